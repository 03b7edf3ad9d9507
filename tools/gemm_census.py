"""Which GEMM shapes the micro-step spends its time in: one eager step with an event pair around every ofa_gemm_bf16 call,
aggregated by (M, N, K, a_mn, b_mn, out) -- launches, total time, TFLOP/s."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, _lib, ops
from musketeer_b200.synthetic import build_model, make_tep_group, to_device

dev = torch.device("cuda", 0)
model, task = build_model("ofa_base", dev, torch.bfloat16)
model.train()
crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
group = to_device(make_tep_group(16), dev, torch.bfloat16)


def step():
    for p in model.parameters():
        p.grad = None
    with ops.grad_accumulation(model):
        loss, _, _ = crit(model, [dict(g, net_input=dict(g["net_input"])) for g in group])
        loss.backward()


step(); step()
torch.cuda.synchronize()
_lib.PROFILE = {}
step()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for e0, e1, work in _lib.PROFILE.get("ofa_gemm_bf16", []):
    a = agg[work[2]]
    a[0] += e0.elapsed_time(e1); a[1] += 1; a[2] += work[1]
tot = sum(v[0] for v in agg.values())
print("ofa_gemm_bf16: %.2f ms in %d calls, %.1f TFLOP/s overall" % (tot, sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values()) / tot / 1e9))
for k, (t, n, f) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("M=%6d N=%6d K=%6d a_mn=%d b_mn=%d out=%d  %4d x %7.1f us = %6.2f ms  %6.1f TFLOP/s" % (*k, n, t / n * 1e3, t, f / t / 1e9))
