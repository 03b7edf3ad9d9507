import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, _lib, ops
from tests.helpers import load_golden, build_case, build_product, to_device
lib = _lib.load()
fx = load_golden("base_tep_b2")
cfg, sd, samples = build_case(fx["case"])
res = {}
for mode in (0, 1, 0):
    lib.ofa_gemm_set_small64(mode)
    model, task = build_product(cfg, sd, dtype=torch.bfloat16)
    model.train()
    crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
    loss, ss, _ = crit(model, to_device(copy.deepcopy(samples), "cuda", torch.bfloat16))
    (loss / ss).backward()
    g = {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}
    key = "%d%s" % (mode, "b" if ("%d" % mode) in res else "")
    res[key] = (float(loss), g)
    print("mode", mode, "loss", float(loss), "gnorm", sum(float(v.norm()) ** 2 for v in g.values()) ** 0.5)
def diff(a, b, flt):
    rows = []
    for n in res[a][1]:
        if flt(n):
            x, y = res[a][1][n], res[b][1][n]
            d = float((x - y).norm()) / (float(y.norm()) + 1e-30)
            rows.append((d, n))
    rows.sort(reverse=True)
    return rows
for flt, nm in ((lambda n: "embed_images" not in n, "transformer"), (lambda n: "embed_images" in n, "stem")):
    r01 = diff("1", "0", flt)
    r00 = diff("0b", "0", flt)
    print(nm, "on-vs-off: median %.3e max %.3e (%s) | off-vs-off (run to run): median %.3e max %.3e" % (
        r01[len(r01) // 2][0], r01[0][0], r01[0][1], r00[len(r00) // 2][0], r00[0][0]))
    for d, n in r01[:8]:
        print("    %.3e  %s" % (d, n))
