"""Per-kernel device-time breakdown of one eager Musketeer micro-step via torch.profiler (CUPTI activity records:
covers the ctypes-launched kernels of libofa_b200.so as well as library kernels).  Writes a table to gpurun_out/."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, ops
from musketeer_b200.synthetic import build_model, make_tep_group, to_device

ap = argparse.ArgumentParser()
ap.add_argument("--task-batch", type=int, default=8)
ap.add_argument("--out", default="gpurun_out/breakdown.txt")
ap.add_argument("--ops", action="store_true", help="also list ATen operators by input shape and Python call site")
a = ap.parse_args()
dev = torch.device("cuda", 0)
model, task = build_model("ofa_base", dev, torch.bfloat16)
model.train()
crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
group = to_device(make_tep_group(a.task_batch), dev, torch.bfloat16)


def step():
    for p in model.parameters():
        p.grad = None
    with ops.grad_accumulation(model):
        loss, _, _ = crit(model, [dict(g, net_input=dict(g["net_input"])) for g in group])
        loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += ev.device_time
        agg[ev.name][1] += 1
tot = sum(v[0] for v in agg.values())
lines = ["one eager micro-step, per-task batch %d: %.2f ms device time in %d kernels" % (a.task_batch, tot / 1e3, sum(v[1] for v in agg.values()))]
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:160]:
    lines.append("%8.3f ms %5.1f%% %6d x %8.1f us  %s" % (t / 1e3, 100 * t / tot, n, t / n, name[:110]))
os.makedirs(os.path.dirname(a.out), exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

if a.ops:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof2:
        step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof2.key_averages(group_by_input_shape=True, group_by_stack_n=6):
        if ev.self_device_time_total > 0 and ev.key.startswith("aten::"):
            stack = [f for f in ev.stack if "musketeer_b200/" in f or "bench" in f][:3]
            rows.append((ev.self_device_time_total, ev.count, ev.key, str(ev.input_shapes)[:70], " <- ".join(x.split("/")[-1][:60] for x in stack)))
    rows.sort(reverse=True)
    out = ["", "ATen / autograd operators with device time (self), by input shape and call site:"]
    for t, n, k, shp, st in rows[:70]:
        out.append("%8.3f ms %5d x  %-34s %-70s %s" % (t / 1e3, n, k[:34], shp, st))
    open(a.out, "a").write("\n".join(out) + "\n")
    print("\n".join(out))

    # who launches the big elementwise kernels?  (CPU op -> parents)
    out = ["", "large ATen elementwise / copy kernels and the operators that launched them:"]
    for ev in prof2.events():
        ks = getattr(ev, "kernels", None) or []
        for k in ks:
            if k.duration > 150 and ("elementwise" in k.name or "copy" in k.name.lower() or "Fill" in k.name):
                chain, e = [], ev
                while e is not None and len(chain) < 6:
                    chain.append("%s%s" % (e.name[:40], str(e.input_shapes)[:60] if e.input_shapes else ""))
                    e = e.cpu_parent
                out.append("%8.1f us  %s  <=  %s" % (k.duration, k.name[:60], "  <-  ".join(chain)))
    open(a.out, "a").write("\n".join(out) + "\n")
    print("\n".join(out))
