"""One eager Musketeer micro-step (per-task batch 16 by default) after one warm-up step: the command profiled under ncu
(launch list with gpu__time_duration, and the --set full capture of the dominant kernel).  Prints the kernel-entry count."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, _lib, ops
from musketeer_b200.optim import FusedAdam
from musketeer_b200.synthetic import build_model, make_tep_group, to_device

ap = argparse.ArgumentParser()
ap.add_argument("--task-batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model, task = build_model("ofa_base", dev, torch.bfloat16)
model.train()
crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
group = to_device(make_tep_group(a.task_batch), dev, torch.bfloat16)
opt = FusedAdam(model.parameters())


def step():
    for p in model.parameters():
        p.grad = None
    with ops.grad_accumulation(model):
        loss, _, _ = crit(model, [dict(g, net_input=dict(g["net_input"])) for g in group])
        loss.backward()
    opt.step()
    return loss


step()
torch.cuda.synchronize()
l0 = _lib.LAUNCHES
for _ in range(a.steps):
    loss = step()
torch.cuda.synchronize()
print("loss %.4f  C-ABI calls per step %d" % (float(loss), (_lib.LAUNCHES - l0) // a.steps))
