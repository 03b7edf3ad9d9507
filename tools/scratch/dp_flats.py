import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, ops
from musketeer_b200.synthetic import build_model, make_tep_group, to_device
dev = torch.device("cuda", 0)
model, task = build_model("ofa_base", dev, torch.bfloat16)
model.train()
crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
group = to_device(make_tep_group(2), dev, torch.bfloat16)
for _ in range(2):
    for p in model.parameters():
        p.grad = None
    with ops.grad_accumulation(model):
        loss, _, _ = crit(model, [dict(g, net_input=dict(g["net_input"])) for g in group])
        loss.backward()
acc = model._ofa_grad_acc
flats = acc.flat_grads()
print("flats:", [(tuple(f.shape), f.dtype, f.numel() * f.element_size() / 1e6) for f in flats])
lo_hi = [(f.data_ptr(), f.data_ptr() + f.numel() * f.element_size()) for f in flats]
rest = [(n, p.numel()) for n, p in model.named_parameters() if p.requires_grad and (p.grad is None or not any(lo <= p.grad.data_ptr() < hi for lo, hi in lo_hi))]
print("rest: %d params, %.3f M elements" % (len(rest), sum(n for _, n in rest) / 1e6))
for n, k in sorted(rest, key=lambda x: -x[1])[:12]:
    print("   ", n, k)
