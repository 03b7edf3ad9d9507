"""Stem-only probe: BN-parameter gradient deviations of (a) this package's bf16 stem and (b) plain torch ops in bf16 on the
GPU, both against torch fp32 (TF32 off), same weights, same image, same random upstream gradient."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import synth, ofa_oracle as oo
from musketeer_b200.resnet import ResNetStem
cfg = synth.make_cfg("ofa_base")
sd = synth.synth_state_dict(cfg, seed=0)
P = "encoder.embed_images."
g = torch.Generator().manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
img = torch.randn(B, 3, 384, 384, generator=g)
dy = torch.randn(B, 576, 1024, generator=g) * 0.01

def run_oracle(dtype):
    s = {k: v.clone().cuda().to(dtype if v.is_floating_point() else v.dtype).requires_grad_(v.is_floating_point()) for k, v in sd.items() if k.startswith(P)}
    f = oo.resnet_forward(s, P[:-1], img.cuda().to(dtype), cfg.resnet_type, training=True)
    f = f.flatten(2).transpose(1, 2)
    f.backward(dy.cuda().to(dtype))
    return {k[len(P):]: v.grad.float().norm().item() for k, v in s.items() if v.grad is not None}, f.detach().float()

ref, fref = run_oracle(torch.float32)
tb, ftb = run_oracle(torch.bfloat16)
stem = ResNetStem(cfg.resnet_type)
stem.load_state_dict({k[len(P):]: v for k, v in sd.items() if k.startswith(P)})
stem = stem.cuda().bfloat16().train()
f = stem(img.cuda().bfloat16())
f.backward(dy.cuda().bfloat16())
ours = {n: p.grad.float().norm().item() for n, p in stem.named_parameters()}
print("feature rel err: torch-bf16 %.3e  ours %.3e" % (((ftb - fref).norm() / fref.norm()).item(), ((f.detach().float() - fref).norm() / fref.norm()).item()))
import collections
agg = collections.defaultdict(list)
for n, r in ref.items():
    kind = "bn" if ".bn" in n or "downsample.1" in n or n.startswith("bn1") else "conv"
    agg[kind].append((abs(ours[n] - r) / r, abs(tb[n] - r) / r, n))
for kind, rows in agg.items():
    o = sorted(x[0] for x in rows); t = sorted(x[1] for x in rows)
    print("%-5s n=%3d  ours: median %.3e p90 %.3e max %.3e | torch-bf16: median %.3e p90 %.3e max %.3e" % (
        kind, len(rows), o[len(o) // 2], o[int(len(o) * .9)], o[-1], t[len(t) // 2], t[int(len(t) * .9)], t[-1]))
for r in sorted(agg["bn"], reverse=True)[:6]:
    print("   %-45s ours %.3e torch-bf16 %.3e" % (r[2], r[0], r[1]))
