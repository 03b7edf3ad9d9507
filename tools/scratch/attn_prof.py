import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from musketeer_b200 import ops
import os
H, B, T, P = 12, int(os.environ.get("AB", 64)), int(os.environ.get("AT", 835)), int(os.environ.get("AP", 576))
SK = int(os.environ.get("AS", T))
g = torch.Generator(device="cpu").manual_seed(0)
D = H * 64
mk = lambda L, sc: (torch.randn(B, L, D, generator=g) * sc).cuda().bfloat16().requires_grad_()
q, pq, k, pk, v = mk(T, 0.3), mk(T, 0.3), mk(SK, 1.0), mk(SK, 1.0), mk(SK, 1.0)
tok_lut = (torch.randn(H, 2047, generator=g) * 0.5).cuda().requires_grad_() if P else None
img_lut = (torch.randn(H, 83 * 83 + 3, generator=g) * 0.5).cuda().requires_grad_() if P else None
cs = torch.ones(H).cuda().bfloat16().requires_grad_()
ar = torch.arange(P)
pid = ((ar // 24) * 42 + ar % 24 + 1).int().cuda()[None].expand(B, P).contiguous() if P else None
cfg = {"H": H, "causal": False, "kpm": torch.zeros(B, SK, dtype=torch.uint8).cuda(), "q_pos_off": 0,
       "bias": {"q_text_off": P, "k_text_off": P, "ibs": 42, "q_pid": pid, "k_pid": pid, "n_img_q": P, "n_img_k": P}}
do = torch.randn(B, T, D, generator=g).cuda().bfloat16()
for _ in range(2):
    ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg).backward(do)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg).backward(do)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += ev.device_time
        agg[ev.name][1] += 1
for name, (t_, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print("%9.1f us x %3d  %s" % (t_ / n, n, name[:100]))
