"""One attention forward + backward at the encoder shape of the bench step (B = 8 by default, H = 12, N = 835 = 576 image + 259
text positions, image + token relative-position bias): the command profiled under `ncu --set full` for the attention kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops
H, B, T, P = 12, int(os.environ.get("ATTN_B", 8)), 835, 576
g = torch.Generator(device="cpu").manual_seed(0)
D = H * 64
mk = lambda L, sc: (torch.randn(B, L, D, generator=g) * sc).cuda().bfloat16().requires_grad_()
q, pq, k, pk, v = mk(T, 0.3), mk(T, 0.3), mk(T, 1.0), mk(T, 1.0), mk(T, 1.0)
tok_lut = (torch.randn(H, 2047, generator=g) * 0.5).cuda().requires_grad_()
img_lut = (torch.randn(H, 83 * 83 + 3, generator=g) * 0.5).cuda().requires_grad_()
cs = torch.ones(H).cuda().bfloat16().requires_grad_()
ar = torch.arange(P)
pid = ((ar // 24) * 42 + ar % 24 + 1).int().cuda()[None].expand(B, P).contiguous()
cfg = {"H": H, "causal": False, "kpm": torch.zeros(B, T, dtype=torch.uint8).cuda(), "q_pos_off": 0,
       "bias": {"q_text_off": P, "k_text_off": P, "ibs": 42, "q_pid": pid, "k_pid": pid, "n_img_q": P, "n_img_k": P}}
do = torch.randn(B, T, D, generator=g).cuda().bfloat16()
for _ in range(3):
    ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg).backward(do)
torch.cuda.synchronize()
print("ok")
