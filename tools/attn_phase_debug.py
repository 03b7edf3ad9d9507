"""Phase timing of one attention-backward CTA (clock64 around the barriers; needs the -DOFA_ATTN_DEBUG build of the library)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import _lib, ops
_lib.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libofa_dbg.so"))
H, B, T, P = 12, 8, 835, 576
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
lib = _lib.load()
buf = (ctypes.c_longlong * 128)()
for it in range(3):
    o = ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg)
    o.backward(do)
    torch.cuda.synchronize()
    lib.ofa_attn_debug_read(buf)
names = ["loop top + row metadata", "wait S/dP (issuer: Q' landed, S, drain, dP)", "tmem ld issue", "softmax math + slab wait + smem", "named barrier", "hist atomics", "wait dQ' MMA", "dQ drain", "named barrier",
         "arrive bar_pds", "-", "-"]
print("CTA (key tile 2: image keys), thread 0, microseconds per phase and query tile (1.9 GHz):")
print("%-24s" % "phase" + "".join("%8d" % i for i in range(7)))
order = [0, 1, 2, 3, 4, 9, 5, 6, 7, 8]
for ph in order:
    print("%-24s" % names[ph] + "".join("%8.2f" % (buf[i * 12 + ph] / 1900.0) for i in range(7)))
print("%-24s" % "total" + "".join("%8.2f" % (sum(buf[i * 12 + ph] for ph in range(12)) / 1900.0) for i in range(7)))
print("prologue %.2f us | prologue + main loop %.2f us | epilogue %.2f us   (accumulated over %d backward calls: divide by 1)" % (
    buf[96] / 1900.0, buf[97] / 1900.0, buf[98] / 1900.0, 1))
