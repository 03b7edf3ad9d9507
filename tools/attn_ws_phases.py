"""Phase accounting of the warp-specialised attention forward (library built with -DOFA_WS_DEBUG, path in OFA_B200_LIB)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops, _lib

H, B, N, P = 12, 64, 835, 576
g = torch.Generator(device="cpu").manual_seed(0)
D = H * 64
mk = lambda L, sc: (torch.randn(B, L, D, generator=g) * sc).cuda().bfloat16()
q, pq, k, pk, v = mk(N, 0.3), mk(N, 0.3), mk(N, 1.0), mk(N, 1.0), mk(N, 1.0)
tok = (torch.randn(H, 2047, generator=g) * 0.5).cuda()
img = (torch.randn(H, 83 * 83 + 3, generator=g) * 0.5).cuda()
ar = torch.arange(P)
pid = ((ar // 24) * 42 + ar % 24 + 1).int().cuda()[None].expand(B, P).contiguous()
cfg = {"H": H, "causal": False, "kpm": torch.zeros(B, N, dtype=torch.uint8).cuda(), "q_pos_off": 0,
       "bias": {"q_text_off": P, "k_text_off": P, "ibs": 42, "q_pid": pid, "k_pid": pid, "n_img_q": P, "n_img_k": P}}
lib = _lib.load()
lib.ofa_attn_ws_debug_read.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_longlong * 64)()
with torch.no_grad():
    for _ in range(3):
        ops.attention(q, pq, k, pk, v, tok, img, None, cfg)
    torch.cuda.synchronize()
    lib.ofa_attn_ws_debug_read(buf)
    ops.attention(q, pq, k, pk, v, tok, img, None, cfg)
    torch.cuda.synchronize()
    lib.ofa_attn_ws_debug_read(buf)
d = list(buf)
names = {0: "loop top", 1: "wait s_full", 2: "tmem_ld S + arrive s_free", 3: "bias + local max", 4: "named barrier + partner max",
         5: "exp + pack + rowsum", 6: "wait o_full(j-1)", 7: "fold O", 8: "P store + arrive p_full",
         20: "issuer: wait Q'(tmem) + K'(0)", 21: "issuer: issue S(0) x2", 22: "issuer: wait k_full(j+1)",
         23: "issuer: wait s_free[0]", 24: "issuer: issue S_0", 25: "issuer: wait s_free[1]", 26: "issuer: issue S_1",
         27: "issuer: wait p_full[0]", 28: "issuer: issue PV_0", 29: "issuer: wait p_full[1]", 30: "issuer: issue PV_1"}
print("CTA (1,3,5): start->main %d  main loop (softmax warp 0) %d  end %d cycles" % (d[40] - d[39], d[41] - d[40], d[42] - d[41]))
for i in sorted(names):
    print("%-40s %8d cycles  (%6.0f per iteration)" % (names[i], d[i], d[i] / 14.0))
