import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musketeer_b200 import ops, _lib
lib = _lib.load()
g = torch.Generator(device="cpu").manual_seed(0)
for (M, N, K) in [(12, 768, 768), (80, 768, 768), (250, 768, 768), (250, 2304, 768), (250, 3072, 768), (250, 768, 3072), (576, 768, 768), (576, 768, 3072), (5, 768, 768), (232, 768, 768)]:
    for b_mn in (False, True):
        A = (torch.randn(M, K, generator=g)).cuda().bfloat16()
        B = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.05).cuda().bfloat16()
        bias = torch.randn(N, generator=g).cuda().bfloat16()
        resid = torch.randn(M, N, generator=g).cuda().bfloat16()
        ref = (A.double() @ (B.double() if b_mn else B.double().t()))
        for name, kw, fn in (("plain", {}, lambda r: r),
                             ("bias+resid", {"bias": bias, "resid": resid}, lambda r: r + bias.double() + resid.double()),
                             ("bias+gelu", {"bias": bias, "act": 1}, lambda r: torch.nn.functional.gelu(r + bias.double())),
                             ("alpha_cols", {"bias": bias, "alpha": 0.125, "alpha_cols": N // 3 // 64 * 64 or 64}, None)):
            res = []
            for mode in (0, 1):
                lib.ofa_gemm_set_small64(mode)
                y = ops.gemm(A, B, M, N, K, b_mn=b_mn, **kw).double()
                if fn is not None:
                    r = fn(ref)
                else:
                    r = ref + bias.double()
                    c = kw["alpha_cols"]
                    r[:, :c] *= 0.125
                res.append(((y - r).abs().max().item(), (y - r).abs().mean().item(), (y - r).mean().item()))
            flag = "  <<<" if res[1][1] > 1.3 * res[0][1] + 1e-6 or abs(res[1][2]) > 3 * abs(res[0][2]) + 1e-4 else ""
            print("M=%4d N=%4d K=%4d b_mn=%d %-11s  split: max %.3e mean %.3e bias %+.2e | 128x64: max %.3e mean %.3e bias %+.2e%s" % (
                M, N, K, b_mn, name, *res[0], *res[1], flag))
