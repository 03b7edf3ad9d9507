"""GEMM micro-benchmark on the OFA-base shapes (per-task batch 8): ofa_gemm_bf16 vs torch.matmul (cuBLAS), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops

SHAPES = [  # (name, M, N, K, a_mn, b_mn)
    ("enc b16 fc1 fwd", 13360, 3072, 768, 0, 0), ("enc b16 fc2 fwd", 13360, 768, 3072, 0, 0),
    ("enc b16 fc1 wgrad", 3072, 768, 13360, 1, 1),
    ("enc qkv fwd", 6680, 768, 768, 0, 0), ("enc fused-qkv fwd", 6680, 2304, 768, 0, 0),
    ("enc fc1 fwd", 6680, 3072, 768, 0, 0), ("enc fc2 fwd", 6680, 768, 3072, 0, 0),
    ("enc fc1 dgrad", 6680, 768, 3072, 0, 1), ("enc fc1 wgrad", 3072, 768, 6680, 1, 1),
    ("enc proj wgrad", 768, 768, 6680, 1, 1), ("dec small fwd", 96, 768, 768, 0, 0),
    ("logits fwd", 1856, 59457, 768, 0, 0), ("logits dgrad", 1856, 768, 59457, 0, 1),
    ("logits wgrad", 59457, 768, 1856, 1, 1),
    # 1x1 convolutions of the ResNet-101 stem as GEMMs on NHWC bytes (64 images of 384x384)
    ("l1 conv1 fwd", 589824, 64, 256, 0, 0), ("l1 conv3 fwd", 589824, 256, 64, 0, 0),
    ("l1 conv3 dgrad", 589824, 64, 256, 0, 1), ("l1 conv3 wgrad", 256, 64, 589824, 1, 1),
    ("l2 conv3 fwd", 147456, 512, 128, 0, 0), ("l3 conv1 fwd", 36864, 256, 1024, 0, 0),
    ("l3 conv3 fwd", 36864, 1024, 256, 0, 0), ("l3 conv3 dgrad", 36864, 256, 1024, 0, 1),
    ("l3 conv3 wgrad", 1024, 256, 36864, 1, 1), ("l3 conv1 wgrad", 256, 1024, 36864, 1, 1),
]
from musketeer_b200 import _lib
if "--one" in sys.argv:
    SHAPES = SHAPES[:2]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if "--no-tma-store" in sys.argv:
    _lib.load().ofa_gemm_set_tma_store(0)
if "--dec" in sys.argv:      # decoder groups of the merged task passes (32 x 250 and 48 x 12 target positions) and the encoder pass
    SHAPES = [("dec8000 qkv fwd", 8000, 2304, 768, 0, 0), ("dec8000 proj fwd", 8000, 768, 768, 0, 0),
              ("dec8000 proj dgrad", 8000, 768, 768, 0, 1), ("dec8000 fc1 fwd", 8000, 3072, 768, 0, 0),
              ("dec8000 fc2 fwd", 8000, 768, 3072, 0, 0), ("dec576 proj fwd", 576, 768, 768, 0, 0),
              ("dec320 proj fwd", 320, 768, 768, 0, 0), ("dec320 qkv fwd", 320, 2304, 768, 0, 0),
              ("dec320 fc1 fwd", 320, 3072, 768, 0, 0), ("dec320 fc2 fwd", 320, 768, 3072, 0, 0),
              ("dec576 fc2 fwd", 576, 768, 3072, 0, 0), ("dec576 fc1 dgrad", 576, 768, 3072, 0, 1),
              ("dec576 proj dgrad", 576, 768, 768, 0, 1), ("enc53440 qkv fwd", 53440, 2304, 768, 0, 0),
              ("enc53440 proj fwd", 53440, 768, 768, 0, 0), ("l3 conv1 dgrad", 36864, 1024, 256, 0, 1),
              ("l3 conv3 fwd", 36864, 1024, 256, 0, 0)]
if "--conv" in sys.argv:
    SHAPES = [s for s in SHAPES if "conv" in s[0]]
for _m in (0, 1, 2):
    if "--mode%d" % _m in sys.argv:
        _lib.load().ofa_gemm_set_pair_mode(_m)
for name, M, N, K, a_mn, b_mn in SHAPES:
    K8, M8, N8 = (K + 7) // 8 * 8, (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.randn((K, M8) if a_mn else (M, K8), device="cuda").bfloat16()
    B = torch.randn((K, N8) if b_mn else (N, K8), device="cuda").bfloat16()
    Av = A[:, :M] if a_mn else A[:, :K]
    Bv = B[:, :N] if b_mn else B[:, :K]
    out = torch.empty(M, N8, dtype=torch.bfloat16, device="cuda")[:, :N]
    def ours():
        ops.gemm(Av, Bv, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out)
    Af = (Av.t() if a_mn else Av)
    Bf = (Bv.t() if b_mn else Bv)
    def ref():
        torch.matmul(Af, Bf.t())
    res = []
    for fn in (ours, ref):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res.append(ts[len(ts) // 2])
    fl = 2.0 * M * N * K
    print("%-20s M=%5d N=%5d K=%5d  ours %8.1f us %7.1f TF/s | cuBLAS %8.1f us %7.1f TF/s" %
          (name, M, N, K, res[0] * 1e3, fl / res[0] / 1e9, res[1] * 1e3, fl / res[1] / 1e9))
