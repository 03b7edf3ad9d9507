"""Microbenchmark of the two decode-side kernels at the captioning shapes (BASELINE configs[4]: 64 sentences x beam 5, OFA-base,
480x480 -> S = 908 encoder positions, V = 59457): cross-attention decode (csrc/decode.cu) and the fused beam tail (csrc/beam.cu).
Prints algorithmic bytes / time against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from musketeer_b200 import ops


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    EB, G, H, S, D, V = 64, 5, 12, 908, 768, 59457
    R = EB * G
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.3).to(dev).bfloat16()
    q, k, v = mk(R, 1, D), mk(EB, S, D), mk(EB, S, D)
    kpm = torch.zeros(EB, S, dtype=torch.uint8, device=dev)
    rows = torch.arange(EB, dtype=torch.int32, device=dev)
    bias = torch.randn(R, H, (S + 3) // 4 * 4, generator=g).to(dev)
    hs = torch.ones(H, device=dev)
    t = timeit(lambda: ops.attention_decode(q, None, k, None, v, S, H, G, rows, None, kpm, hs, bias_in=bias))
    by = EB * S * D * 2 * 2
    print(json.dumps({"kernel": "attn_decode cross (G=5, S=908, 64 sentences)", "us": t, "algorithmic_MB": by / 1e6, "GB/s": by / t / 1e3}))
    t = timeit(lambda: ops.attention_decode(q, None, k, None, None, S, H, G, rows, None, kpm, score_out=bias))
    print(json.dumps({"kernel": "attn_decode score_out", "us": t, "GB/s": by / 2 / t / 1e3}))
    # the same bytes with every (sentence, head) stream contiguous (head-major cache layout): is the strided row pattern the limit?
    q1, k1, v1 = mk(EB * H * G, 1, 64), mk(EB * H, S, 64), mk(EB * H, S, 64)
    kpm1 = torch.zeros(EB * H, S, dtype=torch.uint8, device=dev)
    rows1 = torch.arange(EB * H, dtype=torch.int32, device=dev)
    bias1 = torch.randn(EB * H * G, 1, (S + 3) // 4 * 4, generator=g).to(dev)
    t = timeit(lambda: ops.attention_decode(q1, None, k1, None, v1, S, 1, G, rows1, None, kpm1, None, bias_in=bias1))
    print(json.dumps({"kernel": "attn_decode cross, head-major K / V", "us": t, "GB/s": by / t / 1e3}))
    ld = (V + 7) // 8 * 8
    logits = (torch.randn(R, ld, generator=g) * 3).to(dev).bfloat16()[:, :V]
    prev = torch.randn(R, generator=g).to(dev)
    tokens = torch.randint(4, 50000, (R, 18), generator=g).to(dev)
    ws = [None]

    def tail():
        _, _, ws[0] = ops.beam_topk(logits, G, 2 * G, 1.0, prev, eos=2, pad=1, unk=3, tokens=tokens, step=5, ngram=0, ws=ws[0])
    t = timeit(tail)
    by = R * V * 2
    print(json.dumps({"kernel": "beam_topk (320 rows x 59457)", "us": t, "algorithmic_MB": by / 1e6, "GB/s": by / t / 1e3}))


if __name__ == "__main__":
    main()
