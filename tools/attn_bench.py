"""Attention kernel micro-benchmark at the OFA-base bench shapes (per-task batch 8): CUDA-event times and TFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops

H = 12
def run(name, B, T, S, P, causal, tok, img):
    g = torch.Generator(device="cpu").manual_seed(0)
    D = H * 64
    mk = lambda L, sc: (torch.randn(B, L, D, generator=g) * sc).cuda().bfloat16().requires_grad_()
    q, pq, k, pk, v = mk(T, 0.3), mk(T, 0.3), mk(S, 1.0), mk(S, 1.0), mk(S, 1.0)
    tok_lut = (torch.randn(H, 2047, generator=g) * 0.5).cuda().requires_grad_() if tok else None
    img_lut = (torch.randn(H, 83 * 83 + 3, generator=g) * 0.5).cuda().requires_grad_() if img else None
    cs = torch.ones(H).cuda().bfloat16().requires_grad_()
    pid = None
    if P:
        ar = torch.arange(P)
        pid = ((ar // 24) * 42 + ar % 24 + 1).int().cuda()[None].expand(B, P).contiguous()
    cfg = {"H": H, "causal": causal, "kpm": torch.zeros(B, S, dtype=torch.uint8).cuda(), "q_pos_off": 0,
           "bias": {"q_text_off": P, "k_text_off": P, "ibs": 42, "q_pid": pid, "k_pid": pid, "n_img_q": P, "n_img_k": P}}
    do = torch.randn(B, T, D, generator=g).cuda().bfloat16()
    def fwd():
        return ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg)
    for _ in range(2):
        o = fwd(); o.backward(do)
    torch.cuda.synchronize()
    tf, tb = [], []
    for _ in range(5):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); o = fwd(); e[1].record(); o.backward(do); e[2].record(); torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    tf.sort(); tb.sort()
    ff = 2.0 * B * H * T * S * 192 * (0.5 if causal else 1)
    fb = 2.0 * B * H * T * S * 512 * (0.5 if causal else 1)
    print("%-28s fwd %7.1f us %6.1f TF/s | bwd %7.1f us %6.1f TF/s" % (name, tf[2] * 1e3, ff / tf[2] / 1e9, tb[2] * 1e3, fb / tb[2] / 1e9))

run("enc img+txt bias (N=835)", 8, 835, 835, 576, False, True, True)
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.exit(0)
run("enc no rel bias", 8, 835, 835, 0, False, False, False)
run("enc tok bias only", 8, 835, 835, 0, False, True, False)
run("dec self causal T=232", 8, 232, 232, 0, True, True, False)
run("cross T=232 S=835", 8, 232, 835, 0, False, False, False)
run("cross T=12 S=713", 8, 12, 713, 0, False, False, False)
run("enc text-only N=185", 8, 185, 185, 0, False, True, False)
