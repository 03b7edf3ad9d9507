"""Attention kernel micro-benchmark at the shapes of the OFA-base bench step (per-task batch 16: merged encoder pass B = 64,
decoder groups B = 48 / 32): CUDA-event device times and TFLOP/s.  A spin kernel queued in front of every timed region lets the
host run ahead, so the event pair brackets device time only (the Python wrapper costs ~80 us per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops

H = 12
REP = 6


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6))
    e0.record()
    for _ in range(REP):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP


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
    outs = []

    def fwd():
        outs.append(ops.attention(q, pq, k, pk, v, tok_lut, img_lut, cs, cfg))

    for _ in range(2):
        fwd()
        outs.pop().backward(do)
    tf = min(timed(fwd) for _ in range(3))
    outs.clear()
    tt = []
    for _ in range(3):       # forward + backward together, the forward's share subtracted
        def fb():
            fwd()
            outs.pop().backward(do)
        tt.append(timed(fb))
    tb = min(tt) - tf
    ff = 2.0 * B * H * T * S * 192 * (0.5 if causal else 1)
    fb_ = 2.0 * B * H * T * S * 512 * (0.5 if causal else 1)
    print("%-34s fwd %7.1f us %6.1f TF/s | bwd %7.1f us %6.1f TF/s" % (name, tf * 1e3, ff / tf / 1e9, tb * 1e3, fb_ / tb / 1e9), flush=True)


run("enc merged B=64 N=835 img+txt", 64, 835, 835, 576, False, True, True)
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.exit(0)
run("enc B=8 N=835 img+txt", 8, 835, 835, 576, False, True, True)
run("enc B=64 N=835 no rel bias", 64, 835, 835, 0, False, False, False)
run("dec self causal B=32 T=250", 32, 250, 250, 0, True, True, False)
run("cross B=32 T=250 S=835", 32, 250, 835, 0, False, False, False)
run("dec self causal B=48 T=12", 48, 12, 12, 0, True, True, False)
run("cross B=48 T=12 S=835", 48, 12, 835, 0, False, False, False)
run("enc text-only B=16 N=185", 16, 185, 185, 0, False, True, False)
