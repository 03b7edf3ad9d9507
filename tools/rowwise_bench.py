"""HBM-bound kernels of the OFA step against the copy roofline: BatchNorm (ResNet stem shapes of the b=16 micro-step,
4 task groups) and LayerNorm(+GELU) on the FFN rows.  Per shape: algorithmic bytes / CUDA-event time, L2 flushed between
launches.   python tools/rowwise_bench.py [--bn-sweep] [--ln]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musketeer_b200 import ops, _lib

lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(400000)          # let the host run ahead: the launches below are queued back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


# (name, images, H, C, relu, residual, count in the stem)
BN_SHAPES = [
    ("stem bn1 192^2 x64", 64, 192, 64, 1, 0, 1),
    ("l1 bn1/2 96^2 x64", 64, 96, 64, 1, 0, 6),
    ("l1 bn3 96^2 x256 +res", 64, 96, 256, 1, 1, 3),
    ("l1 ds 96^2 x256", 64, 96, 256, 0, 0, 1),
    ("l2 bn1 96^2 x128", 64, 96, 128, 1, 0, 1),
    ("l2 bn1/2 48^2 x128", 64, 48, 128, 1, 0, 7),
    ("l2 bn3 48^2 x512 +res", 64, 48, 512, 1, 1, 4),
    ("l3 bn1 48^2 x256", 64, 48, 256, 1, 0, 1),
    ("l3 bn1/2 24^2 x256", 64, 24, 256, 1, 0, 45),
    ("l3 bn3 24^2 x1024 +res", 64, 24, 1024, 1, 1, 23),
]


def bn_bench(tag):
    tot_f = tot_b = 0.0
    print("== BatchNorm (%s) ==" % tag)
    for name, n, h, c, relu, has_res, count in BN_SHAPES:
        x = torch.randn(n, c, h, h, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        res = torch.randn_like(x).contiguous(memory_format=torch.channels_last) if has_res else None
        dy = torch.randn_like(x).contiguous(memory_format=torch.channels_last)
        gamma = torch.rand(c, device="cuda").bfloat16() + 0.5
        beta = torch.randn(c, device="cuda").bfloat16() * 0.1
        rm, rv = torch.zeros(c, device="cuda").bfloat16(), torch.ones(c, device="cuda").bfloat16()
        xr = x.detach().requires_grad_(True)
        rr = res.detach().requires_grad_(True) if has_res else None
        g = gamma.detach().requires_grad_(True)
        b = beta.detach().requires_grad_(True)
        out = [None]

        def fwd():
            out[0] = ops.batch_norm(xr, g, b, rm, rv, residual=rr, relu=bool(relu), training=True, groups=4)

        ins = [xr, g, b] + ([rr] if has_res else [])

        def bwd():
            torch.autograd.grad(out[0], ins, dy, retain_graph=True)
        tf = timeit(fwd)
        fwd()
        tb = timeit(bwd)
        nbytes = x.numel() * 2
        bf = (3 + has_res) * nbytes                     # stats read + apply read/write (+ residual read)
        bb = (5 + 2 * has_res) * nbytes                 # stats: x, dy (+y); apply: x, dy (+y) -> dx (+dres)
        print("%-26s %6.1f MB  fwd %7.1f us %5.2f TB/s | bwd %7.1f us %5.2f TB/s  (x%d)" %
              (name, nbytes / 1e6, tf, bf / tf / 1e6, tb, bb / tb / 1e6, count))
        tot_f += tf * count
        tot_b += tb * count
        del x, res, dy, xr, rr, out
    print("stem total: fwd %.2f ms, bwd %.2f ms" % (tot_f / 1e3, tot_b / 1e3))


def ln_bench(tag):
    print("== LayerNorm (%s) ==" % tag)
    for name, rows, c, gelu in [("enc ffn LN(GELU) 3072", 53840, 3072, True), ("dec ffn LN(GELU) 3072", 7712, 3072, True),
                                ("enc LN 768", 53840, 768, False), ("dec LN 768", 7712, 768, False)]:
        x = torch.randn(rows, c, device="cuda").bfloat16().requires_grad_(True)
        w = (torch.rand(c, device="cuda").bfloat16() + 0.5).requires_grad_(True)
        b = torch.zeros(c, device="cuda").bfloat16().requires_grad_(True)
        dy = torch.randn(rows, c, device="cuda").bfloat16()
        out = [None]

        def fwd():
            out[0] = ops.layer_norm(x, w, b, gelu_in=gelu)

        def bwd():
            torch.autograd.grad(out[0], [x, w, b], dy, retain_graph=True)
        tf = timeit(fwd)
        fwd()
        tb = timeit(bwd)
        nbytes = rows * c * 2
        print("%-24s %6.1f MB  fwd %7.1f us %5.2f TB/s | bwd %7.1f us %5.2f TB/s" %
              (name, nbytes / 1e6, tf, 2 * nbytes / tf / 1e6, tb, 3 * nbytes / tb / 1e6))


if __name__ == "__main__":
    if "--ln" in sys.argv or "--all" in sys.argv:
        ln_bench("staged backward")
    if "--bn-sweep" in sys.argv or "--all" in sys.argv:
        for waves, u in ((0, 4), (2, 2)):
            lib.ofa_batchnorm_set_tuning(waves, u)
            bn_bench("waves %d, backward rows in flight %d" % (waves, u))
    elif "--bn" in sys.argv:
        bn_bench("default")
