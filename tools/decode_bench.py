"""Beam-search captioning benchmark (BASELINE.json configs[4] / SURVEY.md 8d C5): OFA-base, bf16, eval mode, B synthetic
480x480 images, prompt " what does the image describe?" (8 source tokens), beam 5, max_len_b 16, min_len 1 -> captions/s.
Random-init weights rarely emit EOS, so the search runs the worst case of 17 decoder steps."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from musketeer_b200.sequence_generator import SequenceGenerator
from musketeer_b200.synthetic import build_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--img", type=int, default=480)
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--profile", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model, task = build_model("ofa_base", dev, torch.bfloat16, seed=0, patch_image_size=a.img)
    model.eval()
    gen = SequenceGenerator([model], task.target_dictionary, beam_size=a.beam, max_len_a=0, max_len_b=16, min_len=1)
    g = torch.Generator(device="cpu").manual_seed(1)
    src = torch.randint(4, 50265, (a.batch, 8), generator=g)
    src[:, 0], src[:, -1] = 0, 2
    sample = {"net_input": {"src_tokens": src.to(dev), "src_lengths": torch.full((a.batch,), 8).to(dev),
                            "patch_images": torch.randn(a.batch, 3, a.img, a.img, generator=g).to(dev).bfloat16(),
                            "patch_masks": torch.ones(a.batch, dtype=torch.bool, device=dev)}}
    times = []
    for it in range(a.iters + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = gen.generate([model], sample)
        torch.cuda.synchronize()
        if it:
            times.append(time.perf_counter() - t0)
    if a.profile:
        import collections
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            gen.generate([model], sample)
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                agg[ev.name][0] += ev.device_time
                agg[ev.name][1] += 1
        tot = sum(v[0] for v in agg.values())
        print("device time %.2f ms in %d kernels" % (tot / 1e3, sum(v[1] for v in agg.values())))
        for name, (t_, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            print("%8.3f ms %5.1f%% %6d x %8.1f us  %s" % (t_ / 1e3, 100 * t_ / tot, n, t_ / n, name[:110]))
    t = sorted(times)[len(times) // 2]
    lens = [len(h[0]["tokens"]) for h in out]
    print(json.dumps({"metric": "beam-5 captions/s", "value": a.batch / t, "latency_ms": t * 1e3, "batch": a.batch,
                      "beam": a.beam, "img": a.img, "mean_len": sum(lens) / len(lens)}))


if __name__ == "__main__":
    main()
