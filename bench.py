#!/usr/bin/env python
"""Benchmark of the hot path: OFA-base Musketeer TEP multi-task training micro-step (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--task-batch b] [--impl reference]

One "step" = one Musketeer micro-step: five sequential task forwards (caption / VQA / VG / SNLI-VE / gigaword, per-task
batch b, 384x384 images) through the criterion + one backward, on random-init OFA-base in bf16 (synthetic data).
`value` = samples/s with the batches resident in HBM; `e2e` = the same step including the pinned-host -> device copy of
every batch and the device -> host read of the loss.  N > 1: one rank per GPU, batch-sharded (weak scaling); the step is a
CUDA-graph replay, after which the flat gradient arenas are all-reduced in place over NCCL (one collective per arena,
musketeer_b200.dp.GradReducer.reduce_flat; the eager path overlaps bucketed all-reduces with backward instead).
`--impl reference`: the reference's own CPU implementation of the same step, timed on the host cores with all threads on a
bounded sample (per-task batch 2, the script's): the UNMODIFIED reference modules from baseline/_ref (placed there by
tools/install_reference.py; they travel to the GPU box) through the fairseq stand-ins of oracle/ref_shim -- `kind:
"reference"` -- or, where that tree is absent, the oracle port (`kind: "port"`).  Rank 0 alone runs it under torchrun.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--task-batch", type=int, default=16)
    ap.add_argument("--arch", default="ofa_base")
    ap.add_argument("--img", type=int, default=384)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--ref-task-batch", type=int, default=2, help="per-task batch of the CPU reference arm (BASELINE.md 3)")
    ap.add_argument("--profile-kernels", type=int, default=1)
    ap.add_argument("--graph", type=int, default=1, help="replay the micro-step as a CUDA graph (0: eager launches)")
    ap.add_argument("--caption-bench", type=int, default=1, help="also time beam-5 captioning (BASELINE configs[4]) at N=1")
    ap.add_argument("--script-flags", type=int, default=1,
                    help="also time the micro-step under the flags of run_scripts/musketeer/train_musketeer.sh:56-71,164 at N=1")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, idx):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def cpu_reference(steps, warmup, arch, img, task_batch=2):
    """The reference's own CPU implementation of the path on the host cores: the UNMODIFIED reference modules
    (models/ofa/*.py + criterions/label_smoothed_cross_entropy.py from baseline/_ref or /root/reference, through the fairseq
    stand-ins of oracle/ref_shim) when that tree is present -> kind "reference"; otherwise the oracle port -> kind "port".
    One step = one five-task group at per-task batch `task_batch` (the script's batch 2): forward + loss +
    backward + the trainer's clip + Adam update, fp32, all host threads.  The reference criterion's multi-task recursion needs
    sample_patch_num > 0 (label_smoothed_cross_entropy.py:175-183), so the plain step is driven task by task and combined with
    the recursion's arithmetic, as in oracle/make_golden.py."""
    from oracle import ofa_oracle as oo, synth, ref_harness as rh
    from musketeer_b200.synthetic import make_tep_group
    torch.set_num_threads(os.cpu_count())
    cfg = synth.make_cfg(arch)
    sd = synth.synth_state_dict(cfg, seed=0)
    kind = "reference" if rh.available() else "port"
    if kind == "reference":
        try:
            model, task = rh.build_model(cfg, sd)
            model.train()
            crit = rh.build_criterion(task, label_smoothing=0.1, sample_patch_num=0)
            leaves = [p for p in model.parameters() if p.requires_grad]
        except Exception:             # a broken reference tree must not cost the arm its number: the oracle port always exists
            import traceback
            traceback.print_exc(file=sys.stderr)
            kind = "port"
    if kind == "port":
        sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        sd["decoder.embed_tokens.weight"] = sd["encoder.embed_tokens.weight"]
        sd["decoder.output_projection.weight"] = sd["encoder.embed_tokens.weight"]
        leaves = [v for k, v in sd.items() if v.requires_grad and not k.startswith(("decoder.embed_tokens", "decoder.output_projection"))]
    # the update the trainer runs after the backward (trainer.py:863-898; fairseq Adam, train_musketeer.sh:136): global-norm
    # clip 0.1 + Adam(lr 3e-5, betas (0.9, 0.999), eps 1e-8, decoupled weight decay 0.01) on the fp32 weights
    opt = torch.optim.AdamW(leaves, lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    times = []
    for it in range(warmup + steps):
        group = make_tep_group(task_batch, img=img, seed=it)
        t0 = time.perf_counter()
        if kind == "reference":
            loss = sum(l / ss for l, ss, _ in (crit(model, smp) for smp in group))
        else:
            loss, ss, _ = oo.criterion_forward(sd, cfg, group, epsilon=0.1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in leaves if p.grad is not None], 0.1)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        opt.zero_grad(set_to_none=True)
    t = sum(times) / len(times)
    return 5.0 * task_batch / t, t, os.cpu_count(), kind


def caption_bench(dev, batch=64, img=480, beam=5, iters=6):
    """BASELINE.json configs[4]: beam-5 captioning, OFA-base bf16, `batch` synthetic 480x480 images, prompt
    of 8 source tokens, max_len_b 16 (random weights rarely emit EOS: the worst case of 17 decoder steps) -> captions/s."""
    from musketeer_b200.sequence_generator import SequenceGenerator
    from musketeer_b200.synthetic import build_model
    model, task = build_model("ofa_base", dev, torch.bfloat16, seed=0, patch_image_size=img)
    model.eval()
    gen = SequenceGenerator([model], task.target_dictionary, beam_size=beam, max_len_a=0, max_len_b=16, min_len=1)
    g = torch.Generator(device="cpu").manual_seed(1)
    src = torch.randint(4, 50265, (batch, 8), generator=g)
    src[:, 0], src[:, -1] = 0, 2
    sample = {"net_input": {"src_tokens": src.to(dev), "src_lengths": torch.full((batch,), 8).to(dev),
                            "patch_images": torch.randn(batch, 3, img, img, generator=g).to(dev).bfloat16(),
                            "patch_masks": torch.ones(batch, dtype=torch.bool, device=dev)}}
    times = []
    for it in range(iters + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = gen.generate([model], sample)
        torch.cuda.synchronize()
        if it:
            times.append(time.perf_counter() - t0)
    t = min(times)
    steps = max(len(h[0]["tokens"]) for h in out)
    # HBM roofline of the decode phase (one token per beam: every step streams the per-sentence cross-attention K / V of the
    # six layers, the shared position keys once, the decoder weights incl. the tied vocabulary projection, and writes + reads
    # the logits), against the WHOLE generate latency -- the encoder pass (tensor-bound) is inside the denominator, so the
    # fraction is a lower bound for the decode kernels
    d, S, R = 768, (img // 16) ** 2 + 8, batch * beam
    V = model.decoder.output_projection.weight.shape[0]
    dec_w = sum(p.numel() for p in model.decoder.parameters()) * 2
    step_bytes = 6 * batch * S * d * 2 * 2 + batch * S * d * 2 + dec_w + 2 * R * V * 2
    pk, _ = peaks()
    ach = steps * step_bytes / t / 1e9
    return {"metric": "beam-5 captions/s (OFA-base, %d x %dx%d images, max_len 16)" % (batch, img, img), "value": batch / t,
            "unit": "captions/s", "latency_ms": t * 1e3, "decoder_steps": steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "bytes_per_decoder_step": step_bytes,
                         "note": "decode-phase algorithmic bytes / whole generate latency (encoder pass included): lower bound"}}


def script_flags_bench(dev, arch, img, steps, warmup):
    """The same 5-task micro-step under the flags run_scripts/musketeer/train_musketeer.sh actually passes (:56-71,164):
    dropout 0.1, encoder / decoder drop-path 0.1, --use-rdrop (every task's batch duplicated, KL term between the halves),
    sample_patch_num 196 (every image task but the last keeps a random patch subset of its 576 patches: 392 of them, because
    construct_rdrop_sample doubles every int of the sample, label_smoothed_cross_entropy.py:61-62), at the headline's per-task batch 16
    and at the script's own batch 2.  samples/s counts the un-duplicated samples.  Device-resident inputs, CUDA-graph replay +
    fused Adam inside the timed region."""
    from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, _lib
    from musketeer_b200.graphed import GraphedTrainStep
    from musketeer_b200.optim import FusedAdam
    from musketeer_b200.synthetic import build_model, make_tep_group, to_device
    out = {"flags": "dropout 0.1, encoder/decoder drop-path 0.1, use_rdrop (reg_alpha 1.0), sample_patch_num 196 (392 kept under R-Drop, as in the reference), label smoothing 0.1"}
    for tb in (16, 2):
        model, task = build_model(arch, dev, torch.bfloat16, seed=0, dropout=0.1, encoder_drop_path_rate=0.1,
                                  decoder_drop_path_rate=0.1)
        model.train()
        crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, use_rdrop=True, reg_alpha=1.0, sample_patch_num=196)
        optim = FusedAdam(model.parameters(), lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, clip_norm=0.1)
        graphed = GraphedTrainStep(model, crit, dev, torch.bfloat16)
        groups = [to_device(make_tep_group(tb, img=img, seed=500 + i), dev, torch.bfloat16) for i in range(2)]
        loss = None
        for i in range(warmup):
            loss, _ = graphed(groups[i % 2])
            optim.step()
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss, _ = graphed(groups[i % 2])
            optim.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["task_batch_%d" % tb] = {"value": 5 * tb / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms,
                                     "loss": float(loss), "rows_per_task_with_rdrop": 2 * tb}
        del graphed, model, optim, groups
        torch.cuda.empty_cache()
    return out


def main():
    # the driver reads ONE JSON line from stdout: libraries that print there (NCCL's version banner) are sent to stderr by
    # pointing fd 1 at fd 2 for the whole run; the JSON line goes to the saved descriptor at the end
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = "OFA-base Musketeer TEP 5-task micro-step (caption S137/T12, VQA S230/T232, VG S259/T5, SNLI-VE " \
               "S250/T250, gigaword S185/T12), %dx%d images, per-task batch %d, fwd+loss+bwd+Adam update" % (a.img, a.img, a.task_batch)

    if a.impl == "reference":
        if rank != 0:
            return
        timed_steps = max(1, min(a.steps, 4))      # bounded sample: a step is 10-30 s of host time
        v, t, cores, kind = cpu_reference(timed_steps, min(a.warmup, 1), a.arch, a.img, a.ref_task_batch)
        what = ("the unmodified reference modules through the fairseq stand-ins of oracle/ref_shim" if kind == "reference"
                else "oracle port of the reference (reference tree not present)")
        print(json.dumps({
            "impl": "reference", "metric": "OFA-base train samples/sec", "value": v, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload.replace("per-task batch %d" % a.task_batch,
                                                    "per-task batch %d (bounded CPU sample; the script's batch)" % a.ref_task_batch)},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                             "sample": "one 5-task group at per-task batch %d per step, %d steps timed after %d warm-up, fp32, %s" % (
                                 a.ref_task_batch, timed_steps, min(a.warmup, 1), what)},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    from musketeer_b200 import _lib, ops, AdjustLabelSmoothedCrossEntropyCriterion
    from musketeer_b200.synthetic import build_model, make_tep_group, to_device, batch_bytes
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    model, task = build_model(a.arch, dev, torch.bfloat16, seed=0)
    model.train()
    crit = AdjustLabelSmoothedCrossEntropyCriterion(task, False, 0.1, sample_patch_num=0)
    reducer = None
    if world > 1:
        from musketeer_b200.dp import GradReducer
        reducer = GradReducer(model, world)
    n_batches = 2
    host = [make_tep_group(a.task_batch, img=a.img, seed=rank * 1000 + i, pin=True) for i in range(n_batches)]
    resident = [to_device(h, dev, torch.bfloat16) for h in host]
    h2d = batch_bytes(host[0])

    from musketeer_b200.optim import FusedAdam
    optim = FusedAdam(model.parameters(), lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, clip_norm=0.1)
    graphed = None
    if a.graph:
        from musketeer_b200.graphed import GraphedTrainStep
        graphed = GraphedTrainStep(model, crit, dev, torch.bfloat16)

    def eager_step(group):
        for p in model.parameters():
            p.grad = None
        with ops.grad_accumulation(model):
            loss, ss, log = crit(model, [dict(g, net_input=dict(g["net_input"])) for g in group])
            loss.backward()
        return loss

    def step(group, e2e=False, eager=False):
        """group: device-resident batches (value) or pinned host batches (e2e: H2D copy inside the step, loss read back)."""
        if reducer is not None:
            reducer.prepare()
        if graphed is not None and not eager:
            loss, _ = graphed(group)
        else:
            if e2e:
                group = to_device(group, dev, torch.bfloat16)
            loss = eager_step(group)
        if reducer is not None:
            if graphed is not None and not eager:
                if not reducer.reduce_flat(model):      # gradients live in the accumulator's flat arenas: in-place all-reduce
                    reducer.reduce_all()
            else:
                reducer.finish()
        optim.step()        # update_freq = 1: every micro-step is followed by clip + Adam (trainer.py:863-898)
        return loss

    def phases(nsteps):
        """Device time of the three phases of a step (graph replay | gradient all-reduce | optimizer), events on this rank."""
        acc = [0.0, 0.0, 0.0, 0.0]
        for i in range(nsteps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            graphed(resident[i % n_batches])
            ev[1].record()
            if reducer is not None:
                dist.barrier()            # separates waiting for the slowest rank from the collective itself
            ev[2].record()
            if reducer is not None and not reducer.reduce_flat(model):
                reducer.reduce_all()
            ev[3].record()
            optim.step()
            ev[4].record()
            torch.cuda.synchronize()
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) / nsteps
        return {"replay_ms": acc[0], "rank_skew_wait_ms": acc[1], "allreduce_ms": acc[2], "optimizer_ms": acc[3]}

    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    losses_read = []

    def timed(nsteps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = None
        for i in range(nsteps):
            loss = step((host if e2e else resident)[i % n_batches], e2e)
            if e2e and graphed is not None and i + 1 < nsteps:
                graphed.prefetch(host[(i + 1) % n_batches])     # next step's H2D copy runs under this step's compute
            if e2e:
                # device -> host read of every step's result, one step behind (asynchronous logging, as a trainer does): the
                # copy into pinned memory is queued behind the step, the host reads step i-1's value while step i runs
                buf = loss_host[i % 2]
                buf.copy_(loss.detach().reshape(1).float(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[1].synchronize()
                    losses_read.append(float(pending[0][0]))
                pending = (buf, ev)
        if pending is not None:
            pending[1].synchronize()
            losses_read.append(float(pending[0][0]))
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(a.warmup):
        step(resident[i % n_batches])
        step(host[i % n_batches], e2e=True)      # also captures the graph for the host-dtype signature
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(a.steps, False)
    ms_e2e = timed(a.steps, True)
    clocks = sampler.stop() if sampler else None
    ph = phases(a.steps) if graphed is not None else None
    # kernel launches of one step (counted on an eager step: a graph replay re-issues exactly these launches)
    l0 = _lib.LAUNCHES
    step(resident[0], eager=True)
    launches = (_lib.LAUNCHES - l0) * a.steps
    # per-kernel CUDA-event timing: the same K steps repeated with an event pair around every C-ABI launch (eager, because
    # events cannot bracket nodes inside a graph replay); used for the roofline object only, never for `value`
    prof = None
    if a.profile_kernels:
        # an eager step is host-bound: without a backlog every start event would also time the wait for the next launch.
        # A spin kernel lets the host queue the whole step ahead of the GPU, so the pairs bracket kernel time only; it is
        # sized from the measured host time of one instrumented eager step (host speed differs from box to box).
        _lib.PROFILE = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(resident[0], eager=True)
        host_s = time.perf_counter() - t0
        torch.cuda.synchronize()
        _lib.PROFILE = {}
        for i in range(a.steps):
            torch.cuda._sleep(int(1.3 * host_s * 1.965e9))
            step(resident[i % n_batches], eager=True)
            torch.cuda.synchronize()
        torch.cuda.synchronize()
        prof = _lib.PROFILE
        _lib.PROFILE = None
        ms_prof = sum(r[0].elapsed_time(r[1]) for recs in prof.values() for r in recs)
    if rank != 0:
        return

    samples = 5 * a.task_batch * world
    value = samples * a.steps / (ms / 1e3)
    pk, pk_kind = peaks()
    roof = None
    if prof:
        tot = {}
        for name, recs in prof.items():
            t = sum(r[0].elapsed_time(r[1]) for r in recs)
            w = {}
            for r in recs:
                if r[2]:
                    w[r[2][0]] = w.get(r[2][0], 0.0) + r[2][1]
                    if len(r[2]) > 2 and len(r[2][2]) >= 6:        # GEMM: operands + result bytes, (M, N, K, a_mn, b_mn, out dtype)
                        M_, N_, K_, od = r[2][2][0], r[2][2][1], r[2][2][2], r[2][2][5]
                        w["byte_alg"] = w.get("byte_alg", 0.0) + 2.0 * (M_ * K_ + N_ * K_) + (2.0 if od == 1 else 4.0) * M_ * N_
            tot[name] = (t, w, len(recs))
        top = max(tot, key=lambda k: tot[k][0])
        t, w, n = tot[top]
        share = {k: round(v[0] / ms_prof, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:6]}
        if "flop" in w:
            ach = w["flop"] / (t / 1e3) / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk_kind + " (sustained: timed inside a long step)",
                    "launches": n, "avg_launch_us": t * 1e3 / n, "share_of_library_kernel_time": share}
        else:
            ach = w.get("byte", 0.0) / (t / 1e3) / 1e9
            roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk_kind, "launches": n,
                    "avg_launch_us": t * 1e3 / n, "share_of_library_kernel_time": share}
        if top == "ofa_gemm_bf16" and a.task_batch == 16 and a.arch == "ofa_base" and a.img == 384:
            # DRAM bytes per launch of the same kernel family in the same step, from the committed ncu pass (cannot be read
            # live): profiles/r02_ncu_launch_list_b16_step.txt, 638 gemm_tc* launches, cold-cache and serialised under ncu
            roof["traffic"] = 89.18e6 + 20.97e6
            roof["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's 638 GEMM launches)"
            roof["traffic_source"] = "profiles/r02_ncu_launch_list_b16_step.txt"
            roof["algorithmic_bytes_per_launch"] = w["byte_alg"] / n if w.get("byte_alg") else None
        g = tot.get("ofa_gemm_bf16")
        if g and top != "ofa_gemm_bf16":
            roof["gemm_tflops"] = g[1].get("flop", 0.0) / (g[0] / 1e3) / 1e12
    caption = script = None
    if world == 1 and (a.caption_bench or a.script_flags):
        del graphed, resident
        torch.cuda.empty_cache()
    # the secondary measurements must never cost the headline line: a failure is reported in their slot
    import traceback
    if a.script_flags and world == 1:
        try:
            script = script_flags_bench(dev, a.arch, a.img, a.steps, a.warmup)
        except Exception as e:
            traceback.print_exc(file=sys.stderr)
            script = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if a.caption_bench and world == 1:
        try:
            caption = caption_bench(dev)
        except Exception as e:
            traceback.print_exc(file=sys.stderr)
            caption = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    cpu = None
    if a.cpu_baseline and world == 1:       # (rank 0 at N = 1 only: the other N reuse that figure)
        try:
            v, t, cores, kind = cpu_reference(1, 1, a.arch, a.img, a.ref_task_batch)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                   "sample": "one 5-task group at per-task batch %d (fp32, %s), %.1f s" % (
                       a.ref_task_batch, "unmodified reference via oracle/ref_shim" if kind == "reference" else "oracle port", t)}
        except Exception as e:
            traceback.print_exc(file=sys.stderr)
            cpu = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    print(json.dumps({
        "metric": "OFA-base train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "arch": a.arch, "l2": "activations and weights per step exceed the 126 MB L2",
                   "optimizer_step": "fused clip + Adam (fp32 master weights) after every micro-step, inside the timed region",
                   "dropout": 0.0,
                   "launch": "CUDA graph replay" if a.graph else "eager"},
        "clocks": clocks, "gpu_launches": launches, "phases": ph,
        "e2e": {"value": samples * a.steps / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps,
                "readback": "loss of every step copied to pinned memory and read by the host one step behind",
                "h2d": "every step's batches copied from pinned host memory inside the timed region; the copy of step i+1 is issued on a copy stream while step i computes (the first step's copy is exposed)",
                "last_loss": losses_read[-1] if losses_read else None},
        "roofline": roof, "cpu_baseline": cpu, "caption_beam5": caption, "script_flags": script}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
