"""Data parallelism on the device (1-GPU vs N-GPU gradient equality): two ranks -- two processes on cuda:0, gloo
carrying the CUDA tensors, so the test runs on the single-GPU CI box; the 8-GPU NCCL path is what bench.py --gpus N runs --
each take half of a micro-step's batch through the real product path (multi-task criterion, merged task passes, fused
gradient accumulation into the flat arenas, GradReducer.reduce_flat in place), and must end up with the gradients one process
computes on the whole batch (trainer.py:755-773,848-852; `_check_grad_norms` :1397-1433: identical on every rank)."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    from oracle import synth
    cfg = synth.make_cfg("ofa_micro", vocab_size=4099, freeze_resnet=True)      # frozen BatchNorm: no per-rank batch statistics
    sd = synth.synth_state_dict(cfg, seed=0)
    # no padding: both halves of every task carry the same number of target tokens, so mean(rank gradients) == full-batch gradient
    spec = [(19, 7, True), (30, 9, True), (25, 6, False)]
    full = [synth.make_batch(4, s, t, img=64, seed=80 + i, vocab=4099, n_pad=0, with_image=im) for i, (s, t, im) in enumerate(spec)]
    return cfg, sd, full


def _slice(sample, lo, hi):
    def f(v):
        if isinstance(v, torch.Tensor) and v.dim() > 0:
            return v[lo:hi].clone()
        if isinstance(v, dict):
            return {k: f(x) for k, x in v.items()}
        return v
    out = f(sample)
    n = hi - lo
    out["nsentences"] = n
    out["ntokens"] = int(out["target"].ne(1).sum())
    return out


def _grads(model, samples):
    from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion, ops
    from tests.helpers import to_device
    crit = AdjustLabelSmoothedCrossEntropyCriterion(model._task, False, 0.1, use_rdrop=False, sample_patch_num=0)
    for p in model.parameters():
        p.grad = None
    with ops.grad_accumulation(model):
        loss, ss, _ = crit(model, to_device(copy.deepcopy(samples), "cuda"))
        loss.backward()
    return float(loss.detach())


def _worker(rank, world, port, outdir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from musketeer_b200.dp import GradReducer
    from tests.helpers import build_product
    cfg, sd, full = _case()
    model, task = build_product(cfg, sd, dtype=torch.float32)
    model.train()
    model._task = task
    red = GradReducer(model, world)
    half = [_slice(s, 2 * rank, 2 * rank + 2) for s in full]
    _grads(model, half)                       # lays out the arenas (first use)
    red.prepare()
    loss = _grads(model, half)
    assert red.reduce_flat(model)
    out = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}
    torch.save((loss, out), os.path.join(outdir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process_on_the_whole_batch(tmp_path):
    from tests.helpers import build_product
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(2)]
    cfg, sd, full = _case()
    model, task = build_product(cfg, sd, dtype=torch.float32)
    model.train()
    model._task = task
    loss = _grads(model, full)
    ref = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    assert abs(0.5 * (res[0][0] + res[1][0]) - loss) <= 1e-5 * abs(loss)
    for n in set(res[0][1]) - set(ref):      # parameters this batch does not reach are reduced as zeros (identical collectives on every rank)
        assert float(res[0][1][n].abs().sum()) == 0.0, n
    assert set(ref) <= set(res[0][1])
    gn = sum(float(v.norm()) ** 2 for v in ref.values()) ** 0.5
    for n, v in ref.items():
        assert torch.equal(res[0][1][n], res[1][1][n]), n                       # identical on both ranks
        assert (res[0][1][n] - v).abs().max().item() <= 2e-5 * max(1.0, gn), n   # == the whole-batch gradient
