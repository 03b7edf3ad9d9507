"""Data-parallel gradient reduction (musketeer_b200.dp.GradReducer) on CPU with the gloo backend, world size 2:
bucketed all-reduce == mean of the per-rank gradients, tied weights reduced once, parameters without gradient reduced
as zeros, no_sync() defers the reduction (trainer.py:755-773 semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(16, 32)
        self.b = torch.nn.Linear(32, 16)
        self.unused = torch.nn.Parameter(torch.ones(7))
        self.tied = torch.nn.Linear(16, 16, bias=False)
        self.tied.weight = self.a.weight if False else self.tied.weight

    def forward(self, x):
        return self.tied(self.b(torch.relu(self.a(x)))).sum()


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from musketeer_b200.dp import GradReducer
    model = _Toy()
    red = GradReducer(model, world, bucket_bytes=1024)
    assert len(red.buckets) > 2
    torch.manual_seed(100 + rank)
    x = torch.randn(8, 16)
    # reference: local grads
    model.zero_grad(set_to_none=True)
    model(x).backward()
    local = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reduced
    model.zero_grad(set_to_none=True)
    red.prepare()
    model(x).backward()
    red.finish()
    out = {n: (p.grad.clone() if p.grad is not None else None) for n, p in model.named_parameters()}
    # no_sync keeps local grads
    model.zero_grad(set_to_none=True)
    with red.no_sync():
        red.prepare()
        model(x).backward()
        red.finish()
    nosync = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reduce_all (CUDA-graph path: hooks do not fire)
    model.zero_grad(set_to_none=True)
    model(x).backward()
    red.reduce_all()
    allred = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    # reduce_flat: gradients are views into flat arenas (ops.grad_accumulation layout) -> in-place flat all-reduce
    model.zero_grad(set_to_none=True)
    model(x).backward()
    named = [(n, p) for n, p in model.named_parameters() if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for _, p in named])
    off = 0
    for _, p in named:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()

    class _Acc:
        def flat_grads(self):
            return [flat]
    model._ofa_grad_acc = _Acc()
    model.unused.grad = torch.zeros_like(model.unused)        # unused parameter: zeros outside the arenas are accepted
    assert red.reduce_flat(model, chunk_bytes=512)
    flatred = {n: p.grad.clone() for n, p in named}
    # a gradient outside the arenas is reduced through the temporary flat buffer
    model.unused.grad = torch.full_like(model.unused, float(rank + 1))
    assert red.reduce_flat(model, chunk_bytes=512)
    assert torch.allclose(model.unused.grad, torch.full_like(model.unused, 1.5))
    for n, p in named:          # (the arenas were averaged a second time: still the mean of equal values)
        assert torch.allclose(p.grad, flatred[n], atol=1e-6), n
    # DistributedOFAModel: the wrapper trainer.py drives (forward / no_sync / all_reduce_grads / .module), update_freq = 2
    from musketeer_b200.dp import DistributedOFAModel
    m2 = _Toy()
    wrapped = DistributedOFAModel(m2, world_size=world, bucket_bytes=1024)
    assert wrapped.module is m2 and wrapped.a is m2.a                           # attribute passthrough
    assert list(wrapped.state_dict().keys()) == list(m2.state_dict().keys())   # no "module." prefix in checkpoints
    x2 = torch.randn(8, 16)
    m2.zero_grad(set_to_none=True)
    m2(x).backward()
    m2(x2).backward()
    local2 = {n: p.grad.clone() for n, p in m2.named_parameters() if p.grad is not None}
    m2.zero_grad(set_to_none=True)
    with wrapped.no_sync():                      # trainer.py:755-773: first micro-batch accumulates locally
        wrapped(x).backward()
    wrapped(x2).backward()                       # last micro-batch: buckets are reduced under the backward
    wrapped.all_reduce_grads()                   # trainer.py:848-852
    ddp2 = {n: p.grad.clone() for n, p in m2.named_parameters() if p.grad is not None}
    torch.save((local, out, nosync, allred, flatred, local2, ddp2), os.path.join(outdir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_world2(tmp_path):
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = {r: torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(2)}
    for n in res[0][0]:
        mean = (res[0][0][n] + res[1][0][n]) / 2
        for r in range(2):
            assert torch.allclose(res[r][1][n], mean, atol=1e-6), n       # hook-driven bucketed reduce
            assert torch.allclose(res[r][3][n], mean, atol=1e-6), n       # reduce_all
            assert torch.allclose(res[r][4][n], mean, atol=1e-6), n       # reduce_flat
            assert torch.allclose(res[r][2][n], res[r][0][n]), n          # no_sync: untouched local grads
    for n in res[0][5]:
        mean2 = (res[0][5][n] + res[1][5][n]) / 2
        for r in range(2):
            assert torch.allclose(res[r][6][n], mean2, atol=1e-6), n      # wrapper: (g1 + g2) averaged over the ranks
    for r in range(2):
        assert res[r][1]["unused"] is not None and float(res[r][1]["unused"].abs().sum()) == 0.0
        # both ranks hold identical gradients afterwards (trainer._check_grad_norms invariant, trainer.py:1397-1433)
    for n in res[0][1]:
        if res[0][1][n] is not None:
            assert torch.equal(res[0][1][n], res[1][1][n]), n
