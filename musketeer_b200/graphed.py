"""CUDA-graph training micro-step.  The OFA micro-step is ~3000 small kernel launches (five task forwards, one backward);
replaying it as one graph removes the Python / launch overhead that otherwise hides the GPU work (the
reference additionally blocks the host four times per forward).  One graph per shape signature; inputs are copied into
static device buffers (pinned host -> device, asynchronous) before each replay.

Constraints: shapes are static per signature (the Musketeer loader would bucket / pad lengths), no host-side randomness
inside the step (patch sampling must come in as a `patch_orders` tensor), gradients live in the graph's memory pool and
are overwritten by every replay (the optimizer reads them in place).

The per-batch counts of the collaters (`ntokens`, `nsentences`: data/mm_data/*_dataset.py) are NOT part of the signature: they
become 0-d device tensors among the graph's static inputs, so the criterion's `loss / sample_size` and the gradient scale it
promises to the loss kernel are computed on the device and one capture serves every batch of a shape.  The cache is a small
LRU.  update_freq > 1 (trainer.py:752-773): call with first= / last= -- the replays' gradients are summed in persistent
buffers and handed back through the same .grad views after the last micro-step."""
from collections import OrderedDict

import torch

from .synthetic import map_tensors

COUNT_KEYS = ("ntokens", "nsentences")      # per-batch python ints that must not key the graph cache


def _signature(obj, key=None):
    if isinstance(obj, torch.Tensor):
        return (tuple(obj.shape), str(obj.dtype))
    if isinstance(obj, dict):
        return tuple((k, _signature(v, k)) for k, v in sorted(obj.items()))
    if isinstance(obj, (list, tuple)):
        return tuple(_signature(v) for v in obj)
    if key in COUNT_KEYS and isinstance(obj, (int, float)) and not isinstance(obj, bool):
        return "count"
    return obj if isinstance(obj, (int, float, bool, str, type(None))) else str(type(obj))


def _make_static(obj, device, key=None):
    """Device twins of a host sample: tensors keep shape / dtype, the collater's counts become 0-d fp32 tensors."""
    if isinstance(obj, torch.Tensor):
        return torch.empty(obj.shape, device=device, dtype=obj.dtype)
    if isinstance(obj, dict):
        return {k: _make_static(v, device, k) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_make_static(v, device) for v in obj]
    if key in COUNT_KEYS and isinstance(obj, (int, float)) and not isinstance(obj, bool):
        return torch.empty((), device=device, dtype=torch.float32)
    return obj


def _copy_into(dst, src):
    if isinstance(dst, torch.Tensor):
        if isinstance(src, torch.Tensor):
            dst.copy_(src, non_blocking=True)
        else:
            dst.fill_(float(src))           # a collater count: scalar kernel argument, no host synchronisation
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])
    elif isinstance(dst, list):
        for d, s in zip(dst, src):
            _copy_into(d, s)


class GraphedTrainStep:
    def __init__(self, model, criterion, device, float_dtype, warmup=2, max_graphs=8):
        self.model, self.criterion, self.device, self.float_dtype = model, criterion, device, float_dtype
        self.warmup = warmup
        self.max_graphs = max_graphs
        self.cache = OrderedDict()          # signature -> captured step, least recently used first
        self._sum = None                    # update_freq > 1: running sums of the flat gradient arenas / stray gradients

    def _run(self, static):
        # float inputs are cast to the model dtype on the device, inside the graph (trainer._prepare_sample semantics,
        # trainer.py:1227-1244) -- the staging buffers keep the host dtype so the H2D copy is a plain DMA
        static = map_tensors(static, lambda t: t.to(self.float_dtype) if t.is_floating_point() else t)
        samples = [dict(s, net_input=dict(s["net_input"])) for s in static] if isinstance(static, list) else \
            dict(static, net_input=dict(static["net_input"]))
        from . import ops
        with ops.grad_accumulation(self.model):
            loss, ss, log = self.criterion(self.model, samples)
            loss.backward()
        return loss, ss

    def _capture(self, samples):
        static = _make_static(samples, self.device)
        _copy_into(static, samples)
        # the eager warm-up passes must leave no trace: BatchNorm running statistics (momentum update per forward,
        # models/ofa/resnet.py:142-143) are restored afterwards, so the first batch of a signature updates them once (the replay)
        stats = {n: b.clone() for n, b in self.model.named_buffers()
                 if n.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in self.model.parameters():
                    p.grad = None
                self._run(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for n, b in self.model.named_buffers():
                if n in stats:
                    b.copy_(stats[n])
        for p in self.model.parameters():
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss, ss = self._run(static)
        return {"graph": graph, "static": static, "loss": loss.detach(), "sample_size": ss}

    def prefetch(self, samples):
        """Start the host -> device copy of the NEXT micro-step's batches on a copy stream, into a staging set next to the
        graph's static inputs, while the current step computes (the data loader's job in a trainer).  The following
        __call__ with the same `samples` object then only does device-to-device copies.  No-op before the first capture."""
        e = self.cache.get(_signature(samples))
        if e is None:
            return
        if "staging" not in e:
            e["staging"] = map_tensors(e["static"], lambda t: torch.empty_like(t))
            e["copy_stream"] = torch.cuda.Stream()
            e["staged_ev"] = torch.cuda.Event()
            e["consumed_ev"] = None
        cs = e["copy_stream"]
        if e["consumed_ev"] is not None:
            cs.wait_event(e["consumed_ev"])          # the previous step's device-to-device copy has drained the staging set
        with torch.cuda.stream(cs):
            _copy_into(e["staging"], samples)
            e["staged_ev"].record(cs)
        e["staged_for"] = id(samples)

    def _grad_tensors(self):
        """Every tensor that holds gradients after a replay: the accumulator's flat arenas + the few gradients autograd
        produced outside them."""
        acc = getattr(self.model, "_ofa_grad_acc", None)
        flats = acc.flat_grads() if acc is not None else []
        lo_hi = [(f.data_ptr(), f.data_ptr() + f.numel() * f.element_size()) for f in flats]
        rest = [p.grad for p in self.model.parameters()
                if p.grad is not None and not any(lo <= p.grad.data_ptr() < hi for lo, hi in lo_hi)]
        return flats + rest

    def __call__(self, samples, first=True, last=True):
        """samples: host (ideally pinned) or device tensors in the reference `sample` layout.  Returns (loss tensor, sample_size);
        parameter .grad fields hold the gradients of this micro-step after the call -- or, with first / last marking the
        micro-steps of one update (update_freq > 1), the sum over the micro-steps after the call with last=True."""
        sig = _signature(samples)
        e = self.cache.get(sig)
        if e is None:
            while len(self.cache) >= self.max_graphs:       # bounded: every capture owns a memory pool
                self.cache.popitem(last=False)
            e = self.cache[sig] = self._capture(samples)
        else:
            self.cache.move_to_end(sig)
        self._replay(e, samples)
        if not (first and last):
            g = self._grad_tensors()
            if first or self._sum is None or len(self._sum) != len(g) or any(a.shape != b.shape for a, b in zip(self._sum, g)):
                self._sum = [t.float().clone() for t in g] if first else None
                if self._sum is None:
                    raise RuntimeError("GraphedTrainStep: micro-steps of one update must start with first=True")
            else:
                torch._foreach_add_(self._sum, [t.float() for t in g])
            if last:
                for t, sacc in zip(g, self._sum):
                    t.copy_(sacc)
        return e["loss"], e["sample_size"]

    def _replay(self, e, samples):
        if e.get("staged_for") == id(samples):
            cur = torch.cuda.current_stream()
            cur.wait_event(e["staged_ev"])
            _copy_into(e["static"], e["staging"])
            e["consumed_ev"] = torch.cuda.Event()
            e["consumed_ev"].record(cur)
            e["staged_for"] = None
        else:
            _copy_into(e["static"], samples)
        e["graph"].replay()
