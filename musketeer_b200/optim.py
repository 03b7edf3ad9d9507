"""Fused Adam with fp32 master weights and global-norm clipping for the data-parallel training loop.

Restates what the reference's trainer drives per update (trainer.py:863-898: `optimizer.multiply_grads`,
`clip_grad_norm`, `optimizer.step`) with fairseq's Adam + FP16Optimizer arithmetic (un-vendored; flags from
run_scripts/musketeer/train_musketeer.sh:136: adam, betas (0.9, 0.999), eps 1e-8, weight decay 0.01, clip-norm 0.1):
the whole model is updated by two kernel launches of libofa_b200.so (`ofa_adam_step`) reading a device-resident chunk
table; nothing is synchronised with the host (the gradient norm stays on the device).
"""
import struct

import torch

from . import _lib
from .ops import _dt, _p, _st

CHUNK = 65536


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _flat(t):
    """1-D view of a dense tensor in storage order."""
    return t.as_strided((t.numel(),), (1,))


class FusedAdam:
    def __init__(self, params, lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, clip_norm=0.1):
        seen, self.params = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("FusedAdam: no trainable parameters")
        p0 = self.params[0]
        if not p0.is_cuda:
            raise _lib.OfaKernelError("FusedAdam needs CUDA parameters (no CPU fallback exists)")
        if any(p.dtype != p0.dtype or p.device != p0.device or not _dense(p) for p in self.params):
            raise ValueError("FusedAdam: parameters must share dtype and device and be dense in memory")
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm = lr, betas, eps, weight_decay, clip_norm
        self.step_count = 0
        n = sum(p.numel() for p in self.params)
        dev = p0.device
        # fp32 state in three flat buffers (master | exp_avg | exp_avg_sq), 12 bytes per parameter
        self.master = torch.empty(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets, off = [], 0
        for p in self.params:
            self.master[off:off + p.numel()].copy_(_flat(p.detach()).float())      # storage order (channels_last weights)
            self.offsets.append(off)
            off += p.numel()
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._table = self._partial = None
        self._sig = None

    def _build_table(self):
        sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in self.params)
        if sig == self._sig:
            return
        recs = bytearray()
        mb, ab, vb = self.master.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        es = self.params[0].element_size()
        n_chunks = 0
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            if g.dtype != p.dtype or g.stride() != p.stride():
                raise ValueError("FusedAdam: gradients must have the parameter's dtype and memory layout")
            for c0 in range(0, p.numel(), CHUNK):
                cn = min(CHUNK, p.numel() - c0)
                recs += struct.pack("PPPPPq", p.data_ptr() + c0 * es, g.data_ptr() + c0 * es, mb + (off + c0) * 4,
                                    ab + (off + c0) * 4, vb + (off + c0) * 4, cn)
                n_chunks += 1
        dev = self.params[0].device
        self._table = torch.frombuffer(recs, dtype=torch.uint8).to(dev)
        self._partial = torch.empty(n_chunks, dtype=torch.float32, device=dev)
        self._n_chunks = n_chunks
        self._sig = sig

    def step(self, grad_scale=1.0):
        """One update over every parameter.  Parameters whose .grad is None get a zero gradient (fairseq semantics: the
        moments still decay).  Returns the device tensor holding the global gradient norm (scaled, before clipping)."""
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            elif p.grad.stride() != p.stride():
                g2 = torch.empty_like(p)          # the parameter's memory layout
                g2.copy_(p.grad)
                p.grad = g2
        self._build_table()
        self.step_count += 1
        _lib.call("ofa_adam_step", _p(self._table), self._n_chunks, _p(self._partial), _p(self.grad_norm), float(self.lr),
                  float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count,
                  float(grad_scale), float(self.clip_norm or 0.0), _dt(self.params[0]), _st(),
                  work=("byte", 30.0 * self.master.numel()))
        # the kernel wrote the parameters through raw pointers: tell autograd's version counters, which key the caches derived
        # from the weights (MultiheadAttention.qkv_eval's concatenated projection, the generator's captured decoder steps)
        torch.autograd.graph.increment_version(self.params)
        return self.grad_norm

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
