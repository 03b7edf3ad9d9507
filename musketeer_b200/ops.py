"""torch.autograd wrappers around the C ABI (include/ofa_b200.h).  PyTorch supplies device memory, streams and the
autograd tape only; every arithmetic step of the hot path below runs in libofa_b200.so.

Two numeric modes, selected by the activation dtype:
  * bfloat16 : operands bf16, fp32 accumulation (tcgen05 GEMMs, flash attention) -- the training mode;
  * float32  : "parity mode" -- GEMM operands are split into three bf16 terms (x = x0+x1+x2 to 2^-24) and contracted
               as six K-blocks on the same tcgen05 kernel; attention runs on the exact SIMT kernels.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import OfaAttnArgs, OfaAttnBias, OfaAttnGrads, OfaDecodeArgs, call

F32, BF16 = 0, 1

# Gradient-accumulation fusion.  A Musketeer micro-step sums five task losses and runs ONE backward, so every parameter
# receives five gradient contributions that autograd adds pairwise (~3000 tiny add kernels, 8 ms at per-task batch 8).
# Inside `with grad_accumulation(model):` the kernels that produce parameter gradients (wgrad GEMM epilogue, colsum,
# LayerNorm reduce, embedding scatter-add) write / accumulate straight into one persistent buffer per parameter and hand
# autograd `None`; on exit the buffers become `param.grad`.  Outside the context everything goes through autograd as usual
# (required when gradient hooks must observe each accumulation, e.g. the eager DDP overlap path).
def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class _GradAccumulator:
    def __init__(self, model):
        self.params = {id(p): p for p in model.parameters() if p.requires_grad}
        self.buf = {}
        self.written = set()
        # the per-parameter buffers of the other producers (bias column sums, LayerNorm / BatchNorm reductions, embedding
        # scatter-add) are carved out of one flat arena in the parameter dtype, so a data-parallel run all-reduces two
        # flat tensors in place instead of ~1000 gradients (dp.GradReducer.reduce_flat)
        self.arena_buf = None
        self.buf_used = 0
        # weight gradients of the Linear / 1x1-conv GEMMs accumulate in fp32: every K slice, micro-batch and task adds into
        # one arena by TMA reduce (ofa_gemm_bf16 out_dtype 2); `finish` casts the arena to the parameter dtype in one pass
        self.arena32 = self.arena_out = None
        self.region = {}           # id(param) -> (offset, numel)
        self.arena_used = 0
        self.written32 = set()
        self.continuing = set()    # parameters whose .grad aliases the arenas at `begin` (gradient accumulation over micro-steps)
        self._buf_live = {}        # id(param) -> the bf16 buffer holds a contribution that belongs to the current .grad

    def target32(self, param):
        """-> fp32 [N, K] accumulation view for a registered >= 2-D leaf parameter, else None."""
        k = id(param)
        dense = param.is_contiguous() or (param.dim() == 4 and param.is_contiguous(memory_format=torch.channels_last))
        if k not in self.params or not dense:
            return None
        if self.arena32 is None or self.arena32.device != param.device or self.arena_out.dtype != param.dtype:
            total = sum((p.numel() + 63) // 64 * 64 for p in self.params.values())
            self.arena32 = torch.zeros(total, dtype=torch.float32, device=param.device)
            self.arena_out = torch.empty(total, dtype=param.dtype, device=param.device)
            self.region, self.arena_used = {}, 0
        r = self.region.get(k)
        if r is None:
            if _capturing():
                raise RuntimeError("grad accumulation arena must be laid out by an eager warm-up step before graph capture")
            r = self.region[k] = (self.arena_used, param.numel())
            self.arena_used += (param.numel() + 63) // 64 * 64
        self.written32.add(k)
        v = self.arena32[r[0]:r[0] + r[1]]
        return v.view(param.shape[0], -1) if param.dim() >= 2 else v      # storage order of the parameter

    def group32(self, params):
        """fp32 accumulation view spanning several parameters whose arena regions are adjacent (laid out together on first
        use).  Returns None if they cannot be adjacent (already placed apart, sizes not multiples of 64)."""
        if any(id(p) not in self.params or not p.is_contiguous() or p.numel() % 64 for p in params):
            return None
        placed = [id(p) in self.region for p in params]
        if not all(placed):
            if any(placed):
                return None
            for p in params:           # first use: consecutive offsets (target32 appends at the end of the arena)
                if self.target32(p) is None:
                    return None
        offs = [self.region[id(p)][0] for p in params]
        for p, o, o2 in zip(params, offs, offs[1:] + [None]):
            if o2 is not None and o + p.numel() != o2:
                return None
        for p in params:
            self.written32.add(id(p))
        n = sum(p.numel() for p in params)
        return self.arena32[offs[0]:offs[0] + n]

    def target(self, param):
        """-> (buffer, accumulate) for a registered leaf parameter, else (None, False)."""
        k = id(param)
        if k not in self.params:
            return None, False
        b = self.buf.get(k)
        if b is None or b.shape != param.shape or b.dtype != param.dtype or b.device != param.device:
            dense = param.is_contiguous() or (param.dim() == 4 and param.is_contiguous(memory_format=torch.channels_last))
            if self.arena_buf is None or self.arena_buf.device != param.device or self.arena_buf.dtype != param.dtype:
                self.arena_buf = torch.zeros(sum((p.numel() + 63) // 64 * 64 for p in self.params.values()),
                                             dtype=param.dtype, device=param.device)
                self.buf, self.buf_used = {}, 0
            if dense and not _capturing() and k not in self.region:
                # (a parameter that already owns an fp32 region -- the tied embedding: output-projection wgrad GEMM + embedding
                # scatter-add -- gets its bf16 buffer OUTSIDE the arena: `finish` folds it into the fp32 region's cast, so it
                # would only add 91 MB of already-merged bytes to the data-parallel all-reduce of the arena)
                n = param.numel()       # same memory format as the parameter (channels_last conv weights)
                b = self.arena_buf[self.buf_used:self.buf_used + n].as_strided(param.shape, param.stride())
                self.buf_used += (n + 63) // 64 * 64
            else:
                b = torch.empty_like(param)
            self.buf[k] = b
        first = k not in self.written
        self.written.add(k)
        if first and k in self.continuing and self._buf_live.get(k):
            first = False                   # the buffer holds the sum of earlier micro-steps: keep adding
        elif first:
            self._buf_live[k] = False       # about to be overwritten by this step's first contribution
        return b, not first

    def _aliases(self, k, p):
        """True when p.grad IS this accumulator's view of parameter k (left there by a previous `finish`): the caller kept
        the gradients of an earlier micro-step (update_freq > 1, trainer.py:752-773) and this step must ADD to them."""
        g = p.grad
        if g is None:
            return False
        r = self.region.get(k)
        if r is not None and self.arena_out is not None and g.dtype == self.arena_out.dtype and \
                g.data_ptr() == self.arena_out.data_ptr() + r[0] * self.arena_out.element_size():
            return True
        b = self.buf.get(k)
        return b is not None and g.data_ptr() == b.data_ptr()

    def begin(self):
        self.written, self.written32 = set(), set()
        # Parameters whose .grad still aliases the arenas carry the running sum of earlier micro-steps: their fp32 regions are
        # NOT zeroed (the arena keeps the running fp32 sum, `finish` re-casts it) and their bf16 buffers are accumulated into
        # from the first write on.  Everything else starts from zero.  (Under CUDA-graph capture gradients are None.)
        self.continuing = set()
        if not _capturing():
            self.continuing = {k for k, p in self.params.items() if p.grad is not None and self._aliases(k, p)}
        self._buf_live = {k: True for k in self.continuing if self._buf_live.get(k)}
        if self.arena32 is not None and self.arena_used:
            if not self.continuing:
                self.arena32[:self.arena_used].zero_()
            else:
                for k, (off, n) in self.region.items():
                    if k not in self.continuing:
                        self.arena32[off:off + n].zero_()

    def finish(self):
        if self.written32 or (self.continuing and self.arena_used):
            self.arena_out[:self.arena_used].copy_(self.arena32[:self.arena_used])      # one cast for the whole model
        cont = self.continuing
        for k in self.written32 | {k for k in cont if k in self.region}:
            p = self.params[k]
            off, n = self.region[k]
            g = self.arena_out[off:off + n].as_strided(p.shape, p.stride())
            if k in self.written or (k in cont and k in self.buf and self._buf_live.get(k)):
                g.add_(self.buf[k])          # e.g. the tied embedding: GEMM part + scatter-add part
            p.grad = g if (p.grad is None or k in cont) else p.grad + g
        for k in self.written - self.written32 - {k for k in cont if k in self.region}:
            p, b = self.params[k], self.buf[k]
            p.grad = b if (p.grad is None or k in cont) else p.grad + b
        for k in self.written:
            self._buf_live[k] = True
        self.written, self.written32 = set(), set()

    def flat_grads(self):
        """The flat tensors every gradient produced inside the context lives in (valid after `finish`)."""
        out = []
        if self.arena_out is not None and self.arena_used:
            out.append(self.arena_out[:self.arena_used])
        if self.arena_buf is not None and self.buf_used:
            out.append(self.arena_buf[:self.buf_used])
        return out


_ACC = None


class grad_accumulation:
    def __init__(self, model):
        acc = getattr(model, "_ofa_grad_acc", None)
        if acc is None:
            acc = model._ofa_grad_acc = _GradAccumulator(model)
        self.acc = acc

    def __enter__(self):
        global _ACC
        self.prev, _ACC = _ACC, self.acc
        self.acc.begin()
        return self.acc

    def __exit__(self, et, ev, tb):
        global _ACC
        _ACC = self.prev
        if et is None:
            self.acc.finish()
        return False


def _acc_target(param):
    if _ACC is None or param is None:
        return None, False
    return _ACC.target(param)


def _acc_target32(param):
    if _ACC is None or param is None:
        return None
    return _ACC.target32(param)


def _acc_group32(params):
    """One contiguous fp32 arena region covering `params` back to back (fused q|k|v weight / bias gradients), or None."""
    if _ACC is None:
        return None
    return _ACC.group32(params)


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("musketeer_b200 kernels support float32 and bfloat16, got %s" % t.dtype)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(t):
    if not t.is_cuda:
        raise _lib.OfaKernelError("musketeer_b200 ops need CUDA tensors (no CPU fallback exists)")


def _ceil8(n):
    return (n + 7) // 8 * 8


# ---------------------------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------------------------
def _split3(x2d, rows, cols, k_is_cols, pattern):
    """fp32 [rows, cols] (row stride x2d.stride(0)) -> bf16 six-block operand.  k_is_cols: contraction runs along
    columns (K-major operand) -> blocks side by side [rows, 6*Kp]; else along rows (MN-major) -> stacked [6*Kp, cols8]."""
    if k_is_cols:
        kp = _ceil8(cols)
        out = torch.zeros(rows, 6 * kp, dtype=torch.bfloat16, device=x2d.device)
        call("ofa_split3_bf16", _p(x2d), x2d.stride(0), rows, cols, _p(out), 6 * kp, kp, pattern, _st())
        return out, 6 * kp, 6 * kp
    kp = _ceil8(rows)
    c8 = _ceil8(cols)
    out = torch.zeros(6 * kp, c8, dtype=torch.bfloat16, device=x2d.device)
    call("ofa_split3_bf16", _p(x2d), x2d.stride(0), rows, cols, _p(out), c8, kp * c8, pattern, _st())
    return out, c8, 6 * kp


_NO_SPLITK = bool(int(os.environ.get("OFA_GEMM_NO_SPLITK", "0")))     # A/B switch: no split-K workspace -> small GEMMs run unsplit


def gemm(A, B, M, N, K, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, alpha=1.0, act=0, resid=None,
         ldd=None, acc32=False, rowsum=None, alpha_cols=0):
    """D[M,N] = act((A.B^T + bias) * alpha) + resid.   A: [M,K] (or [K,M] if a_mn); B: [N,K] (or [K,N] if b_mn);
    2-D tensors with unit inner stride.  fp32 operands take the split path."""
    _need_cuda(A)
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    out_dtype = out_dtype or A.dtype
    if out is None:
        ldd = ldd or N
        out = torch.empty(M, ldd, dtype=out_dtype, device=A.device)
    ldd = out.stride(0)
    Kk = K
    if A.dtype == torch.float32:
        A, lda, Kk = _split3(A, A.shape[0], A.shape[1], not a_mn, 0)
        B, ldb, Kb = _split3(B, B.shape[0], B.shape[1], not b_mn, 1)
        assert Kb == Kk
    else:
        lda, ldb = A.stride(0), B.stride(0)
    od = F32 if out_dtype == torch.float32 else BF16
    if acc32:                   # out (fp32) += alpha * A.B^T   (TMA reduce-add; K slices need no workspace)
        assert out is not None and out.dtype == torch.float32 and bias is None and resid is None and act == 0
        od = 2
        if rowsum is not None:  # fp32 [M] += alpha * row sums of A (bias gradient), computed inside the same GEMM
            assert rowsum.dtype == torch.float32 and rowsum.numel() == M and rowsum.is_contiguous()
            bias = rowsum
    elif bias is not None:
        assert bias.dtype == out_dtype
    if resid is not None:
        assert resid.dtype == out_dtype and resid.stride(-1) == 1
    wsb = 0 if (acc32 or _NO_SPLITK) else _lib.load().ofa_gemm_workspace_bytes(M, N, Kk, 1)
    ws = torch.empty(wsb // 4, dtype=torch.float32, device=A.device) if wsb > 0 else None
    call("ofa_gemm_bf16", _p(A), _p(B), _p(out), M, N, Kk, 1, lda, ldb, ldd, 0, 0, 0, int(a_mn), int(b_mn), od,
         _p(bias), float(alpha), int(act), _p(resid), resid.stride(0) if resid is not None else 0, 0, _p(ws), wsb,
         int(alpha_cols), _st(),
         work=("flop", 2.0 * M * N * K, (M, N, K, int(a_mn), int(b_mn), od)))
    return out


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, alpha, resid, out_pad, fork=False):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        M, K = x2.shape
        N = w.shape[0]
        ctx.w_param = w                       # leaf parameter (a 1x1 conv passes its [Cout, Cin, 1, 1] weight)
        w = w.reshape(N, -1)
        r2 = resid.reshape(-1, N) if resid is not None else None
        ldd = _ceil8(N) if out_pad else N
        if ldd == N:
            # the result is allocated in its final shape and handed out as is (not as a view made inside this Function): callers
            # may modify it in place -- the reference criterion masks the logits it gets back (label_smoothed_cross_entropy.py:233-236)
            yn = torch.empty(*shp[:-1], N, dtype=x2.dtype, device=x2.device)
            gemm(x2, w, M, N, K, bias=b, alpha=alpha, resid=r2, out=yn.view(M, N))
        else:
            yn = gemm(x2, w, M, N, K, bias=b, alpha=alpha, resid=r2, ldd=ldd)[:, :N].reshape(*shp[:-1], N)
        ctx.save_for_backward(x2, w)
        ctx.bias_param = b
        ctx.alpha, ctx.has_b, ctx.has_r, ctx.shp = alpha, b is not None, resid is not None, shp
        ctx.set_materialize_grads(False)
        if fork:        # second output = x: the branch that bypasses this projection; its gradient is added in the dgrad epilogue
            return yn, x
        return yn

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        if dy is None:
            return dskip, None, None, None, None, None, None
        dy2 = dy.reshape(-1, N)
        if dy2.stride(-1) != 1 or (dy2.stride(0) % 8 != 0 and dy2.dtype == torch.bfloat16):
            pad = torch.zeros(M, _ceil8(N), dtype=dy2.dtype, device=dy2.device)
            pad[:, :N].copy_(dy2)
            dy2 = pad[:, :N]
        dx = dw = db = dr = None
        bias_done = False
        if ctx.needs_input_grad[0]:
            sk2 = None
            if dskip is not None:
                sk2 = dskip.reshape(-1, K)
                if sk2.stride(-1) != 1 or (sk2.stride(0) % 8 != 0 and sk2.dtype == torch.bfloat16):
                    sk2 = sk2.contiguous()
            dx = gemm(dy2, w, M, K, N, a_mn=False, b_mn=True, alpha=ctx.alpha, resid=sk2).reshape(ctx.shp)
        elif dskip is not None:
            dx = dskip
        if ctx.needs_input_grad[1]:
            tgt32 = _acc_target32(ctx.w_param) if K % 4 == 0 else None
            tgt, accum = (None, False) if tgt32 is not None else _acc_target(ctx.w_param)
            if tgt32 is not None:
                b32 = None
                if ctx.has_b and ctx.needs_input_grad[2] and x2.dtype == torch.bfloat16:
                    b32 = _acc_target32(ctx.bias_param)          # bias gradient = row sums of dY^T, from the same GEMM
                gemm(dy2, x2, N, K, M, a_mn=True, b_mn=True, alpha=ctx.alpha, out=tgt32, out_dtype=torch.float32, acc32=True,
                     rowsum=b32)
                bias_done = b32 is not None
            elif tgt is not None:
                tgt = tgt.view(N, K)
                gemm(dy2, x2, N, K, M, a_mn=True, b_mn=True, alpha=ctx.alpha, out=tgt, out_dtype=w.dtype,
                     resid=tgt if accum else None)
            else:
                dw = gemm(dy2, x2, N, K, M, a_mn=True, b_mn=True, alpha=ctx.alpha, out_dtype=w.dtype)
                dw = dw.view(ctx.w_param.shape)
        if ctx.has_b and ctx.needs_input_grad[2] and not bias_done:
            tgt, accum = _acc_target(ctx.bias_param)
            if tgt is not None:
                colsum(dy2, alpha=ctx.alpha, out=tgt, accumulate=accum)
            else:
                db = colsum(dy2, alpha=ctx.alpha)
        if ctx.has_r and ctx.needs_input_grad[4]:
            dr = dy
        return dx, dw, db, None, dr, None, None


class _FusedLinear(torch.autograd.Function):
    """Several projections of the same input as ONE GEMM over the concatenated weights: q|k|v of a self-attention
    ([3D, d]) or k|v of a cross-attention ([2D, d]); forward, dgrad and wgrad are one launch each instead of one per
    projection (+ the adds of the separate input gradients).  With `scaling`, the first projection is the query: the
    softmax scaling multiplies its columns in the forward epilogue (alpha_cols) and dq on its way out of the attention
    backward (dq_scale), so the arithmetic per element is that of q = (x Wq^T + bq) * s, k = x Wk^T + bk, v = x Wv^T + bv
    (unify_multihead_attention.py:213-232).  The gradients arrive side by side in one buffer when the attention backward
    wrote them there (cfg fused_qkv / fused_kv); otherwise they are concatenated."""

    @staticmethod
    def forward(ctx, x, scaling, *params):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        M, K = x2.shape
        n = len(params) // 2
        D = params[0].shape[0]
        w = torch.cat(params[0::2], 0)
        b = torch.cat(params[1::2], 0)
        if scaling is not None:
            y = gemm(x2, w, M, n * D, K, bias=b, alpha=scaling, alpha_cols=D)
        else:
            y = gemm(x2, w, M, n * D, K, bias=b)
        ctx.save_for_backward(x2, w)
        ctx.params = params
        ctx.shp, ctx.D, ctx.n = shp, D, n
        y = y.view(*shp[:-1], n * D)
        return tuple(y[..., i * D:(i + 1) * D] for i in range(n))

    @staticmethod
    def backward(ctx, *dys):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        D, n = ctx.D, ctx.n
        d0 = dys[0]
        es = d0.element_size()
        fused = (d0.dim() == 3 and d0.stride(-1) == 1 and d0.stride(-2) == n * D and d0.stride(0) == d0.shape[1] * n * D and
                 all(d.stride() == d0.stride() and d.data_ptr() == d0.data_ptr() + i * D * es for i, d in enumerate(dys)))
        if fused:       # written side by side by the attention backward
            dy2 = d0.as_strided((M, n * D), (n * D, 1))
        else:
            dy2 = torch.cat([d.reshape(M, D) for d in dys], 1)
        dx = gemm(dy2, w, M, K, n * D, a_mn=False, b_mn=True).reshape(ctx.shp) if ctx.needs_input_grad[0] else None
        params = ctx.params
        grads = [None] * (2 * n)
        t32 = _acc_group32(list(params[0::2])) if x2.dtype == torch.bfloat16 else None
        b32 = _acc_group32(list(params[1::2])) if t32 is not None else None
        if t32 is not None and b32 is not None:
            gemm(dy2, x2, n * D, K, M, a_mn=True, b_mn=True, out=t32.view(n * D, K), out_dtype=torch.float32, acc32=True,
                 rowsum=b32)
        else:
            dw = gemm(dy2, x2, n * D, K, M, a_mn=True, b_mn=True, out_dtype=params[0].dtype)
            db = colsum(dy2)
            for i in range(n):
                grads[2 * i], grads[2 * i + 1] = dw[i * D:(i + 1) * D], db[i * D:(i + 1) * D]
        return (dx, None, *grads)


def qkv_linear(x, q_proj, k_proj, v_proj, scaling):
    """-> (q * scaling, k, v) as views of one [.., 3D] buffer (q_proj / k_proj / v_proj: nn.Linear parameter holders)."""
    return _FusedLinear.apply(x, scaling, q_proj.weight, q_proj.bias, k_proj.weight, k_proj.bias, v_proj.weight, v_proj.bias)


def kv_linear(x, k_proj, v_proj):
    """-> (k, v) as views of one [.., 2D] buffer (cross-attention keys / values of the encoder output)."""
    return _FusedLinear.apply(x, None, k_proj.weight, k_proj.bias, v_proj.weight, v_proj.bias)


def linear(x, w, b=None, alpha=1.0, resid=None, out_pad=False, fork=False):
    """(x W^T + b) * alpha + resid   (nn.Linear forward/backward through ofa_gemm_bf16).  fork=True returns (y, x): use
    the second value for a branch that bypasses the projection, its gradient is then added in the dgrad GEMM epilogue."""
    return _Linear.apply(x, w, b, alpha, resid, out_pad, fork)


class _Subsample2(torch.autograd.Function):
    """x[:, :, ::2, ::2] of a channels_last tensor as one gather pass; the backward writes zeros and the scattered gradient in
    one pass (csrc/pool.cu) instead of autograd's fill + strided copy."""

    @staticmethod
    def forward(ctx, x):
        N, Cc, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((N, Cc, (H + 1) // 2, (W + 1) // 2), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        call("ofa_subsample2", _p(x), _p(y), N, H, W, Cc, x.element_size(), 0, _st(), work=("byte", 2 * y.numel() * x.element_size()))
        ctx.shape = (N, Cc, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, Cc, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, Cc, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        call("ofa_subsample2", _p(dy), _p(dx), N, H, W, Cc, dy.element_size(), 1, _st(), work=("byte", dx.numel() * dy.element_size()))
        return dx


def conv1x1(x, weight, stride=1, fork=False):
    """1x1 convolution (no bias) on a channels_last [N, C, H, W] tensor = the GEMM [N*H*W, Cin] x [Cout, Cin]^T on the
    NHWC bytes (models/ofa/resnet.py:105-126: conv1 / conv3 / downsample of every bottleneck).  Returns a channels_last
    [N, Cout, H', W'] tensor; with fork=True (stride 1) also the input again, for the identity / downsample branch of the
    bottleneck: that branch's gradient is added in the epilogue of this convolution's dgrad GEMM."""
    _need_cuda(x)
    if stride != 1:
        assert not fork
        if stride == 2 and (x.shape[1] * x.element_size()) % 16 == 0:
            x = _Subsample2.apply(x)
        else:
            x = x[:, :, ::stride, ::stride]
    x = x.contiguous(memory_format=torch.channels_last)
    N, Cin, H, W = x.shape
    x2 = x.permute(0, 2, 3, 1).reshape(N * H * W, Cin)
    if fork:
        y, xs = linear(x2, weight, fork=True)
        return y.view(N, H, W, weight.shape[0]).permute(0, 3, 1, 2), xs.view(N, H, W, Cin).permute(0, 3, 1, 2)
    y = linear(x2, weight)
    return y.view(N, H, W, weight.shape[0]).permute(0, 3, 1, 2)


class _Conv3x3(torch.autograd.Function):
    """3x3 / stride 1 / padding 1 convolution as an implicit GEMM (csrc/conv.cu).  x: channels_last [N, Cin, H, W] bf16;
    w: [Cout, Cin, 3, 3] stored channels_last, i.e. the bytes are [Cout][3][3][Cin]."""

    @staticmethod
    def forward(ctx, x, w):
        _need_cuda(x)
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        x = x.contiguous(memory_format=torch.channels_last)
        wc = w.detach().contiguous(memory_format=torch.channels_last)
        y = torch.empty((N, Cout, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        call("ofa_conv3x3_bf16", _p(x), _p(wc), _p(y), N, H, W, Cin, Cout, 0, _st(),
             work=("flop", 2.0 * N * H * W * Cout * Cin * 9))
        ctx.save_for_backward(x, wc)
        ctx.w_param = w
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wc = ctx.saved_tensors
        N, Cin, H, W = x.shape
        Cout = wc.shape[0]
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x, memory_format=torch.channels_last)
            call("ofa_conv3x3_bf16", _p(dy), _p(wc), _p(dx), N, H, W, Cin, Cout, 1, _st(),
                 work=("flop", 2.0 * N * H * W * Cout * Cin * 9))
        if ctx.needs_input_grad[1]:
            w = ctx.w_param
            tgt32 = _acc_target32(w) if w.is_contiguous(memory_format=torch.channels_last) else None
            if tgt32 is not None:       # fp32 arena region in the weight's storage order [Cout][3][3][Cin]
                call("ofa_conv3x3_wgrad_bf16", _p(x), _p(dy), _p(tgt32), N, H, W, Cin, Cout, 2, _p(None), 0, _st(),
                     work=("flop", 2.0 * N * H * W * Cout * Cin * 9))
                return dx, None
            tgt, accum = _acc_target(w)
            if tgt is not None and not tgt.is_contiguous(memory_format=torch.channels_last):
                tgt = None
            out = tgt if tgt is not None else torch.empty_like(wc, memory_format=torch.channels_last)
            wsb = _lib.load().ofa_conv3x3_wgrad_workspace_bytes(N, H, W, Cin, Cout)
            ws = torch.empty(wsb // 4, dtype=torch.float32, device=x.device)
            call("ofa_conv3x3_wgrad_bf16", _p(x), _p(dy), _p(out), N, H, W, Cin, Cout, int(tgt is not None and accum),
                 _p(ws), wsb, _st(), work=("flop", 2.0 * N * H * W * Cout * Cin * 9))
            if tgt is None:
                dw = out
        return dx, dw


def conv3x3(x, weight):
    """3x3, stride 1, padding 1, no bias, bf16 (models/ofa/resnet.py conv2 of the stride-1 bottlenecks)."""
    return _Conv3x3.apply(x, weight)


def colsum(x2, alpha=1.0, out=None, accumulate=False):
    rows, Cc = x2.shape
    if out is None:
        out = torch.empty(Cc, dtype=x2.dtype, device=x2.device)
    ws = torch.empty(64 * Cc, dtype=torch.float32, device=x2.device)
    call("ofa_colsum", _p(x2), x2.stride(0), rows, Cc, _p(out), _p(ws), float(alpha), int(accumulate), _dt(x2), _st())
    return out


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm (+ fused GELU prologue / residual epilogue)
# ---------------------------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    """fork=True also returns x itself (second output): the residual branch that leaves before the LayerNorm
    (unify_transformer_layer.py:259-262).  Its gradient comes back into THIS node and is added inside the backward kernel
    (dskip) instead of by autograd's separate accumulation pass over the activations."""

    @staticmethod
    def forward(ctx, x, gamma, beta, resid, gelu_in, eps, fork):
        _need_cuda(x)
        shp = x.shape
        Cc = shp[-1]
        x2 = x.reshape(-1, Cc).contiguous()
        rows = x2.shape[0]
        r2 = resid.reshape(-1, Cc).contiguous() if resid is not None else None
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        call("ofa_layernorm_fwd", _p(x2), _p(gamma), _p(beta), _p(r2), _p(y), _p(mean), _p(rstd), rows, Cc, eps,
             int(gelu_in), _dt(x2), _st(), work=("byte", (2 + (resid is not None)) * rows * Cc * x2.element_size()))
        ctx.save_for_backward(x2, gamma, mean, rstd)
        ctx.beta_param = beta
        ctx.gelu_in, ctx.shp, ctx.has_r, ctx.fork = gelu_in, shp, resid is not None, fork
        ctx.set_materialize_grads(False)
        if fork:
            return y.reshape(shp), x
        return y.reshape(shp)

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x2, gamma, mean, rstd = ctx.saved_tensors
        rows, Cc = x2.shape
        if dy is None:                      # only the skip branch was used
            return dskip, None, None, None, None, None, None
        dy2 = dy.reshape(-1, Cc).contiguous()
        sk2 = dskip.reshape(-1, Cc).contiguous() if dskip is not None else None
        dx = torch.empty_like(x2)
        (tg, ag), (tb, ab) = _acc_target(gamma), _acc_target(ctx.beta_param)
        fused = tg is not None and tb is not None and ag == ab
        acc = fused and ag
        dg = tg if fused else torch.empty_like(gamma)
        db = tb if fused else torch.empty_like(gamma)
        nparts = _lib.load().ofa_layernorm_bwd_nparts(rows)
        ws = torch.empty(2 * nparts * Cc, dtype=torch.float32, device=x2.device)
        call("ofa_layernorm_bwd", _p(dy2), _p(x2), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db), _p(ws), rows,
             Cc, int(ctx.gelu_in), int(acc), _p(sk2), _dt(x2), _st(),
             work=("byte", (3 + (sk2 is not None)) * rows * Cc * x2.element_size()))
        return (dx.reshape(ctx.shp), (None if fused else dg), (None if fused else db), (dy if ctx.has_r else None), None,
                None, None)


def layer_norm(x, gamma, beta, resid=None, gelu_in=False, eps=1e-5, fork=False):
    """LN(f(x)) * gamma + beta (+ resid), f = gelu if gelu_in.  fork=True returns (LN(x), x): use the second value for the
    residual connection so that its gradient is added inside the LayerNorm backward kernel."""
    return _LayerNorm.apply(x, gamma, beta, resid, gelu_in, eps, fork)


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        call("ofa_gelu", _p(x), _p(None), _p(y), x.numel(), 0, _dt(x), _st())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        call("ofa_gelu", _p(x), _p(dy), _p(dx), x.numel(), 1, _dt(x), _st())
        return dx


def gelu(x):
    return _Gelu.apply(x)


# ---------------------------------------------------------------------------------------------------------------------
# embeddings and glue
# ---------------------------------------------------------------------------------------------------------------------
class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, table, addvec, padding_idx):
        _need_cuda(table)
        idx = idx.contiguous()
        rows = idx.numel()
        Cc = table.shape[1]
        out = torch.empty(*idx.shape, Cc, dtype=table.dtype, device=table.device)
        call("ofa_embed_gather", _p(idx), _p(table), _p(addvec), _p(out), Cc, rows, Cc, _dt(table), _st())
        ctx.save_for_backward(idx)
        ctx.table_param = table
        ctx.tshape, ctx.pad, ctx.has_add = table.shape, padding_idx, addvec is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dout = dout.contiguous()
        Cc = ctx.tshape[1]
        dt = da = None
        if ctx.needs_input_grad[1]:
            tgt, accum = _acc_target(ctx.table_param)
            if tgt is not None and not accum:
                tgt.zero_()
            dt = tgt if tgt is not None else torch.zeros(ctx.tshape, dtype=dout.dtype, device=dout.device)
            call("ofa_embed_scatter_add", _p(idx), _p(dout), Cc, _p(dt), idx.numel(), Cc,
                 -1 if ctx.pad is None else ctx.pad, _dt(dout), _st())
            if tgt is not None:
                dt = None
        if ctx.has_add and ctx.needs_input_grad[2]:
            da = colsum(dout.reshape(-1, Cc))
        return None, dt, da, None


def embedding(idx, table, addvec=None, padding_idx=None):
    """table[idx] (+ addvec broadcast over rows); the padding row receives no gradient."""
    return _Embedding.apply(idx, table, addvec, padding_idx)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        call("ofa_add", _p(a), _p(b), _p(out), a.numel(), _dt(a), _st())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


class _MaskRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowmask):
        x = x.contiguous().clone()
        rm = rowmask.contiguous().view(torch.uint8)
        call("ofa_mask_rows", _p(x), _p(rm), rm.numel(), x.shape[-1], _dt(x), _st())
        ctx.save_for_backward(rm)
        return x

    @staticmethod
    def backward(ctx, g):
        (rm,) = ctx.saved_tensors
        g = g.contiguous().clone()
        call("ofa_mask_rows", _p(g), _p(rm), rm.numel(), g.shape[-1], _dt(g), _st())
        return g, None


def mask_rows(x, rowmask):
    """x * (1 - rowmask[..., None])   (unify_transformer.py:892-893)"""
    return _MaskRows.apply(x, rowmask)


class _StemConv7x7(torch.autograd.Function):
    """conv1 of the ResNet stem (7x7, stride 2, padding 3, 3 -> Cout channels, no bias) as patch matrix + tcgen05 GEMM.
    x: channels_last [N, 3, H, W] bf16 (no gradient: it is the image); w: [Cout, 3, 7, 7]."""

    @staticmethod
    def forward(ctx, x, w):
        _need_cuda(x)
        N, Cin, H, W = x.shape
        assert Cin == 3 and tuple(w.shape[1:]) == (3, 7, 7)
        x = x.contiguous(memory_format=torch.channels_last)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        M, Cout = N * OH * OW, w.shape[0]
        col = torch.empty(M, 152, dtype=x.dtype, device=x.device)
        call("ofa_stem_patches", _p(x), _p(col), N, H, W, _st())
        wp = torch.zeros(Cout, 152, dtype=w.dtype, device=w.device)
        wp[:, :147] = w.detach().reshape(Cout, 147)
        y = gemm(col[:, :147], wp[:, :147], M, Cout, 147)
        ctx.save_for_backward(col)
        ctx.w_param, ctx.dims = w, (N, Cout, OH, OW)
        return y.view(N, OH, OW, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        N, Cout, OH, OW = ctx.dims
        M = N * OH * OW
        dw = None
        if ctx.needs_input_grad[1]:
            dy2 = dy.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(M, Cout)
            dwp = gemm(dy2, col[:, :147], Cout, 147, M, a_mn=True, b_mn=True, out_dtype=dy.dtype, ldd=152)
            dw = dwp[:, :147].reshape(ctx.w_param.shape)
        return None, dw


def stem_conv7x7(x, weight):
    return _StemConv7x7.apply(x, weight)


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        N, Cc, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, Cc, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty(N * OH * OW * Cc, dtype=torch.uint8, device=x.device)
        call("ofa_maxpool3x3s2_fwd", _p(x), _p(y), _p(idx), N, H, W, Cc, _st())
        ctx.save_for_backward(idx)
        ctx.shape = (N, Cc, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, Cc, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, Cc, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        call("ofa_maxpool3x3s2_bwd", _p(dy), _p(idx), _p(dx), N, H, W, Cc, _st())
        return dx


class _MaxPool3x3s2Any(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        N, Cc, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, Cc, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty(N * OH * OW * Cc, dtype=torch.uint8, device=x.device)
        call("ofa_maxpool3x3s2_any", _p(x), _p(idx), _p(y), N, H, W, Cc, 0, _dt(x), _st())
        ctx.save_for_backward(idx)
        ctx.shape = (N, Cc, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, Cc, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, Cc, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        call("ofa_maxpool3x3s2_any", _p(dy), _p(idx), _p(dx), N, H, W, Cc, 1, _dt(dy), _st())
        return dx


def max_pool3x3s2(x):
    """nn.MaxPool2d(3, 2, 1) on a channels_last [N, C, H, W] tensor: vectorised bf16 kernels when C % 8 == 0, else (and in the
    fp32 parity mode) the generic kernels of csrc/im2col.cu."""
    if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
        return _MaxPool3x3s2.apply(x)
    return _MaxPool3x3s2Any.apply(x)


class _Im2col(torch.autograd.Function):
    """NHWC patch matrix [N*OH*OW, ceil8(KH*KW*C)] (columns ordered (kh, kw, c), zero padded) and its adjoint (csrc/im2col.cu)."""

    @staticmethod
    def forward(ctx, x, KH, KW, stride, pad):
        _need_cuda(x)
        N, Cc, H, W = x.shape
        x = x.contiguous(memory_format=torch.channels_last)
        OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
        K = KH * KW * Cc
        ld = _ceil8(K)
        col = torch.empty(N * OH * OW, ld, dtype=x.dtype, device=x.device)
        if ld != K:
            col[:, K:].zero_()
        call("ofa_im2col", _p(x), _p(col), N, H, W, Cc, KH, KW, stride, pad, ld, _dt(x), _st(),
             work=("byte", 2.0 * col.numel() * x.element_size()))
        ctx.dims = (N, Cc, H, W, KH, KW, stride, pad, ld)
        return col              # [N*OH*OW, ceil8(K)]: columns beyond K are zero

    @staticmethod
    def backward(ctx, dcol):
        N, Cc, H, W, KH, KW, stride, pad, ld = ctx.dims
        if dcol.stride(-1) != 1 or dcol.stride(0) != ld:
            buf = torch.empty(dcol.shape[0], ld, dtype=dcol.dtype, device=dcol.device)
            buf[:, :dcol.shape[1]].copy_(dcol)
            dcol = buf
        dx = torch.empty((N, Cc, H, W), dtype=dcol.dtype, device=dcol.device, memory_format=torch.channels_last)
        call("ofa_col2im", _p(dcol), _p(dx), N, H, W, Cc, KH, KW, stride, pad, ld, _dt(dcol), _st(),
             work=("byte", float(dcol.numel() + dx.numel()) * dcol.element_size()))
        return dx, None, None, None, None


def conv_im2col(x, weight, stride, pad):
    """k x k convolution (no bias) as patch matrix x weight^T on the tcgen05 GEMM: the two stride-2 3x3 convolutions of the stem
    and every k > 1 convolution of the fp32 parity mode (models/ofa/resnet.py:34-37,107-121,176,214).  x: channels_last
    [N, Cin, H, W]; weight [Cout, Cin, KH, KW] (any memory format; a channels_last weight is used in place)."""
    N, Cin, H, W = x.shape
    Cout, _, KH, KW = weight.shape
    OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    needs_x_grad = x.requires_grad
    col = _Im2col.apply(x if needs_x_grad else x.detach(), KH, KW, stride, pad)
    w2 = weight.permute(0, 2, 3, 1).reshape(Cout, KH * KW * Cin)         # (kh, kw, c) column order; a view for channels_last weights
    K = KH * KW * Cin
    if K % 8:                   # the patch matrix has zero columns up to a multiple of 8 (16-byte TMA rows): pad the weight alike
        w2 = torch.nn.functional.pad(w2, (0, _ceil8(K) - K))
    y = linear(col, w2)
    return y.view(N, OH, OW, Cout).permute(0, 3, 1, 2)


_BN_SCRATCH = {}


def _bn_scratch(device):
    """The self-cleaning per-channel sum scratch of the BatchNorm kernels: zero on entry, left zero by the kernels."""
    key = (device.type, device.index)
    ws = _BN_SCRATCH.get(key)
    if ws is None:
        ws = _BN_SCRATCH[key] = torch.zeros(_lib.load().ofa_batchnorm_workspace_floats(0), dtype=torch.float32, device=device)
    return ws


class _BatchNorm(torch.autograd.Function):
    """x: conv output, logically [N,C,H,W] in channels_last memory format (== [N*H*W, C] row-major)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, relu, training, momentum, eps, groups):
        _need_cuda(x)
        N, Cc, Hh, Ww = x.shape
        assert N % groups == 0
        x = x.contiguous(memory_format=torch.channels_last)
        res = residual.contiguous(memory_format=torch.channels_last) if residual is not None else None
        y = torch.empty_like(x, memory_format=torch.channels_last)
        R = N * Hh * Ww
        stats = torch.empty(groups * 4 * Cc, dtype=torch.float32, device=x.device)
        ws = _bn_scratch(x.device)
        call("ofa_batchnorm_fwd", _p(x), _p(res), _p(y), _p(gamma), _p(beta), _p(running_mean), _p(running_var), R, Cc,
             float(eps), float(momentum), int(training), int(relu), _p(stats), _p(ws), int(groups), _dt(x), _st(),
             work=("byte", (2 + training + (res is not None)) * R * Cc * x.element_size()))
        # with a residual the ReLU mask needs y; without one it is recomputed from x (saves a full read in the backward)
        ctx.save_for_backward(x, y if (relu and res is not None) else None, gamma, stats)
        ctx.beta_param = beta
        ctx.relu, ctx.training, ctx.has_res, ctx.groups = relu, training, residual is not None, groups
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, stats = ctx.saved_tensors
        N, Cc, Hh, Ww = x.shape
        R = N * Hh * Ww
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        dres = torch.empty_like(x, memory_format=torch.channels_last) if ctx.has_res else None
        want_pg = ctx.needs_input_grad[1]
        dg = db = None
        fused = False
        if want_pg:
            (tg, ag), (tb, ab) = _acc_target(gamma), _acc_target(ctx.beta_param)
            fused = tg is not None and tb is not None and ag == ab
            dg = tg if fused else torch.empty_like(gamma)
            db = tb if fused else torch.empty_like(gamma)
        ws = _bn_scratch(x.device)
        call("ofa_batchnorm_bwd", _p(x), _p(dy), _p(y), _p(gamma), _p(stats), _p(dx), _p(dres), _p(dg), _p(db),
             int(fused and ag), R, Cc, int(ctx.training), int(ctx.relu), _p(ws), int(ctx.groups), _dt(x), _st(),
             work=("byte", (4 + 2 * ctx.relu + ctx.has_res) * R * Cc * x.element_size()))
        if fused or not want_pg:
            dg = db = None
        return dx, dg, db, None, None, dres, None, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, residual=None, relu=False, training=True, momentum=0.1,
               eps=1e-5, groups=1):
    """relu?(BN(x) + residual) on a channels_last [N,C,H,W] tensor; training=True uses batch statistics and updates the
    running buffers in place (nn.BatchNorm2d semantics), False uses the running statistics (eval / FrozenBatchNorm2d).
    groups > 1: the batch is `groups` equal consecutive groups normalised independently (several tasks in one stem pass)."""
    return _BatchNorm.apply(x, gamma, beta, running_mean, running_var, residual, relu, training, momentum, eps, groups)


class _DropoutResidual(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, resid, p, row_scale, seed):
        _need_cuda(x)
        x = x.contiguous()
        r = resid.contiguous() if resid is not None else None
        y = torch.empty_like(x)
        Cc = x.shape[-1]
        rps = x.shape[-2] if x.dim() >= 3 else 1
        call("ofa_dropout_residual", _p(x), _p(r), _p(y), x.numel(), Cc, rps, float(p), _p(row_scale), _p(seed), _dt(x), _st())
        ctx.save_for_backward(row_scale, seed)
        ctx.p, ctx.rps, ctx.has_r = p, rps, resid is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        row_scale, seed = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("ofa_dropout_residual", _p(dy), _p(None), _p(dx), dy.numel(), dy.shape[-1], ctx.rps, float(ctx.p),
             _p(row_scale), _p(seed), _dt(dy), _st())
        return dx, (dy if ctx.has_r else None), None, None, None


def dropout_residual(x, resid=None, p=0.0, drop_path=0.0, training=True):
    """resid + drop_path(dropout(x)).  x: [B, L, C]; the drop-path Bernoulli is per sample (dim 0), as in the reference
    (mask shape (1, B, 1) on T x B x C).  Identity (plain add) when not training or both rates are 0."""
    if not training or (p == 0.0 and drop_path == 0.0):
        return add(x, resid) if resid is not None else x
    row_scale = None
    if drop_path > 0.0:
        keep = 1.0 - drop_path
        row_scale = torch.floor(keep + torch.rand(x.shape[0], device=x.device, dtype=torch.float32)) / keep
    seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64)
    return _DropoutResidual.apply(x, resid, p, row_scale, seed)


# ---------------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------------
def _fill_args(q, pq, k, pk, v, o, lse, H, causal, q_pos_off, kpm, head_scale, bias, p_round):
    a = OfaAttnArgs()
    B, T = q.shape[0], q.shape[1]
    S = k.shape[1]
    for name, t in (("q", q), ("pq", pq), ("k", k), ("pk", pk), ("v", v), ("o", o)):
        assert t.stride(2) == 1, name
        setattr(a, name, t.data_ptr())
        setattr(a, "ld" + name, t.stride(1))
        setattr(a, "bs" + name, t.stride(0))
    a.lse = lse.data_ptr()
    a.B, a.H, a.T, a.S = B, H, T, S
    a.causal, a.q_pos_off = int(causal), int(q_pos_off)
    a.kpm = kpm.data_ptr() if kpm is not None else None
    a.head_scale = head_scale.data_ptr() if head_scale is not None else None
    a.p_round_bf16 = int(p_round)
    bz = OfaAttnBias()
    bz.tok_max = 1024
    bz.tok_lut = bias["tok_lut"].data_ptr() if bias.get("tok_lut") is not None else None
    bz.q_text_off, bz.k_text_off = bias.get("q_text_off", 0), bias.get("k_text_off", 0)
    bz.img_lut = bias["img_lut"].data_ptr() if bias.get("img_lut") is not None else None
    bz.n_img_rel = bias["img_lut"].shape[1] if bias.get("img_lut") is not None else 0
    bz.ibs = bias.get("ibs", 42)
    bz.q_pid = bias["q_pid"].data_ptr() if bias.get("q_pid") is not None else None
    bz.k_pid = bias["k_pid"].data_ptr() if bias.get("k_pid") is not None else None
    bz.n_img_q, bz.n_img_k = bias.get("n_img_q", 0), bias.get("n_img_k", 0)
    a.bias = bz
    return a


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, pq, k, pk, v, tok_lut, img_lut, head_scale, cfg):
        _need_cuda(q)
        B, T, D = q.shape
        H = cfg["H"]
        o = torch.empty(B, T, D, dtype=q.dtype, device=q.device)
        lse = torch.empty(B, H, T, dtype=torch.float32, device=q.device)
        hs = head_scale.float().contiguous() if head_scale is not None else None
        bias = dict(cfg["bias"])
        bias["tok_lut"], bias["img_lut"] = tok_lut, img_lut
        kpm = cfg.get("kpm")
        a = _fill_args(q, pq, k, pk, v, o, lse, H, cfg["causal"], cfg.get("q_pos_off", 0), kpm, hs, bias,
                       q.dtype == torch.bfloat16)
        if q.dtype == torch.bfloat16 and cfg.get("use_tc", True):
            call("ofa_attn_fwd_tc", C.byref(a), _st(), work=("flop", 2.0 * B * H * T * k.shape[1] * 192))
        else:
            call("ofa_attn_fwd_simt", C.byref(a), _dt(q), _st(), work=("flop", 2.0 * B * H * T * k.shape[1] * 192))
        ctx.save_for_backward(q, pq, k, pk, v, o, lse, tok_lut, img_lut, hs, head_scale)
        ctx.cfg, ctx.bias = cfg, bias
        return o

    @staticmethod
    def backward(ctx, do):
        q, pq, k, pk, v, o, lse, tok_lut, img_lut, hs, head_scale = ctx.saved_tensors
        cfg = ctx.cfg
        B, T, D = q.shape
        S, H = k.shape[1], cfg["H"]
        do = do.contiguous()
        a = _fill_args(q, pq, k, pk, v, o, lse, H, cfg["causal"], cfg.get("q_pos_off", 0), cfg.get("kpm"), hs,
                       ctx.bias, q.dtype == torch.bfloat16)
        g = OfaAttnGrads()
        # pos_q / pos_k are shared by every layer of a stack: with a `pos_sink` (one dict per forward pass of the stack, see
        # ofa.py) the layers' backward kernels sum their contributions into ONE pair of buffers (first call writes, later calls
        # add) and only the last call hands them to autograd -- instead of autograd adding 2 x (layers - 1) full tensors
        sink = cfg.get("pos_sink") if (q.dtype == torch.bfloat16 and cfg.get("use_tc", True)) else None
        last_of_stack = True
        if sink is not None:
            if sink.get("dpq") is None:
                sink["dpq"] = torch.empty(B, T, D, dtype=q.dtype, device=q.device)
                sink["dpk"] = torch.empty(B, S, D, dtype=q.dtype, device=q.device)
                sink["seen"] = 0
            else:
                g.acc_pos = 1
            dpq, dpk = sink["dpq"], sink["dpk"]
            sink["seen"] += 1
            last_of_stack = sink["seen"] == sink["n"]
        else:
            dpq = torch.empty(B, T, D, dtype=q.dtype, device=q.device)
            dpk = torch.empty(B, S, D, dtype=q.dtype, device=q.device)
        dq_scale = float(cfg.get("dq_scale", 1.0))
        if cfg.get("fused_kv"):
            # cross-attention fed by the fused k|v projection: dk and dv side by side
            dq = torch.empty(B, T, D, dtype=q.dtype, device=q.device)
            dkv = torch.empty(B, S, 2 * D, dtype=q.dtype, device=q.device)
            dk, dv = dkv[..., :D], dkv[..., D:]
        elif cfg.get("fused_qkv") and T == S:
            # self-attention fed by the fused q|k|v projection: the three gradients land side by side in one [B, L, 3D]
            # buffer, which is the A operand of that projection's single dgrad / wgrad GEMM (no concatenation copy)
            dqkv = torch.empty(B, T, 3 * D, dtype=q.dtype, device=q.device)
            dq, dk, dv = dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:]
        else:
            dq = torch.empty(B, T, D, dtype=q.dtype, device=q.device)
            dk, dv = (torch.empty(B, S, D, dtype=q.dtype, device=q.device) for _ in range(2))
        g.dout = do.data_ptr()
        for name, t in (("dq", dq), ("dpq", dpq), ("dk", dk), ("dpk", dpk), ("dv", dv)):
            setattr(g, name, t.data_ptr())
            setattr(g, "ld" + name, t.stride(1))
            setattr(g, "bs" + name, t.stride(0))
        dtok = torch.zeros_like(tok_lut) if tok_lut is not None else None
        dimg = torch.zeros_like(img_lut) if img_lut is not None else None
        g.dtok_lut = dtok.data_ptr() if dtok is not None else None
        g.dimg_lut = dimg.data_ptr() if dimg is not None else None
        delta = torch.empty(B, H, T, dtype=torch.float32, device=q.device)
        g.delta = delta.data_ptr()
        if q.dtype == torch.bfloat16 and cfg.get("use_tc", True):
            dq_acc = torch.empty(B * T * H * 128, dtype=torch.float32, device=q.device)
            g.dq_scale = dq_scale
            call("ofa_attn_bwd_tc", C.byref(a), C.byref(g), _p(dq_acc), _st(), work=("flop", 2.0 * B * H * T * S * 512))
        else:
            P = torch.empty(B, H, T, S, dtype=torch.float32, device=q.device)
            dS = torch.empty(B, H, T, S, dtype=torch.float32, device=q.device)
            g.P, g.dS = P.data_ptr(), dS.data_ptr()
            call("ofa_attn_bwd_simt", C.byref(a), C.byref(g), _dt(q), _st(), work=("flop", 2.0 * B * H * T * S * 512))
            if dq_scale != 1.0:
                dq.mul_(dq_scale)
        dhs = None
        if head_scale is not None:
            dhs = (delta.sum(dim=(0, 2)) / hs).to(head_scale.dtype)
        if not last_of_stack:
            return dq, None, dk, None, dv, dtok, dimg, dhs, None
        return dq, dpq, dk, dpk, dv, dtok, dimg, dhs, None


def attention(q, pq, k, pk, v, tok_lut, img_lut, head_scale, cfg):
    """softmax(q k^T + pq pk^T + rel-pos LUT bias + masks) v * c_attn.   q/pq/k/pk/v: [B, L, H*64] (unit inner
    stride); tok_lut [H, 2047] / img_lut [H, n_rel] fp32 (differentiable); cfg: H, causal, kpm, q_pos_off, bias{...}."""
    return _Attention.apply(q, pq, k, pk, v, tok_lut, img_lut, head_scale, cfg)


def attention_decode(q, pq, k, pk, v, S, H, G=1, kv_row=None, pk_row=None, kpm=None, head_scale=None, tok_lut=None,
                     q_pos=0, bias_in=None, score_out=None, page=None):
    """Single-token attention over a KV cache (inference only; csrc/decode.cu).  q, pq: [R, 1, H*64] (pre-scaled);
    k, v: caches [rows, cap, H*64] of which the first S positions are valid; pk likewise; G consecutive query rows share
    cache row kv_row[group] (int32).  pq / pk None: no absolute-position term.  bias_in [R, H, ld] fp32 is added to the
    scores; with score_out [R, H, ld] the launch only writes its raw scores (the shared cross-attention position term,
    computed once per step).  page = (table int32 [rows, max_pages], page_len, page_stride): k / v are page pools.
    Returns [R, 1, H*64] (None in score_out mode)."""
    _need_cuda(q)
    R, D = q.shape[0], q.shape[-1]
    q2 = q.reshape(R, D)
    a = OfaDecodeArgs()
    a.q, a.ldq = q2.data_ptr(), q2.stride(0)
    if pq is not None:
        pq2 = pq.reshape(R, D)
        a.pq, a.ldpq = pq2.data_ptr(), pq2.stride(0)
        a.pk, a.ldpk, a.bspk = pk.data_ptr(), pk.stride(1), pk.stride(0)
    a.k = k.data_ptr()
    if page is not None:
        table, page_len, page_stride = page
        a.page_table, a.page_len, a.max_pages, a.page_stride = table.data_ptr(), int(page_len), table.shape[1], int(page_stride)
        a.ldk = a.ldv = D
        a.bsk = a.bsv = 0
    else:
        a.ldk, a.bsk = k.stride(1), k.stride(0)
    o = None
    if score_out is None:
        a.v = v.data_ptr()
        if page is None:
            a.ldv, a.bsv = v.stride(1), v.stride(0)
        o = torch.empty(R, D, dtype=q.dtype, device=q.device)
        a.o, a.ldo = o.data_ptr(), D
    else:
        a.score_out, a.bias_ld = score_out.data_ptr(), score_out.stride(1)
    if bias_in is not None:
        a.bias_in, a.bias_ld = bias_in.data_ptr(), bias_in.stride(1)
    a.kv_row = kv_row.data_ptr() if kv_row is not None else None
    a.pk_row = pk_row.data_ptr() if pk_row is not None else None
    a.kpm = kpm.data_ptr() if kpm is not None else None
    a.kpm_stride = kpm.stride(0) if kpm is not None else 0
    hs = head_scale.float().contiguous() if head_scale is not None else None
    a.head_scale = hs.data_ptr() if hs is not None else None
    a.tok_lut = tok_lut.data_ptr() if tok_lut is not None else None
    a.tok_max, a.q_pos = 1024, int(q_pos)
    a.R, a.G, a.H, a.S = R, int(G), int(H), int(S)
    nb = (2 if score_out is not None or pq is None else 3) - (1 if score_out is not None else 0)
    call("ofa_attn_decode", C.byref(a), _dt(q), _st(),
         work=("byte", (R // max(G, 1)) * S * D * float(nb) * q.element_size()))
    return o.view(R, 1, D) if o is not None else None


def page_reorder(pool, table_src, table_dst, order, rows, page, off, parity, page_len):
    """Beam reorder of the paged self-attention cache: table entries of the full pages are copied, the partial last page is
    copied on write into the row's own slot (csrc/decode.cu)."""
    n_slots, planes, pl, D = pool.shape
    max_pages = table_src.shape[1]
    call("ofa_page_reorder", _p(pool), _p(table_src), _p(table_dst), _p(order), rows, max_pages, int(page), int(off), int(parity),
         int(page_len), D, planes, _dt(pool), _st())


def page_write(pool, table, k, v, layer, page, off, page_len):
    """Append one layer's new K / V token of every row at position `off` of page `page`."""
    n_slots, planes, pl, D = pool.shape
    rows, max_pages = table.shape
    k2, v2 = k.reshape(rows, D), v.reshape(rows, D)
    call("ofa_page_write", _p(pool), _p(table), _p(k2), _p(v2), k2.stride(0), v2.stride(0), rows, max_pages, int(page), int(off),
         int(page_len), D, planes, 2 * layer, _dt(pool), _st())


def beam_topk(logits, beam, K, temperature=1.0, prev_scores=None, step0=False, eos=2, pad=1, unk=3, unk_penalty=0.0,
              block_eos=False, force_eos=False, eos_one=False, crange=None, range_post=False, trie=None, node=None,
              trie_post=False, tokens=None, step=0, ngram=0, ws=None, prefix_tok=None, prefix_fill=None):
    """Fused tail of a beam-search step (csrc/beam.cu): logits [R, V] -> (cand_scores [R/beam, K] fp32, cand_index int64).
    trie = (ptr, tok, child) int32 CSR tensors; node int32 [R] (-2: row not constrained); tokens int64 [R, L] for n-gram
    blocking; prefix_tok int64 [R] (pad: none) + prefix_fill fp32 [1] or None (-inf): forced prefix tokens."""
    from ._lib import OfaBeamArgs
    _need_cuda(logits)
    R, V = logits.shape
    assert logits.stride(1) == 1
    a = OfaBeamArgs()
    a.logits, a.ld, a.dtype = logits.data_ptr(), logits.stride(0), _dt(logits)
    a.R, a.beam, a.V, a.K = R, int(beam), V, int(K)
    a.temperature = float(temperature)
    a.prev_scores = prev_scores.data_ptr() if prev_scores is not None else None
    a.step0 = int(step0)
    a.eos, a.pad, a.unk, a.unk_penalty = int(eos), int(pad), int(unk), float(unk_penalty)
    a.block_eos, a.force_eos, a.eos_one = int(block_eos), int(force_eos), int(eos_one)
    a.range_lo, a.range_hi = (int(crange[0]), int(crange[1])) if crange is not None else (-1, -1)
    a.range_post = int(range_post)
    if trie is not None:
        a.trie_ptr, a.trie_tok, a.node, a.trie_post = trie[0].data_ptr(), trie[1].data_ptr(), node.data_ptr(), int(trie_post)
    if ngram > 0:
        a.tokens, a.ldtok, a.step, a.ngram = tokens.data_ptr(), tokens.stride(0), int(step), int(ngram)
    if prefix_tok is not None:
        assert prefix_tok.dtype == torch.long and prefix_tok.numel() == R and prefix_tok.is_contiguous()
        a.prefix_tok = prefix_tok.data_ptr()
        if prefix_fill is not None:
            assert prefix_fill.dtype == torch.float32 and prefix_fill.numel() == 1
            a.prefix_fill = prefix_fill.data_ptr()
    kw = _lib.load().ofa_beam_topk_width(int(K))
    if ws is None or ws[0].numel() < R * kw:
        ws = (torch.empty(R * kw, dtype=torch.float32, device=logits.device),
              torch.empty(R * kw, dtype=torch.int32, device=logits.device))
    cs = torch.empty(R // beam, K, dtype=torch.float32, device=logits.device)
    ci = torch.empty(R // beam, K, dtype=torch.int64, device=logits.device)
    a.row_val, a.row_idx, a.cand_scores, a.cand_index = ws[0].data_ptr(), ws[1].data_ptr(), cs.data_ptr(), ci.data_ptr()
    call("ofa_beam_topk", C.byref(a), _st(), work=("byte", 2.0 * R * V * logits.element_size()))
    return cs, ci, ws


def beam_advance(cand_scores, cand_index, ignore_in, tok_in, sc_in, tok_out, sc_out, ignore_out, active_bbsz, eos_n, beam, V,
                 eos, step):
    """Beam bookkeeping of one step without finalisation (csrc/beam.cu beam_advance_kernel); all buffers preallocated."""
    _need_cuda(cand_scores)
    bsz, C2 = cand_scores.shape
    assert ignore_in.dtype == torch.bool and ignore_out.dtype == torch.bool and eos_n.dtype == torch.int32
    assert tok_in.stride(1) == 1 and tok_out.stride(0) == tok_in.stride(0) and sc_out.stride(0) == sc_in.stride(0)
    call("ofa_beam_advance", _p(cand_scores), _p(cand_index), C2, _p(ignore_in), _p(tok_in), tok_in.stride(0), _p(sc_in),
         sc_in.stride(0), _p(tok_out), _p(sc_out), _p(ignore_out), _p(active_bbsz), _p(eos_n), bsz, int(beam), int(V), int(eos),
         int(step), _st())


def trie_advance(trie, node_in, parent, tok):
    """node_out[r] = child of node_in[parent[r]] along tok[r] (-1 when the prefix leaves the trie)."""
    out = torch.empty_like(node_in)
    call("ofa_trie_advance", _p(trie[0]), _p(trie[1]), _p(trie[2]), _p(node_in), _p(parent), _p(tok), tok.stride(0), _p(out),
         node_in.numel(), _st())
    return out


def trie_score(logits, seg_off, node, target, trie, pad):
    """All-candidate scores (csrc/beam.cu trie_score_kernel): logits [P, V] of the counted positions, seg_off int32 [rows + 1],
    node int32 [P], target int64 [P]; trie = (ptr, tok, child) CSR or None when every node is -1 / -2.  Returns fp32 [rows]."""
    _need_cuda(logits)
    P, V = logits.shape
    assert logits.stride(1) == 1 and node.numel() == P and target.numel() == P
    rows = seg_off.numel() - 1
    out = torch.empty(rows, dtype=torch.float32, device=logits.device)
    call("ofa_trie_score", _p(logits), logits.stride(0), _dt(logits), V, _p(seg_off), _p(node), _p(target),
         _p(trie[0]) if trie is not None else None, _p(trie[1]) if trie is not None else None, int(pad), _p(out), rows, _st())
    return out


def cache_gather(src, dst, order, rows, L):
    """dst[p, r, :L] = src[p, order[r], :L] for caches [planes, cap_rows, cap_len, D] (beam reorder, valid prefix only)."""
    planes, cap_rows, cap_len, D = src.shape
    call("ofa_cache_gather", _p(src), _p(dst), _p(order), int(rows), int(L), D, cap_len * D, cap_rows * cap_len * D, planes,
         _dt(src), _st())


# ---------------------------------------------------------------------------------------------------------------------
# fused label-smoothed cross-entropy (+ R-Drop)
# ---------------------------------------------------------------------------------------------------------------------
class _LsCe(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, cmask, conf, eps, pad_idx, crange, rdrop, reg_alpha, drop_ratio):
        """logits [B,T,V] (row stride may be padded) are CONSUMED: the kernel overwrites the buffer with
        d loss / d logits behind autograd's back (nothing upstream saves the logits; see decoder.output_layer)."""
        _need_cuda(logits)
        B, T, V = logits.shape
        assert logits.stride(2) == 1 and logits.stride(0) == T * logits.stride(1)
        R = B * T
        tgt = target.contiguous()
        loss_rows = torch.empty(R, dtype=torch.float32, device=logits.device)
        nll_rows = torch.empty(R, dtype=torch.float32, device=logits.device)
        kl_rows = torch.zeros(R // 2 if rdrop else 1, dtype=torch.float32, device=logits.device)
        cm = cmask.contiguous().view(torch.uint8) if cmask is not None else None
        cf = conf.float().contiguous() if conf is not None else None
        cs, ce = crange if crange is not None else (-1, -1)
        call("ofa_ls_ce_fwd_bwd", _p(logits), logits.stride(1), _p(tgt), _p(cm), _p(cf), T, R, V, pad_idx, eps, cs, ce,
             int(rdrop), reg_alpha, _p(loss_rows), _p(nll_rows), _p(kl_rows), _p(None), _dt(logits), _st(),
             work=("byte", 2 * R * V * logits.element_size()))
        ctx.dlogits = logits.detach()
        ctx.row_keep = None
        if drop_ratio > 0:
            # drop-worst (criterions/label_smoothed_cross_entropy.py:100-111): keep the int(n * (1 - ratio)) non-pad rows with
            # the smallest loss (of the first R-Drop half; the second half keeps the same rows); dropped rows leave the
            # loss, the KL term and -- through row_keep in the backward -- the gradient
            valid = tgt.reshape(-1).ne(pad_idx)
            half = R // 2 if rdrop else R
            v1 = valid[:half]
            k = int(int(v1.sum()) * (1 - drop_ratio))       # host sync, as the reference's boolean indexing + topk
            idx = torch.topk(loss_rows[:half].masked_fill(~v1, float("inf")), k=k, largest=False).indices
            keep1 = torch.zeros(half, dtype=torch.bool, device=logits.device)
            keep1[idx] = True
            keep = torch.cat([keep1, keep1]) if rdrop else keep1
            loss_rows = loss_rows * keep
            nll_rows = nll_rows * keep
            if rdrop:
                kl_rows = kl_rows * keep1
            ctx.row_keep = keep.view(torch.uint8)
        ctx.mark_non_differentiable(nll_rows)
        loss = loss_rows.sum() + (reg_alpha * kl_rows.sum() if rdrop else 0.0)
        return loss, nll_rows

    @staticmethod
    def backward(ctx, gloss, gnll):
        dlogits = ctx.dlogits
        B, T, V = dlogits.shape
        scale = gloss.float().reshape(1).contiguous()
        call("ofa_scale_rows", _p(dlogits), dlogits.stride(1), B * T, V, _p(scale), _p(ctx.row_keep), 0, _dt(dlogits), _st())
        return dlogits, None, None, None, None, None, None, None, None, None


class _LsCeRows(torch.autograd.Function):
    """Per-row variant (no R-Drop / drop-worst): returns the differentiable vector of row losses, so that rows of several
    tasks decoded in one batch can be summed and normalised per task; the backward scales every row of the in-place
    d(logits) by its own upstream gradient."""

    @staticmethod
    def forward(ctx, logits, target, cmask, conf, eps, pad_idx, crange, expected_grad):
        _need_cuda(logits)
        B, T, V = logits.shape
        assert logits.stride(2) == 1 and logits.stride(0) == T * logits.stride(1)
        R = B * T
        tgt = target.contiguous()
        # expected_grad [R] fp32: the caller's promise of d(total loss) / d(loss_rows); the gradient rows are written already
        # multiplied by it, and the backward only rescales the rows whose actual upstream gradient differs (none, normally)
        gsc = expected_grad.float().contiguous() if expected_grad is not None else None
        ctx.gsc = gsc
        loss_rows = torch.empty(R, dtype=torch.float32, device=logits.device)
        nll_rows = torch.empty(R, dtype=torch.float32, device=logits.device)
        kl_rows = torch.zeros(1, dtype=torch.float32, device=logits.device)
        cm = cmask.contiguous().view(torch.uint8) if cmask is not None else None
        cf = conf.float().contiguous() if conf is not None else None
        cs, ce = crange if crange is not None else (-1, -1)
        call("ofa_ls_ce_fwd_bwd", _p(logits), logits.stride(1), _p(tgt), _p(cm), _p(cf), T, R, V, pad_idx, eps, cs, ce,
             0, 1.0, _p(loss_rows), _p(nll_rows), _p(kl_rows), _p(gsc), _dt(logits), _st(),
             work=("byte", 2 * R * V * logits.element_size()))
        ctx.dlogits = logits.detach()
        ctx.mark_non_differentiable(nll_rows)
        return loss_rows, nll_rows

    @staticmethod
    def backward(ctx, g_rows, gnll):
        dlogits = ctx.dlogits
        B, T, V = dlogits.shape
        scale = g_rows.float().contiguous()
        if ctx.gsc is not None:     # exactly 1 where the promise held (same fp32 value), 0 / 0 -> rows that carry no gradient
            scale = torch.where(ctx.gsc != 0, scale / ctx.gsc, torch.ones_like(scale))
        call("ofa_scale_rows", _p(dlogits), dlogits.stride(1), B * T, V, _p(scale), _p(None), 1, _dt(dlogits), _st())
        return dlogits, None, None, None, None, None, None, None


def ls_cross_entropy_rows(logits, target, eps, pad_idx, cmask=None, conf=None, crange=None, expected_grad=None):
    """-> (loss_rows [B*T], nll_rows [B*T]); pad rows are 0.  `logits` is overwritten with its own gradient.
    expected_grad [B*T]: what the caller will multiply every row's loss by (1 / sample_size of its task); when the backward
    then receives exactly that, the logits-sized gradient needs no second pass."""
    return _LsCeRows.apply(logits, target, cmask, conf, eps, pad_idx, crange, expected_grad)


def ls_cross_entropy(logits, target, eps, pad_idx, cmask=None, conf=None, crange=None, rdrop=False, reg_alpha=1.0,
                     drop_worst_ratio=0.0):
    """Sum over non-pad rows of the label-smoothed NLL (+ reg_alpha * symmetric KL between the two R-Drop halves),
    optionally over the (1 - drop_worst_ratio) fraction of rows with the smallest loss.  Returns (loss, nll_rows; dropped
    rows are 0).  `logits` is overwritten with its own gradient (one read + one write of M x V)."""
    return _LsCe.apply(logits, target, cmask, conf, eps, pad_idx, crange, rdrop, reg_alpha, float(drop_worst_ratio))
