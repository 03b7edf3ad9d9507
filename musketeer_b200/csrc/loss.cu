// Fused label-smoothed cross-entropy (+ R-Drop symmetric KL) over the ~59k vocabulary: one CTA per target row (or R-Drop
// row pair); online-softmax statistics in one sweep, then the gradient is written IN PLACE over the logits, so the
// M x V matrix crosses HBM once in each direction (the second/third sweeps of a row hit L2: a row pair is <= 476 KB).
//   restates criterions/label_smoothed_cross_entropy.py:81-126 (label_smoothed_nll_loss), :74-78 (kl_loss),
//   :228-260 (constraint fill, fp32 log-softmax, conf scaling, pad-row filtering).
// Masked vocabulary entries (constraint mask false / outside the constraint range) contribute 0 everywhere, which is
// the torch-1.8.1 kl_div behaviour the reference pins.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 512;

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return (float)*p; }

struct RowCtx {
  const unsigned char* mask;  // row of the constraint mask or null
  int cs, ce;                 // constraint range (cs < 0: none)
  __device__ __forceinline__ bool allowed(int v) const {
    if (mask) return mask[v] != 0;
    if (cs >= 0) return v < 4 || (v >= cs && v < ce);
    return true;
  }
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t = fmaxf(t, sh[i]);
  return t;
}

// statistics of one row: lse over the allowed set, sum of x over the allowed set, count of allowed entries
template <typename T>
__device__ void row_stats(const T* x, int V, const RowCtx& rc, float* sh, float& lse, float& sumx, float& cnt) {
  float m = -CUDART_INF_F, s = 0.f, sx = 0.f, n = 0.f;
  for (int v = threadIdx.x; v < V; v += kThreads) {
    if (!rc.allowed(v)) continue;
    const float xv = ldf(x + v);
    if (xv > m) { s = s * expf(m - xv) + 1.f; m = xv; }
    else s += expf(xv - m);
    sx += xv;
    n += 1.f;
  }
  const float M = block_max(m, sh);
  s = (m == -CUDART_INF_F) ? 0.f : s * expf(m - M);
  s = block_sum(s, sh);
  sumx = block_sum(sx, sh);
  cnt = block_sum(n, sh);
  lse = M + logf(s);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) ls_ce_kernel(T* __restrict__ logits, long long ld, const long long* __restrict__ target,
                                                         const unsigned char* __restrict__ cmask, const float* __restrict__ conf,
                                                         int rows_per_sample, int R, int V, long long pad_idx, float eps,
                                                         int cs, int ce, int rdrop, float reg_alpha,
                                                         float* __restrict__ loss_out, float* __restrict__ nll_out,
                                                         float* __restrict__ kl_out, const float* __restrict__ grad_row_scale) {
  pdl_sync();
  __shared__ float sh[kThreads / 32];
  const int half = rdrop ? R / 2 : R;
  const int r0 = blockIdx.x;  // < half
  const int nrow = rdrop ? 2 : 1;
  T* xr[2];
  RowCtx rc[2];
  long long tg[2];
  float c[2], lse[2], sumx[2], cnt[2];
  for (int k = 0; k < nrow; ++k) {
    const int r = r0 + k * half;
    xr[k] = logits + (size_t)r * ld;
    rc[k].mask = cmask ? cmask + (size_t)r * V : nullptr;
    rc[k].cs = cs; rc[k].ce = ce;
    tg[k] = target[r];
    c[k] = conf ? conf[r / rows_per_sample] : 1.f;
  }
  const bool ignored = tg[0] == pad_idx;  // R-Drop halves carry identical targets (construct_rdrop_sample)
  if (ignored) {
    for (int k = 0; k < nrow; ++k) {
      for (int v = threadIdx.x; v < V; v += kThreads) xr[k][v] = (T)0.f;
      if (threadIdx.x == 0) { loss_out[r0 + k * half] = 0.f; nll_out[r0 + k * half] = 0.f; }
    }
    if (rdrop && threadIdx.x == 0 && kl_out) kl_out[r0] = 0.f;   // kl_out has R/2 entries only under R-Drop
    return;
  }
  for (int k = 0; k < nrow; ++k) row_stats(xr[k], V, rc[k], sh, lse[k], sumx[k], cnt[k]);

  // R-Drop row sums:  Zp = sum e^{c p}, Ap = sum e^{c p}(q - p)  and the q-side twins (p, q = log-probs of the pair)
  float Z[2] = {1.f, 1.f}, A[2] = {0.f, 0.f};
  if (rdrop) {
    float z0 = 0.f, z1 = 0.f, a0 = 0.f, a1 = 0.f;
    for (int v = threadIdx.x; v < V; v += kThreads) {
      if (!rc[0].allowed(v)) continue;
      const float p = ldf(xr[0] + v) - lse[0], q = ldf(xr[1] + v) - lse[1];
      const float ep = expf(c[0] * p), eq = expf(c[1] * q);
      z0 += ep; z1 += eq;
      a0 += ep * (q - p);
      a1 += eq * (p - q);
    }
    Z[0] = block_sum(z0, sh); Z[1] = block_sum(z1, sh);
    A[0] = block_sum(a0, sh); A[1] = block_sum(a1, sh);
  }
  float eps_i[2];
  for (int k = 0; k < nrow; ++k) {
    const bool constrained = rc[k].mask || cs >= 0;
    eps_i[k] = constrained ? eps / (cnt[k] - 1.f + 1e-6f) : eps / (float)(V - 1);
    if (threadIdx.x == 0) {
      const float lp_y = ldf(xr[k] + tg[k]) - lse[k];
      const float nll = -c[k] * lp_y;
      const float smooth = -c[k] * (sumx[k] - cnt[k] * lse[k]);
      loss_out[r0 + k * half] = (1.f - eps - eps_i[k]) * nll + eps_i[k] * smooth;
      nll_out[r0 + k * half] = nll;
    }
  }
  if (rdrop && threadIdx.x == 0 && kl_out) {
    // S = 1/2 sum (e^{c q} - e^{c p}) c (q - p)   (c identical in both halves)
    kl_out[r0] = -0.5f * (c[0] * A[0] + c[1] * A[1]);
  }
  __syncthreads();  // all reads of x[target] above happen before the in-place overwrite below
  // gradient sweep (in place).  d/dx_u = c(1-eps-eps_i)(P_u - [u=y]) + c eps_i (n P_u - 1) + alpha (g_u - P_u sum_v g_v)
  float sg[2], gs[2] = {1.f, 1.f};
  for (int k = 0; k < nrow; ++k) sg[k] = 0.5f * (-c[k] * c[k] * A[k] - c[k] * (Z[1 - k] - Z[k]));
  if (grad_row_scale)       // the caller's promise of the upstream gradient of this row's loss (the loss value stays unscaled)
    for (int k = 0; k < nrow; ++k) gs[k] = grad_row_scale[r0 + k * half];
  for (int v = threadIdx.x; v < V; v += kThreads) {
    const bool ok = rc[0].allowed(v);
    float x0 = 0.f, x1 = 0.f;
    if (ok) { x0 = ldf(xr[0] + v); if (rdrop) x1 = ldf(xr[1] + v); }
    float g0 = 0.f, g1 = 0.f;
    if (ok) {
      const float p = x0 - lse[0];
      const float P = expf(p);
      g0 = c[0] * (1.f - eps - eps_i[0]) * (P - (v == tg[0] ? 1.f : 0.f)) + c[0] * eps_i[0] * (cnt[0] * P - 1.f);
      if (rdrop) {
        const float q = x1 - lse[1];
        const float Q = expf(q);
        g1 = c[1] * (1.f - eps - eps_i[1]) * (Q - (v == tg[1] ? 1.f : 0.f)) + c[1] * eps_i[1] * (cnt[1] * Q - 1.f);
        const float ep = expf(c[0] * p), eq = expf(c[1] * q);
        const float gp = 0.5f * (-c[0] * c[0] * ep * (q - p) - c[0] * (eq - ep));
        const float gq = 0.5f * (-c[1] * c[1] * eq * (p - q) - c[1] * (ep - eq));
        g0 += reg_alpha * (gp - P * sg[0]);
        g1 += reg_alpha * (gq - Q * sg[1]);
      }
    }
    xr[0][v] = (T)(g0 * gs[0]);
    if (rdrop) xr[1][v] = (T)(g1 * gs[1]);
  }
}

template <typename T>
__global__ void scale_rows_kernel(T* __restrict__ x, long long ld, int V, const float* __restrict__ scale,
                                  const unsigned char* __restrict__ row_keep, int per_row) {
  pdl_sync();
  const int r = blockIdx.x;
  const float s = (row_keep && !row_keep[r]) ? 0.f : scale[per_row ? r : 0];
  if (s == 1.0f) return;       // exact no-op: the row is neither read nor written
  T* xr = x + (size_t)r * ld;
  int v0 = 0;
  if (sizeof(T) == 2 && (reinterpret_cast<uintptr_t>(xr) & 15) == 0) {       // 16-byte vectors over the aligned body
    const int nv = V / 8;
    uint4* x4 = reinterpret_cast<uint4*>(xr);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      uint4 u = x4[i];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); h[k] = __floats2bfloat162_rn(f.x * s, f.y * s); }
      x4[i] = u;
    }
    v0 = nv * 8;
  }
  for (int v = v0 + threadIdx.x; v < V; v += blockDim.x) xr[v] = (T)((float)xr[v] * s);
}

}  // namespace

extern "C" int ofa_ls_ce_fwd_bwd(void* logits, long long ld, const long long* target, const unsigned char* cmask,
                                 const float* conf, int rows_per_sample, int R, int V, long long pad_idx, float eps,
                                 int cs, int ce, int rdrop, float reg_alpha, float* loss_rows, float* nll_rows,
                                 float* kl_rows, const float* grad_row_scale, int dtype, void* stream) {
  OFA_CHECK(R > 0 && V > 1, "ofa_ls_ce_fwd_bwd: R=%d V=%d", R, V);
  OFA_CHECK(!rdrop || (R % 2 == 0 && kl_rows), "ofa_ls_ce_fwd_bwd: R-Drop needs an even row count and kl_rows");
  OFA_CHECK(!(cmask && cs >= 0), "ofa_ls_ce_fwd_bwd: constraint mask and constraint range are exclusive");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = rdrop ? R / 2 : R;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(ls_ce_kernel<__nv_bfloat16>, grid, kThreads, 0, st, (__nv_bfloat16*)logits, ld, target, cmask, conf, rows_per_sample, R, V, pad_idx, eps, cs, ce, rdrop, reg_alpha, loss_rows, nll_rows, kl_rows, grad_row_scale));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(ls_ce_kernel<float>, grid, kThreads, 0, st, (float*)logits, ld, target, cmask, conf, rows_per_sample, R, V, pad_idx, eps, cs, ce, rdrop, reg_alpha, loss_rows, nll_rows, kl_rows, grad_row_scale));
  else
    return ofa_set_error("ofa_ls_ce_fwd_bwd: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("ls_ce_kernel");
  return 0;
}

// dlogits[r, :] *= scale (device scalar, or scale[r] with scale_per_row: the upstream gradient of every row's loss -- rows of
// several tasks normalised by their own sample sizes), rows with row_keep==0 zeroed (drop-worst, :100-111)
extern "C" int ofa_scale_rows(void* x, long long ld, int R, int V, const float* scale, const unsigned char* row_keep,
                              int scale_per_row, int dtype, void* stream) {
  OFA_CHECK(R > 0 && V > 0, "ofa_scale_rows: R=%d V=%d", R, V);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16) OFA_CUDA(ofa_launch_pdl(scale_rows_kernel<__nv_bfloat16>, R, 512, 0, st, (__nv_bfloat16*)x, ld, V, scale, row_keep, scale_per_row));
  else if (dtype == OFA_F32) OFA_CUDA(ofa_launch_pdl(scale_rows_kernel<float>, R, 512, 0, st, (float*)x, ld, V, scale, row_keep, scale_per_row));
  else return ofa_set_error("ofa_scale_rows: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("scale_rows_kernel");
  return 0;
}
