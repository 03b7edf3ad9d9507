// HBM-bound row-wise kernels of the OFA hot path: LayerNorm fwd/bwd (warp-shuffle, 128-bit vectorised, optional
// fused GELU prologue and residual epilogue), embedding gather / scatter-add, column sums (bias grads), the bf16x3
// operand split used by the fp32 parity mode, and small elementwise helpers.
//   reference call sites: fairseq LayerNorm uses in models/ofa/unify_transformer_layer.py:259-283,466-560 and
//   unify_transformer.py:731-747,898-904,951,1300,1486-1493,1567; embeddings unify_transformer.py:725-744,885,1450,1475.
#include "common.cuh"

namespace {

template <typename T>
struct Vec8 {};  // 8 elements per thread-vector
template <>
struct Vec8<float> {
  struct Raw { float4 a, b; };
  __device__ static Raw ldraw(const float* p) { Raw r; r.a = reinterpret_cast<const float4*>(p)[0]; r.b = reinterpret_cast<const float4*>(p)[1]; return r; }
  __device__ static void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  __device__ static void unpack2(const Raw& r, float2 (&v)[4]) {
    v[0] = make_float2(r.a.x, r.a.y); v[1] = make_float2(r.a.z, r.a.w);
    v[2] = make_float2(r.b.x, r.b.y); v[3] = make_float2(r.b.z, r.b.w);
  }
  __device__ static void store2(float* p, const float2 (&v)[4]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  }
  __device__ static void load(const float* p, float (&v)[8]) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ static void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  struct Raw { uint4 u; };
  __device__ static Raw ldraw(const __nv_bfloat16* p) { Raw r; r.u = *reinterpret_cast<const uint4*>(p); return r; }
  __device__ static void unpack(const Raw& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void unpack2(const Raw& r, float2 (&v)[4]) {     // one shift / one mask per element
    const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
  }
  __device__ static void store2(__nv_bfloat16* p, const float2 (&v)[4]) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y), pack_bf16(v[3].x, v[3].y));
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

// Packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): one issue slot per TWO IEEE fp32 operations, bit-identical to
// the scalar forms.  The wide LayerNorm(+GELU) kernels are issue-bound (ncu: 31 instructions per element, 67% issue
// utilisation at 2.9 TB/s), so halving the FP32 instruction count is what moves them towards the HBM roofline.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)),
      "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 k2(float c) { return make_float2(c, c); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// erf-GELU for the bf16 kernels: Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, far below bf16 resolution) with
// approximate MUFU.RCP / MUFU.EX2 (no IEEE fix-up paths): h = Phi(-|x|) = 0.5*erfc(|x|/sqrt2) costs 2 MUFU + 8 FP32
// instructions and also yields exp(-x^2/2) for the derivative; the fp32 parity mode keeps erff.
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float half_erfc_abs(float x, float& e) {    // e = exp(-x^2 / 2)
  const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  e = fast_ex2(x * x * (-0.5f * 1.44269504088896340736f));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  return poly * t * e;
}
template <typename T>
__device__ __forceinline__ float gelu_value(float x) {
  if (sizeof(T) == 4) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  float e;
  const float h = half_erfc_abs(x, e);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));                // x >= 0: x - x*h;  x < 0: x*h
}
// two elements at once; nh = -Phi(-|x|) comes out of the (negated) polynomial, e = exp(-x^2/2), ax = |x|
__device__ __forceinline__ float2 neg_half_erfc_abs2(float2 x, float2& ax, float2& e) {
  ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 u = ffma2(ax, k2(0.3275911f * 0.70710678118654752440f), k2(1.0f));
  const float2 t = make_float2(fast_rcp(u.x), fast_rcp(u.y));
  const float2 w = fmul2(fmul2(x, x), k2(-0.5f * 1.44269504088896340736f));
  e = make_float2(fast_ex2(w.x), fast_ex2(w.y));
  float2 poly = ffma2(k2(-0.5f * 1.061405429f), t, k2(0.5f * 1.453152027f));
  poly = ffma2(poly, t, k2(-0.5f * 1.421413741f));
  poly = ffma2(poly, t, k2(0.5f * 0.284496736f));
  poly = ffma2(poly, t, k2(-0.5f * 0.254829592f));
  return fmul2(fmul2(poly, t), e);
}
template <typename T>
__device__ __forceinline__ float2 gelu_value2(float2 x) {
  if (sizeof(T) == 4) return make_float2(gelu_value<T>(x.x), gelu_value<T>(x.y));
  float2 ax, e;
  const float2 nh = neg_half_erfc_abs2(x, ax, e);
  return ffma2(ax, nh, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));       // x >= 0: x - x*h;  x < 0: x*h
}
template <typename T>
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_x) {   // cdf = Phi(x), pdf_x = x * phi(x)
  if (sizeof(T) == 4) {
    cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    pdf_x = x * 0.39894228040143267794f * __expf(-0.5f * x * x);
  } else {
    float e;
    const float h = half_erfc_abs(x, e);
    cdf = x >= 0.f ? 1.0f - h : h;
    pdf_x = x * 0.39894228040143267794f * e;
  }
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward.  One warp per row, row cached in registers (NV vectors of 8 per lane), two-pass variance.
//   y = LN(f(x)) * gamma + beta (+ resid),  f = identity | gelu
// ------------------------------------------------------------------------------------------------
template <typename T, int NV, int GELU>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                     const T* __restrict__ beta, const T* __restrict__ resid,
                                                     T* __restrict__ y, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int rows, int C, float eps) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const T* xr = x + (size_t)row * C;
  // the whole row is requested before any arithmetic (NV independent 16-byte loads per lane in flight)
  typename Vec8<T>::Raw raw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    raw[i] = Vec8<T>::ldraw(xr + (c < C ? c : 0));
  }
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    Vec8<T>::unpack(raw[i], v[i]);
    if (c < C) {
      if (GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = gelu_value<T>(v[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q = fmaf(d, d, q); }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  const float nmr = -mean * rstd;
  // narrow rows: the residual row is requested up front as well (it is the only other HBM stream)
  constexpr bool kHoistResid = NV <= 4;
  typename Vec8<T>::Raw rres[kHoistResid ? NV : 1];
  if (kHoistResid && resid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      rres[i] = Vec8<T>::ldraw(resid + (size_t)row * C + (c < C ? c : 0));
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < C) {
      float g[8], b[8], o[8];
      Vec8<T>::load(gamma + c, g);
      Vec8<T>::load(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(v[i][j], rstd, nmr), g[j], b[j]);
      if (resid) {
        float r[8];
        if (kHoistResid) Vec8<T>::unpack(rres[i], r);
        else Vec8<T>::load(resid + (size_t)row * C + c, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      Vec8<T>::store(y + (size_t)row * C + c, o);
    }
  }
}

// Wide rows (C > 1536, the FFN's LN(GELU(fc1))): a warp per row needs ~100 live registers per lane, which leaves 3-4
// warps per scheduler and the row loads exposed (ncu: long-scoreboard stalls dominate, 2.9 TB/s).  Here a 128-thread CTA
// owns a row (VPT 16-byte vectors per thread, ~64 registers, 7-8 CTAs per SM), loops over rows persistently with the
// NEXT row's loads already in flight, keeps gamma / beta as fp32 in shared memory, and does the arithmetic two elements
// per instruction (FFMA2).
template <typename T, int VPT, int GELU>
__global__ void __launch_bounds__(128, 7) ln_fwd_wide_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                             const T* __restrict__ beta, const T* __restrict__ resid,
                                                             T* __restrict__ y, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, int rows, int C, float eps) {
  extern __shared__ __align__(16) float ln_gb[];     // gamma[C] | beta[C] in fp32, converted once per CTA
  __shared__ float red_s[2][4], red_q[2][4];
  pdl_sync();
  for (int c = threadIdx.x * 8; c < C; c += 128 * 8) {
    float g[8], b[8];
    Vec8<T>::load(gamma + c, g);
    Vec8<T>::load(beta + c, b);
    Vec8<float>::store(ln_gb + c, g);
    Vec8<float>::store(ln_gb + C + c, b);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_c = 1.f / (float)C;
  int cs[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) cs[i] = (i * 128 + threadIdx.x) * 8;
  typename Vec8<T>::Raw nxt[VPT];
  int row = blockIdx.x;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) nxt[i] = Vec8<T>::ldraw(x + (size_t)row * C + (cs[i] < C ? cs[i] : 0));
  }
  for (int it = 0; row < rows; row += gridDim.x, it ^= 1) {
    float2 v[VPT][4];
#pragma unroll
    for (int i = 0; i < VPT; ++i) Vec8<T>::unpack2(nxt[i], v[i]);
    const int row_n = row + gridDim.x;
    if (row_n < rows) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) nxt[i] = Vec8<T>::ldraw(x + (size_t)row_n * C + (cs[i] < C ? cs[i] : 0));
    }
    float2 s2 = k2(0.f);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (cs[i] < C) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (GELU) v[i][j] = gelu_value2<T>(v[i][j]);
          s2 = fadd2(s2, v[i][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = k2(0.f);
      }
    }
    const float sw = warp_sum(s2.x + s2.y);
    if (lane == 0) red_s[it][warp] = sw;
    __syncthreads();
    const float mean = ((red_s[it][0] + red_s[it][1]) + (red_s[it][2] + red_s[it][3])) * inv_c;
    const float2 nm2 = k2(-mean);
    float2 q2 = k2(0.f);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (cs[i] < C) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 d = fadd2(v[i][j], nm2); q2 = ffma2(d, d, q2); }
      }
    }
    const float qw = warp_sum(q2.x + q2.y);
    if (lane == 0) red_q[it][warp] = qw;
    __syncthreads();
    const float rstd = rsqrtf(((red_q[it][0] + red_q[it][1]) + (red_q[it][2] + red_q[it][3])) * inv_c + eps);
    if (threadIdx.x == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    const float2 rs2 = k2(rstd), nmr2 = k2(-mean * rstd);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = cs[i];
      if (c < C) {
        const float4 g0 = *reinterpret_cast<const float4*>(ln_gb + c), g1 = *reinterpret_cast<const float4*>(ln_gb + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(ln_gb + C + c), b1 = *reinterpret_cast<const float4*>(ln_gb + C + c + 4);
        const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
        const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        float2 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ffma2(ffma2(v[i][j], rs2, nmr2), g[j], b[j]);
        if (resid) {
          float2 r[4];
          Vec8<T>::unpack2(Vec8<T>::ldraw(resid + (size_t)row * C + c), r);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fadd2(o[j], r[j]);
        }
        Vec8<T>::store2(y + (size_t)row * C + c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  Persistent CTAs loop over rows; dgamma/dbeta partials stay in registers and are written as
// [gridDim.x, C] fp32 partials that ln_bwd_reduce_kernel sums.   dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) [* gelu'(x)]
// ------------------------------------------------------------------------------------------------
template <typename T, int GELU, int U>
__global__ void __launch_bounds__(U == 2 ? 256 : 512, U == 2 ? 3 : 2) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                     const T* __restrict__ gamma, const float* __restrict__ mean_in,
                                                     const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                     float* __restrict__ part_g, float* __restrict__ part_b, int rows,
                                                     int C, const T* __restrict__ dskip) {
  // U = rows per iteration: 2 for the plain variant up to 2048 columns (latency-bound), 1 otherwise (GELU is erf-bound)
  // blockDim.x = ceil(C/8) rounded up to a warp: thread t owns columns 8t..8t+7 of every row this CTA visits, so the
  // dgamma / dbeta partials need no cross-thread reduction; erf is evaluated once per element (GELU value and derivative).
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  __shared__ float red[2][4][16];
  const int c = threadIdx.x * 8;
  const bool act = c < C;
  float ag[8], ab[8], gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ag[j] = 0.f; ab[j] = 0.f; gm[j] = 0.f; }
  if (act) Vec8<T>::load(gamma + c, gm);
  // two rows per iteration: both rows' loads are in flight together and one barrier serves both reductions
  int it = 0;
  for (int row = blockIdx.x; row < rows; row += U * gridDim.x, it ^= 1) {
    const int row1 = row + gridDim.x;
    const bool has1 = U == 2 && row1 < rows;
    const int rr[2] = {row, has1 ? row1 : row};
    float mean[U], rstd[U], xh[U][8], gp[U][8], s1[U], s2[U];
    typename Vec8<T>::Raw rd[U], rs[U], rx[U];     // dy and the skip-branch gradient stay packed until they are used
#pragma unroll
    for (int u = 0; u < U; ++u) {
      mean[u] = mean_in[rr[u]]; rstd[u] = rstd_in[rr[u]];
      if (act) {
        rd[u] = Vec8<T>::ldraw(dy + (size_t)rr[u] * C + c);
        rx[u] = Vec8<T>::ldraw(x + (size_t)rr[u] * C + c);
        if (dskip) rs[u] = Vec8<T>::ldraw(dskip + (size_t)rr[u] * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s1[u] = 0.f; s2[u] = 0.f;
      if (act && (u == 0 || has1)) {
        float d[8], raw[8];
        Vec8<T>::unpack(rd[u], d);
        Vec8<T>::unpack(rx[u], raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float fx = raw[j];
          gp[u][j] = 1.f;
          if (GELU) {
            float cdf, pdfx;
            gelu_parts<T>(raw[j], cdf, pdfx);
            fx = raw[j] * cdf;
            gp[u][j] = cdf + pdfx;
          }
          xh[u][j] = (fx - mean[u]) * rstd[u];
          const float g = d[j] * gm[j];
          s1[u] += g;
          s2[u] += g * xh[u][j];
          ag[j] += d[j] * xh[u][j];
          ab[j] += d[j];
        }
      }
      s1[u] = warp_sum(s1[u]);
      s2[u] = warp_sum(s2[u]);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) { red[it][2 * u][warp] = s1[u]; red[it][2 * u + 1][warp] = s2[u]; }
    }
    __syncthreads();
    float t[2 * U];
#pragma unroll
    for (int k = 0; k < 2 * U; ++k) t[k] = 0.f;
    for (int w = 0; w < nwarp; ++w) {
#pragma unroll
      for (int k = 0; k < 2 * U; ++k) t[k] += red[it][k][w];
    }
    if (act) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u == 0 || has1) {
          const float m1 = t[2 * u] / C, m2 = t[2 * u + 1] / C;
          float o[8], d[8];
          Vec8<T>::unpack(rd[u], d);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd[u] * (d[j] * gm[j] - m1 - xh[u][j] * m2) * (GELU ? gp[u][j] : 1.f);
          if (dskip) {
            float sk[8];
            Vec8<T>::unpack(rs[u], sk);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += sk[j];
          }
          Vec8<T>::store(dx + (size_t)rr[u] * C + c, o);
        }
      }
    }
  }
  if (act) {
    float* pg = part_g + (size_t)blockIdx.x * C + c;
    float* pb = part_b + (size_t)blockIdx.x * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[j] = ag[j]; pb[j] = ab[j]; }
  }
}

// Wide rows (the FFN's LN(GELU(fc1)) over 3072 columns): a CTA spans a whole row, so only 2 CTAs fit an SM and loads held
// in registers keep too few bytes in flight (one row per CTA: latency-bound at ~40% of HBM).  Here the rows stream through
// a shared-memory ring filled by 1-D bulk copies (TMA, mbarrier completion) kStages rows ahead of the math; registers
// only hold the row being processed.  Same thread mapping, reductions and partials as ln_bwd_kernel.
template <typename T, int GELU, int MAXT>
__global__ void __launch_bounds__(MAXT, 2) ln_bwd_staged_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                const T* __restrict__ gamma, const float* __restrict__ mean_in,
                                                                const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                                float* __restrict__ part_g, float* __restrict__ part_b,
                                                                int rows, int C, int stages) {
  extern __shared__ __align__(128) unsigned char ln_ring[];     // stages x (dy row | x row)
  __shared__ __align__(8) uint64_t full[8];
  __shared__ __align__(16) float red[2][2][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)C * sizeof(T);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_sync();
  const int n_my = (rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // rows blockIdx.x + i * gridDim.x
  auto fill = [&](int i, int s) {
    const size_t row = (size_t)blockIdx.x + (size_t)i * gridDim.x;
    unsigned char* dst = ln_ring + (size_t)s * 2 * row_bytes;
    mbar_expect_tx(&full[s], 2 * row_bytes);
    bulk_load_1d(dst, dy + row * C, row_bytes, &full[s]);
    bulk_load_1d(dst + row_bytes, x + row * C, row_bytes, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < stages && i < n_my; ++i) fill(i, i);
  const int c = threadIdx.x * 8;
  const bool act = c < C;
  const float inv_c = 1.f / (float)C;
  float ag[8], ab[8], gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ag[j] = 0.f; ab[j] = 0.f; gm[j] = 0.f; }
  if (act) Vec8<T>::load(gamma + c, gm);
  float mean_n = n_my > 0 ? mean_in[blockIdx.x] : 0.f, rstd_n = n_my > 0 ? rstd_in[blockIdx.x] : 0.f;
  if (threadIdx.x < 64) (&red[0][0][0])[threadIdx.x] = 0.f;      // warps beyond nwarp contribute zeros to the fixed-size sums
  __syncthreads();
  int s = 0;
  uint32_t phase = 0;
  for (int i = 0; i < n_my; ++i) {
    const int it = i & 1;
    const size_t row = (size_t)blockIdx.x + (size_t)i * gridDim.x;
    const float rstd = rstd_n, nmr = -mean_n * rstd_n;
    if (i + 1 < n_my) { mean_n = mean_in[row + gridDim.x]; rstd_n = rstd_in[row + gridDim.x]; }
    mbar_wait(&full[s], phase);
    float d[8], xh[8], gp[8], s1 = 0.f, s2 = 0.f;
    if (act) {
      const T* srow = reinterpret_cast<const T*>(ln_ring + (size_t)s * 2 * row_bytes);
      float raw[8];
      Vec8<T>::load(srow + c, d);
      Vec8<T>::load(srow + C + c, raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float fx = raw[j];
        gp[j] = 1.f;
        if (GELU) {
          float cdf, pdfx;
          gelu_parts<T>(raw[j], cdf, pdfx);
          fx = raw[j] * cdf;
          gp[j] = cdf + pdfx;
        }
        xh[j] = fmaf(fx, rstd, nmr);
        const float g = d[j] * gm[j];
        s1 += g;
        s2 = fmaf(g, xh[j], s2);
        ag[j] = fmaf(d[j], xh[j], ag[j]);
        ab[j] += d[j];
        d[j] = g;
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { red[it][0][warp] = s1; red[it][1][warp] = s2; }
    __syncthreads();                      // every thread has its slice of the stage in registers: the slot can be refilled
    if (threadIdx.x == 0 && i + stages < n_my) fill(i + stages, s);
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float4 p1 = reinterpret_cast<const float4*>(red[it][0])[w], p2 = reinterpret_cast<const float4*>(red[it][1])[w];
      t1 += (p1.x + p1.y) + (p1.z + p1.w);
      t2 += (p2.x + p2.y) + (p2.z + p2.w);
    }
    if (act) {
      const float m1r = -t1 * inv_c * rstd, m2r = -t2 * inv_c * rstd;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[j], m2r, fmaf(d[j], rstd, m1r)) * (GELU ? gp[j] : 1.f);
      Vec8<T>::store(dx + row * C + c, o);
    }
    if (++s == stages) { s = 0; phase ^= 1u; }
  }
  if (act) {
    float* pg = part_g + (size_t)blockIdx.x * C + c;
    float* pb = part_b + (size_t)blockIdx.x * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[j] = ag[j]; pb[j] = ab[j]; }
  }
}

// sums the [nparts, C] fp32 partials: block = 8 columns x 32 partial-lanes (one 32-byte sector per row and lane group);
// C/8 x 2 CTAs, so that the few hundred KB .. MB of partials are read by ~200 CTAs instead of ~50
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b,
                                                            int nparts, int C, T* __restrict__ dgamma, T* __restrict__ dbeta,
                                                            int accumulate) {
  pdl_sync();
  const int cx = threadIdx.x & 7, py = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cx;
  const float* src = blockIdx.y ? part_b : part_g;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    int p = py;
    for (; p + 32 < nparts; p += 64) { s0 += src[(size_t)p * C + c]; s1 += src[(size_t)(p + 32) * C + c]; }
    if (p < nparts) s0 += src[(size_t)p * C + c];
  }
  __shared__ float red[32][9];
  red[py][cx] = s0 + s1;
  __syncthreads();
  if (py == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][cx];
    T* dst = (blockIdx.y ? dbeta : dgamma) + c;
    if (accumulate) t += (float)*dst;
    *dst = (T)t;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums: out[c] = sum_r x[r, c]      (bias gradients)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, long long ld, int rows, int C,
                                                             float* __restrict__ part, int vec_ok) {
  pdl_sync();
  // block = 32 lanes (8 consecutive columns each -> 256 columns) x 8 row-lanes; grid.x = column tiles, grid.y = row slices
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cx * 8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (c < C) {
    if (vec_ok && c + 8 <= C) {
      for (int r = blockIdx.y * 8 + ry; r < rows; r += gridDim.y * 8) {
        float v[8];
        Vec8<T>::load(x + (size_t)r * ld + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      }
    } else {
      for (int r = blockIdx.y * 8 + ry; r < rows; r += gridDim.y * 8)
        for (int j = 0; j < 8 && c + j < C; ++j) s[j] += (float)x[(size_t)r * ld + c + j];
    }
  }
  __shared__ float red[8][32 * 8 + 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cx * 8 + j] = s[j];
  __syncthreads();
  const int col = threadIdx.x;   // 256 threads -> 256 columns of this tile
  if (blockIdx.x * 256 + col < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][col];
    part[(size_t)blockIdx.y * C + blockIdx.x * 256 + col] = t;
  }
}
template <typename T>
__global__ void colsum_final_kernel(const float* __restrict__ part, int nparts, int C, T* __restrict__ out, float alpha,
                                    int accumulate) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * C + c];
  s *= alpha;
  if (accumulate) s += (float)out[c];
  out[c] = (T)s;
}

// ------------------------------------------------------------------------------------------------
// embedding gather (+ optional broadcast add vector, e.g. the type embedding row) and scatter-add backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_gather_kernel(const long long* __restrict__ idx, const T* __restrict__ table,
                                    const T* __restrict__ addvec, T* __restrict__ out, long long ldo, int rows, int C) {
  pdl_sync();
  const int row = blockIdx.x;
  const T* src = table + (size_t)idx[row] * C;
  T* dst = out + (size_t)row * ldo;
  for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
    float v[8];
    Vec8<T>::load(src + c, v);
    if (addvec) {
      float a[8];
      Vec8<T>::load(addvec + c, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    Vec8<T>::store(dst + c, v);
  }
}
__device__ __forceinline__ void atomic_add2(float* p, float a, float b) { atomicAdd(p, a); atomicAdd(p + 1, b); }
__device__ __forceinline__ void atomic_add2(__nv_bfloat16* p, float a, float b) {
  atomicAdd(reinterpret_cast<__nv_bfloat162*>(p), __floats2bfloat162_rn(a, b));
}
template <typename T>
__global__ void embed_scatter_add_kernel(const long long* __restrict__ idx, const T* __restrict__ dout, long long ldo,
                                         T* __restrict__ dtable, int rows, int C, long long skip_idx) {
  pdl_sync();
  const int row = blockIdx.x;
  const long long id = idx[row];
  if (id == skip_idx) return;  // padding_idx rows receive no gradient (nn.Embedding semantics)
  const T* src = dout + (size_t)row * ldo;
  T* dst = dtable + (size_t)id * C;
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) atomic_add2(dst + c, (float)src[c], (float)src[c + 1]);
}

// ------------------------------------------------------------------------------------------------
// fp32 -> 3-way bf16 split (x = x0 + x1 + x2 to ~2^-24), laid out as 6 blocks for the K-concatenated GEMM:
//   pattern 0 (A side): x0 x1 x2 x0 x1 x0      pattern 1 (B side): x0 x0 x0 x1 x1 x2
//   out[blk * blk_stride + r * ldo + c]
// ------------------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ x, long long ldx, int rows, int C, __nv_bfloat16* __restrict__ out,
                              long long ldo, long long blk_stride, int pattern) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  const int r = (int)(i / C), c = (int)(i % C);
  const float v = x[(size_t)r * ldx + c];
  const __nv_bfloat16 h0 = __float2bfloat16(v);
  const float r1 = v - __bfloat162float(h0);
  const __nv_bfloat16 h1 = __float2bfloat16(r1);
  const __nv_bfloat16 h2 = __float2bfloat16(r1 - __bfloat162float(h1));
  const __nv_bfloat16 pa[6] = {h0, h1, h2, h0, h1, h0};
  const __nv_bfloat16 pb[6] = {h0, h0, h0, h1, h1, h2};
  __nv_bfloat16* o = out + (size_t)r * ldo + c;
#pragma unroll
  for (int b = 0; b < 6; ++b) o[b * blk_stride] = pattern ? pb[b] : pa[b];
}

// out = a + b  |  out = a * rowmask (zero padded rows)  -- glue that has no natural producer to fuse into yet
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float x[8], y[8];
    Vec8<T>::load(a + i, x);
    Vec8<T>::load(b + i, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    Vec8<T>::store(out + i, x);
  } else {
    for (long long k = i; k < n; ++k) out[k] = (T)((float)a[k] + (float)b[k]);
  }
}
template <typename T>
__global__ void mask_rows_kernel(T* __restrict__ x, const unsigned char* __restrict__ rowmask, int rows, int C) {
  pdl_sync();
  const int row = blockIdx.x;
  if (!rowmask[row]) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) x[(size_t)row * C + c] = (T)0.f;
}
template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (T)gelu_erf((float)x[i]);
}
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = (T)((float)dy[i] * gelu_grad((float)x[i]));
}

int g_ln_bwd_staged = 1;   // A/B switch (ofa_layernorm_set_staged)

template <typename T, int NV>
int ln_fwd_launch(const void* x, const void* g, const void* b, const void* r, void* y, float* mean, float* rstd,
                  int rows, int C, float eps, int gelu_in, cudaStream_t st) {
  if (NV >= 8) {      // wide rows: one CTA per row, persistent
    constexpr int VPT = NV >= 8 ? NV / 4 : 1;
    auto kern = gelu_in ? ln_fwd_wide_kernel<T, VPT, 1> : ln_fwd_wide_kernel<T, VPT, 0>;
    const int cap = 148 * 7;
    OFA_CUDA(ofa_launch_pdl(kern, rows < cap ? rows : cap, 128, 2 * (size_t)C * sizeof(float), st, (const T*)x, (const T*)g,
                            (const T*)b, (const T*)r, (T*)y, mean, rstd, rows, C, eps));
  } else {
    auto kern = gelu_in ? ln_fwd_kernel<T, NV, 1> : ln_fwd_kernel<T, NV, 0>;
    OFA_CUDA(ofa_launch_pdl(kern, (rows + 3) / 4, 128, 0, st, (const T*)x, (const T*)g, (const T*)b, (const T*)r, (T*)y, mean,
                            rstd, rows, C, eps));
  }
  OFA_LAUNCH_CHECK("ln_fwd_kernel");
  return 0;
}
template <typename T>
int ln_bwd_launch(const void* dy, const void* x, const void* g, const float* mean, const float* rstd, void* dx,
                  float* pg, float* pb, int nparts, int rows, int C, int gelu_in, const void* dskip, cudaStream_t st) {
  const int threads = ((C / 8 + 31) / 32) * 32;
  // wide rows: shared-memory ring (2 CTAs per SM, up to ~100 KB each)
  const size_t row_pair = 2 * (size_t)C * sizeof(T);
  int stages = (int)((100 * 1024) / row_pair);
  if (stages > 6) stages = 6;
  if (g_ln_bwd_staged && threads > 256 && stages >= 2 && (row_pair % 32) == 0 && !dskip) {
    const size_t smem = stages * row_pair;
    const int v = (gelu_in ? 1 : 0) + (threads <= 384 ? 2 : 0);
    auto kern = v == 3 ? ln_bwd_staged_kernel<T, 1, 384> : v == 2 ? ln_bwd_staged_kernel<T, 0, 384>
              : v == 1 ? ln_bwd_staged_kernel<T, 1, 512> : ln_bwd_staged_kernel<T, 0, 512>;
    static bool attr_done[4] = {false, false, false, false};
    if (!attr_done[v]) {
      OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      attr_done[v] = true;
    }
    OFA_CUDA(ofa_launch_pdl(kern, nparts, threads, smem, st, (const T*)dy, (const T*)x, (const T*)g, mean, rstd, (T*)dx, pg, pb, rows, C, stages));
    OFA_LAUNCH_CHECK("ln_bwd_staged_kernel");
    return 0;
  }
  if (gelu_in)
    OFA_CUDA(ofa_launch_pdl(ln_bwd_kernel<T, 1, 1>, nparts, threads, 0, st, (const T*)dy, (const T*)x, (const T*)g, mean, rstd, (T*)dx, pg, pb, rows, C, (const T*)dskip));
  else if (threads <= 256)
    OFA_CUDA(ofa_launch_pdl(ln_bwd_kernel<T, 0, 2>, nparts, threads, 0, st, (const T*)dy, (const T*)x, (const T*)g, mean, rstd, (T*)dx, pg, pb, rows, C, (const T*)dskip));
  else
    OFA_CUDA(ofa_launch_pdl(ln_bwd_kernel<T, 0, 1>, nparts, threads, 0, st, (const T*)dy, (const T*)x, (const T*)g, mean, rstd, (T*)dx, pg, pb, rows, C, (const T*)dskip));
  OFA_LAUNCH_CHECK("ln_bwd_kernel");
  return 0;
}

}  // namespace

#define DISPATCH_NV(NVV, CALL)                                  \
  switch (NVV) {                                                \
    case 1: { constexpr int NV = 1; return CALL; }              \
    case 2: { constexpr int NV = 2; return CALL; }              \
    case 3: { constexpr int NV = 3; return CALL; }              \
    case 4: { constexpr int NV = 4; return CALL; }              \
    case 5: case 6: { constexpr int NV = 6; return CALL; }      \
    case 7: case 8: { constexpr int NV = 8; return CALL; }      \
    case 9: case 10: case 11: case 12: { constexpr int NV = 12; return CALL; } \
    case 13: case 14: case 15: case 16: { constexpr int NV = 16; return CALL; } \
    default: return ofa_set_error("layernorm: C=%d too wide (max 4096)", C); \
  }

#define DISPATCH_NV_BWD(NVV, CALL)                              \
  switch (NVV) {                                                \
    case 1: { constexpr int NV = 1; return CALL; }              \
    case 2: { constexpr int NV = 2; return CALL; }              \
    case 3: { constexpr int NV = 3; return CALL; }              \
    case 4: { constexpr int NV = 4; return CALL; }              \
    default: return ofa_set_error("layernorm bwd: C=%d too wide (max 4096)", C); \
  }

extern "C" int ofa_layernorm_fwd(const void* x, const void* gamma, const void* beta, const void* resid, void* y,
                                 float* mean, float* rstd, int rows, int C, float eps, int gelu_in, int dtype,
                                 void* stream) {
  OFA_CHECK(rows > 0 && C > 0 && C % 8 == 0, "ofa_layernorm_fwd: rows=%d C=%d (C must be a multiple of 8)", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (C + 255) / 256;
  if (dtype == OFA_BF16) {
    DISPATCH_NV(nv, (ln_fwd_launch<__nv_bfloat16, NV>(x, gamma, beta, resid, y, mean, rstd, rows, C, eps, gelu_in, st)))
  } else if (dtype == OFA_F32) {
    DISPATCH_NV(nv, (ln_fwd_launch<float, NV>(x, gamma, beta, resid, y, mean, rstd, rows, C, eps, gelu_in, st)))
  }
  return ofa_set_error("ofa_layernorm_fwd: bad dtype %d", dtype);
}

extern "C" int ofa_layernorm_set_staged(int enabled) {
  const int old = g_ln_bwd_staged;
  g_ln_bwd_staged = enabled;
  return old;
}

extern "C" int ofa_layernorm_bwd_nparts(int rows) {
  return rows < 1184 ? rows : 1184;  // 8 CTAs per SM on 148 SMs
}
// wide rows (C >= 2048: 256+ threads per CTA, at most 2-3 CTAs resident per SM) use fewer persistent CTAs, which also
// quarters the [nparts, C] fp32 partials the reduce kernel has to re-read
static int ln_bwd_nparts_for(int rows, int C) {
  const int cap = C >= 2048 ? 296 : (C >= 1024 ? 592 : 1184);
  return rows < cap ? rows : cap;
}

extern "C" int ofa_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                 const float* rstd, void* dx, void* dgamma, void* dbeta, float* workspace, int rows,
                                 int C, int gelu_in, int accumulate, const void* dskip, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && C > 0 && C % 8 == 0, "ofa_layernorm_bwd: rows=%d C=%d", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int nparts = ln_bwd_nparts_for(rows, C);   // <= ofa_layernorm_bwd_nparts(rows): the workspace bound
  float* pg = workspace;
  float* pb = workspace + (size_t)nparts * C;  // workspace: 2 * nparts * C floats
  OFA_CHECK(C <= 4096, "ofa_layernorm_bwd: C=%d too wide (max 4096)", C);
  int rc = 1;
  if (dtype == OFA_BF16) {
    rc = ln_bwd_launch<__nv_bfloat16>(dy, x, gamma, mean, rstd, dx, pg, pb, nparts, rows, C, gelu_in, dskip, st);
    if (rc) return rc;
    OFA_CUDA(ofa_launch_pdl(ln_bwd_reduce_kernel<__nv_bfloat16>, dim3((C + 7) / 8, 2), 256, 0, st, pg, pb, nparts, C, (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta, accumulate));
  } else if (dtype == OFA_F32) {
    rc = ln_bwd_launch<float>(dy, x, gamma, mean, rstd, dx, pg, pb, nparts, rows, C, gelu_in, dskip, st);
    if (rc) return rc;
    OFA_CUDA(ofa_launch_pdl(ln_bwd_reduce_kernel<float>, dim3((C + 7) / 8, 2), 256, 0, st, pg, pb, nparts, C, (float*)dgamma, (float*)dbeta, accumulate));
  } else {
    return ofa_set_error("ofa_layernorm_bwd: bad dtype %d", dtype);
  }
  OFA_LAUNCH_CHECK("ln_bwd_reduce_kernel");
  return 0;
}

extern "C" int ofa_colsum(const void* x, long long ld, int rows, int C, void* out, float* workspace, float alpha,
                          int accumulate, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && C > 0, "ofa_colsum: rows=%d C=%d", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  int ny = (rows + 31) / 32;
  if (ny > 32) ny = 32;  // workspace: 64 * C floats
  dim3 grid((C + 255) / 256, ny);
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  const int vec_ok = (((uintptr_t)x & 15) == 0) && ((ld * esz) % 16 == 0);
  if (dtype == OFA_BF16) {
    OFA_CUDA(ofa_launch_pdl(colsum_partial_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, ld, rows, C, workspace, vec_ok));
    OFA_CUDA(ofa_launch_pdl(colsum_final_kernel<__nv_bfloat16>, (C + 127) / 128, 128, 0, st, workspace, ny, C, (__nv_bfloat16*)out, alpha, accumulate));
  } else if (dtype == OFA_F32) {
    OFA_CUDA(ofa_launch_pdl(colsum_partial_kernel<float>, grid, 256, 0, st, (const float*)x, ld, rows, C, workspace, vec_ok && (ld * 4) % 32 == 0 && ((uintptr_t)x & 31) == 0));
    OFA_CUDA(ofa_launch_pdl(colsum_final_kernel<float>, (C + 127) / 128, 128, 0, st, workspace, ny, C, (float*)out, alpha, accumulate));
  } else {
    return ofa_set_error("ofa_colsum: bad dtype %d", dtype);
  }
  OFA_LAUNCH_CHECK("colsum");
  return 0;
}

extern "C" int ofa_embed_gather(const long long* idx, const void* table, const void* addvec, void* out, long long ldo,
                                int rows, int C, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && C % 8 == 0, "ofa_embed_gather: rows=%d C=%d", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(embed_gather_kernel<__nv_bfloat16>, rows, 128, 0, st, idx, (const __nv_bfloat16*)table, (const __nv_bfloat16*)addvec, (__nv_bfloat16*)out, ldo, rows, C));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(embed_gather_kernel<float>, rows, 128, 0, st, idx, (const float*)table, (const float*)addvec, (float*)out, ldo, rows, C));
  else
    return ofa_set_error("ofa_embed_gather: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("embed_gather_kernel");
  return 0;
}

extern "C" int ofa_embed_scatter_add(const long long* idx, const void* dout, long long ldo, void* dtable, int rows,
                                     int C, long long skip_idx, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && C % 2 == 0, "ofa_embed_scatter_add: rows=%d C=%d", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(embed_scatter_add_kernel<__nv_bfloat16>, rows, 128, 0, st, idx, (const __nv_bfloat16*)dout, ldo, (__nv_bfloat16*)dtable, rows, C, skip_idx));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(embed_scatter_add_kernel<float>, rows, 128, 0, st, idx, (const float*)dout, ldo, (float*)dtable, rows, C, skip_idx));
  else
    return ofa_set_error("ofa_embed_scatter_add: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("embed_scatter_add_kernel");
  return 0;
}

extern "C" int ofa_split3_bf16(const float* x, long long ldx, int rows, int C, void* out, long long ldo,
                               long long blk_stride, int pattern, void* stream) {
  OFA_CHECK(rows > 0 && C > 0, "ofa_split3_bf16: rows=%d C=%d", rows, C);
  const long long n = (long long)rows * C;
  OFA_CUDA(ofa_launch_pdl(split3_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, x, ldx, rows, C, (__nv_bfloat16*)out, ldo, blk_stride, pattern));
  OFA_LAUNCH_CHECK("split3_kernel");
  return 0;
}

extern "C" int ofa_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream) {
  OFA_CHECK(n > 0, "ofa_add: n=%lld", n);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 2047) / 2048);
  if (dtype == OFA_BF16) OFA_CUDA(ofa_launch_pdl(add_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, n));
  else if (dtype == OFA_F32) OFA_CUDA(ofa_launch_pdl(add_kernel<float>, grid, 256, 0, st, (const float*)a, (const float*)b, (float*)out, n));
  else return ofa_set_error("ofa_add: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("add_kernel");
  return 0;
}

extern "C" int ofa_mask_rows(void* x, const unsigned char* rowmask, int rows, int C, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && C > 0, "ofa_mask_rows: rows=%d C=%d", rows, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16) OFA_CUDA(ofa_launch_pdl(mask_rows_kernel<__nv_bfloat16>, rows, 128, 0, st, (__nv_bfloat16*)x, rowmask, rows, C));
  else if (dtype == OFA_F32) OFA_CUDA(ofa_launch_pdl(mask_rows_kernel<float>, rows, 128, 0, st, (float*)x, rowmask, rows, C));
  else return ofa_set_error("ofa_mask_rows: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("mask_rows_kernel");
  return 0;
}

extern "C" int ofa_gelu(const void* x, const void* dy, void* out, long long n, int backward, int dtype, void* stream) {
  OFA_CHECK(n > 0, "ofa_gelu: n=%lld", n);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == OFA_BF16) {
    if (backward) OFA_CUDA(ofa_launch_pdl(gelu_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)out, n));
    else OFA_CUDA(ofa_launch_pdl(gelu_fwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)out, n));
  } else if (dtype == OFA_F32) {
    if (backward) OFA_CUDA(ofa_launch_pdl(gelu_bwd_kernel<float>, grid, 256, 0, st, (const float*)x, (const float*)dy, (float*)out, n));
    else OFA_CUDA(ofa_launch_pdl(gelu_fwd_kernel<float>, grid, 256, 0, st, (const float*)x, (float*)out, n));
  } else return ofa_set_error("ofa_gelu: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("gelu_kernel");
  return 0;
}
