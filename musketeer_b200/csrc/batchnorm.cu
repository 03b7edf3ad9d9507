// BatchNorm2d (+ReLU, +residual) for the NHWC ResNet patch embedder: training-mode batch statistics with running-stat
// update, eval / frozen mode, forward and backward.   Tensors are viewed as [R = N*H*W, C] row-major (channels-last).
//   restates nn.BatchNorm2d / FrozenBatchNorm2d as used by models/ofa/resnet.py:113-133,211-220 and frozen_bn.py:36-57,
//   with the ReLU and the residual add of the bottleneck fused into the same passes:
//     forward : stats (1 read)              -> apply: y = relu(x*scale_c + shift_c + res)      (1-2 reads, 1 write)
//     backward: reduce (dy, x, y -> s1, s2) -> apply: dx = a_c*(g - s1/n - xhat*s2/n), dres = g (g = dy masked by y > 0)
// HBM-bound: every thread owns 8 consecutive channels (one 16-byte vector of bf16, two of fp32).
// Groups: the R rows may be G equal groups (the images of G tasks of a Musketeer micro-step pushed through the stem as ONE
// batch); statistics, normalisation and backward are per group (blockIdx.z), exactly as if each task had run alone, the
// running statistics are updated group after group and dgamma / dbeta are summed over the groups.
#include "common.cuh"

namespace {

template <typename T>
struct V8 {};
template <>
struct V8<float> {
  struct Raw { float4 a, b; };
  __device__ static Raw ldraw(const float* p) { Raw r; r.a = reinterpret_cast<const float4*>(p)[0]; r.b = reinterpret_cast<const float4*>(p)[1]; return r; }
  __device__ static void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
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
struct V8<__nv_bfloat16> {
  struct Raw { uint4 u; };
  __device__ static Raw ldraw(const __nv_bfloat16* p) { Raw r; r.u = *reinterpret_cast<const uint4*>(p); return r; }
  __device__ static void unpack(const Raw& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

constexpr int kT = 256;

// Thread mapping shared by all four kernels: the channel axis is cut into slabs of CS = min(C, 256) channels
// (blockIdx.y); inside a slab tpr = CS/8 threads cover one row (one 16-byte vector each) and the kT/tpr row-lanes of the
// CTA walk the rows with stride gridDim.x * rpi, kUnroll rows in flight per thread.  A thread keeps its 8 channels for the
// whole kernel, so per-channel coefficients live in registers instead of being re-read for every element.
struct Map {
  int tpr, rpi, cx, ry, c0;
  long long r0, stride;
};
__device__ __forceinline__ Map make_map(int CS) {
  Map m;
  m.tpr = CS >> 3;
  m.rpi = kT / m.tpr;
  m.cx = threadIdx.x % m.tpr;
  m.ry = threadIdx.x / m.tpr;
  m.c0 = blockIdx.y * CS + m.cx * 8;
  m.r0 = (long long)blockIdx.x * m.rpi + m.ry;
  m.stride = (long long)gridDim.x * m.rpi;
  return m;
}

// block-level reduction of per-thread (a[8], b[8]) over the row-lanes, then one atomicAdd per channel and CTA into
// sums[0][C] / sums[1][C] (zeroed by the host wrapper).  fp32 atomics: the summation order varies run to run in the last
// bits, as with the library BatchNorm this replaces.
__device__ __forceinline__ void block_accumulate(const Map& m, int C, int CS, const float (&a)[8], const float (&b)[8],
                                                 float* __restrict__ sums) {
  __shared__ float sa[kT * 8], sb[kT * 8];   // [rpi][CS]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sa[m.ry * CS + m.cx * 8 + j] = a[j];
    sb[m.ry * CS + m.cx * 8 + j] = b[j];
  }
  __syncthreads();
  if ((int)threadIdx.x < CS) {
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < m.rpi; ++q) { s0 += sa[q * CS + threadIdx.x]; s1 += sb[q * CS + threadIdx.x]; }
    atomicAdd(sums + blockIdx.y * CS + threadIdx.x, s0);
    atomicAdd(sums + C + blockIdx.y * CS + threadIdx.x, s1);
  }
}

// ---- order-independent (bit-reproducible) accumulation of the FORWARD statistics ----------------------------------------
// fp32 atomics add the CTAs' partial sums in arrival order, so batch statistics differed in the last bits from run to run;
// behind 90+ BatchNorm / ReLU layers that noise moved whole-model gradient norms by percents on small batches.
// Here every partial sum is added EXACTLY: a float is an integer mantissa times a power of two, the mantissa is shifted to
// the fixed binary position of one of kLimbs 64-bit integer accumulators per value (limb k collects least-significant-bit
// positions [kP0 + 24k, kP0 + 24k + 24)) and added with an integer atomic -- integer addition is associative, so the limbs
// hold the same bits whatever the arrival order.  The last CTA of the launch (atomic ticket) folds the limbs into the fp32
// sums the apply kernel reads, most significant first, in double precision, and clears them for the next call.  Range: values
// from 2^-57 (smaller ones lose low bits) to 2^54; non-finite partial sums poison the fp32 slot directly so NaN / inf reach
// the loss as before.  The backward statistics keep fp32 atomics: their rounding noise meets no ReLU mask on its way to the
// parameter gradients and stays at the 1e-7 level.
constexpr int kLimbs = 4, kP0 = -80;
constexpr int kSumsN = 2 * 2048 * 8;     // = 2 * kMaxC * kMaxGroups floats of fp32 sums at the start of the scratch
__device__ __forceinline__ void exact_add(unsigned long long* __restrict__ limbs, float* __restrict__ sums, int slot, float p) {
  const uint32_t bits = __float_as_uint(p);
  if ((bits & 0x7fffffffu) == 0u) return;
  int e = (int)((bits >> 23) & 0xffu);
  if (e == 0xff) { atomicAdd(sums + slot, p); return; }
  long long mant = (long long)((bits & 0x7fffffu) | (e ? 0x800000u : 0u));
  if (e == 0) e = 1;
  const int pos = e - 150 - kP0;          // position of the mantissa's least significant bit above the bottom of limb 0
  int k = 0, o = 0;
  if (pos < 0) mant >>= min(-pos, 31);
  else { k = min(pos / 24, kLimbs - 1); o = min(pos - 24 * k, 38); }
  long long v = mant << o;
  if (bits >> 31) v = -v;
  atomicAdd(limbs + (size_t)k * kSumsN + slot, (unsigned long long)v);
}
// called by every thread of every CTA of bn_stats_kernel after its exact_add calls.  One ticket per (channel slab, group): the
// CTAs that share blockIdx.y / blockIdx.z are exactly the contributors of that slab's 2 * CS slots, and the last of them folds
// just those (two values per thread) -- a single last CTA folding all 2 * C * groups values (first version) walked 8192 values
// alone behind everybody else: 14 -> 28 us per launch.
constexpr int kMaxTickets = 64;
__device__ __forceinline__ void exact_finalize(unsigned long long* __restrict__ limbs, float* __restrict__ sums, int C, int CS) {
  __shared__ int last_cta;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(limbs + (size_t)kLimbs * kSumsN) + (blockIdx.z * gridDim.y + blockIdx.y);
    last_cta = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last_cta) *ticket = 0u;
  }
  __syncthreads();
  if (!last_cta) return;
  __threadfence();
  for (int e = threadIdx.x; e < 2 * CS; e += kT) {
    const int s = (int)blockIdx.z * 2 * C + (e < CS ? 0 : C) + (int)blockIdx.y * CS + (e < CS ? e : e - CS);
    double acc = 0.0;
#pragma unroll
    for (int k = kLimbs - 1; k >= 0; --k) {
      const long long L = (long long)__ldcg(limbs + (size_t)k * kSumsN + s);
      acc += (double)L * __longlong_as_double((long long)(1023 + kP0 + 24 * k) << 52);      // 2^(kP0 + 24 k), exact
      limbs[(size_t)k * kSumsN + s] = 0ull;
    }
    sums[s] = __ldcg(sums + s) + (float)acc;      // (+ a non-finite poison value, if any)
  }
}

// The per-channel sums live in a caller-provided scratch that is ZERO ON ENTRY and LEFT ZERO ON EXIT: every CTA of the apply
// kernel bumps a counter once it has read the sums, and the last one clears sums and counter for the next BatchNorm call in
// the stream (no memset node per layer; 752 of them per training step before).  scratch = [2*C sums | ... | counter @ 2*kMaxC].
constexpr int kMaxC = 2048, kMaxGroups = 8;
__device__ __forceinline__ void release_sums(float* __restrict__ sums /* group 0 */, int C) {
  __shared__ int last;
  __syncthreads();                       // every thread of this CTA has read its sums
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(sums + 2 * kMaxC * kMaxGroups);
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    last = atomicAdd(counter, 1u) == total - 1;
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (last)
    for (int c = threadIdx.x; c < 2 * C * (int)gridDim.z; c += kT) sums[c] = 0.f;
}

constexpr int kUnroll = 4;

// forward statistics: sums[0][c] += sum x, sums[1][c] += sum x^2
template <typename T>
__global__ void __launch_bounds__(kT, 4) bn_stats_kernel(const T* __restrict__ x, long long R, int C, int CS,
                                                      float* __restrict__ sums_base, unsigned long long* __restrict__ limbs) {
  pdl_sync();
  const Map m = make_map(CS);
  x += (long long)blockIdx.z * R * C;            // this group's rows
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
  for (long long r = m.r0; r < R; r += kUnroll * m.stride) {
    typename V8<T>::Raw raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long ru = r + u * m.stride;
      raw[u] = V8<T>::ldraw(x + (ru < R ? ru : m.r0) * C + m.c0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (r + u * m.stride < R) {
        float xv[8];
        V8<T>::unpack(raw[u], xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += xv[j]; b[j] = fmaf(xv[j], xv[j], b[j]); }
      }
    }
  }
  // block-level reduction over the row-lanes (fixed order), then one exact integer accumulation per channel and CTA
  {
    __shared__ float sa[kT * 8], sb[kT * 8];   // [rpi][CS]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sa[m.ry * CS + m.cx * 8 + j] = a[j];
      sb[m.ry * CS + m.cx * 8 + j] = b[j];
    }
    __syncthreads();
    if ((int)threadIdx.x < CS) {
      float s0 = 0.f, s1 = 0.f;
      for (int q = 0; q < m.rpi; ++q) { s0 += sa[q * CS + threadIdx.x]; s1 += sb[q * CS + threadIdx.x]; }
      const int slot = (int)blockIdx.z * 2 * C + (int)blockIdx.y * CS + (int)threadIdx.x;
      exact_add(limbs, sums_base, slot, s0);
      exact_add(limbs, sums_base, slot + C, s1);
    }
  }
  exact_finalize(limbs, sums_base, C, CS);
}

// forward apply: y = relu?(x * scale_c + shift_c + res).  Every thread derives scale / shift of its 8 channels from the
// sums (training) or the running statistics (eval / frozen); CTA column 0 also publishes mean | rstd | scale | shift for
// the backward and updates the running statistics.
template <typename T>
__global__ void __launch_bounds__(kT, 3) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                      const T* __restrict__ gamma, const T* __restrict__ beta,
                                                      T* __restrict__ running_mean, T* __restrict__ running_var,
                                                      const float* __restrict__ sums, float* __restrict__ stats,
                                                      long long R, int C, int CS, float eps, float momentum, int training,
                                                      int relu) {
  pdl_sync();
  const Map m = make_map(CS);
  const float* sums0 = sums;                     // group 0 (running statistics, release)
  {
    const long long go = (long long)blockIdx.z * R * C;
    x += go; y += go;
    if (res) res += go;
    if (sums) sums += (size_t)blockIdx.z * 2 * C;
    stats += (size_t)blockIdx.z * 4 * C;
  }
  float sc[8], sh[8];
  {
    const float n = (float)R;
    const bool publish = blockIdx.x == 0 && m.ry == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = m.c0 + j;
      float mean, var;
      if (training) {
        mean = sums[c] / n;
        var = fmaxf(sums[C + c] / n - mean * mean, 0.f);
      } else {
        mean = (float)running_mean[c];
        var = (float)running_var[c];
      }
      const float rstd = rsqrtf(var + eps);
      sc[j] = (float)gamma[c] * rstd;
      sh[j] = (float)beta[c] - mean * sc[j];
      if (publish) {
        stats[c] = mean; stats[C + c] = rstd; stats[2 * C + c] = sc[j]; stats[3 * C + c] = sh[j];
        if (training && running_mean && blockIdx.z == 0) {     // group after group, as sequential per-task forwards would
          float rm = (float)running_mean[c], rvv = (float)running_var[c];
          for (int g = 0; g < (int)gridDim.z; ++g) {
            const float mg = sums0[(size_t)g * 2 * C + c] / n;
            const float vg = fmaxf(sums0[(size_t)g * 2 * C + C + c] / n - mg * mg, 0.f);
            rm = (float)(T)((1.f - momentum) * rm + momentum * mg);
            rvv = (float)(T)((1.f - momentum) * rvv + momentum * vg * n / fmaxf(n - 1.f, 1.f));
          }
          running_mean[c] = (T)rm;
          running_var[c] = (T)rvv;
        }
      }
    }
  }
  for (long long r = m.r0; r < R; r += kUnroll * m.stride) {
    typename V8<T>::Raw rx[kUnroll], rr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long ru = r + u * m.stride < R ? r + u * m.stride : m.r0;
      rx[u] = V8<T>::ldraw(x + ru * C + m.c0);
      if (res) rr[u] = V8<T>::ldraw(res + ru * C + m.c0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long ru = r + u * m.stride;
      if (ru < R) {
        float xv[8], rv[8];
        V8<T>::unpack(rx[u], xv);
        if (res) V8<T>::unpack(rr[u], rv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = fmaf(xv[j], sc[j], sh[j]);
          if (res) v += rv[j];
          xv[j] = relu ? fmaxf(v, 0.f) : v;
        }
        V8<T>::store(y + ru * C + m.c0, xv);
      }
    }
  }
  if (training) release_sums(const_cast<float*>(sums0), C);   // after the stream: nothing waits on it
}

// backward statistics: sums[0][c] += sum g, sums[1][c] += sum g * xhat, g = dy * (relu ? y > 0 : 1).  Without a residual
// the ReLU mask is recomputed from x with the forward's scale / shift (saves the read of y).
template <typename T, int U>
__global__ void __launch_bounds__(kT, U == 2 ? 3 : 2) bn_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                          const T* __restrict__ y, const float* __restrict__ stats,
                                                          long long R, int C, int CS, int relu, float* __restrict__ sums) {
  pdl_sync();
  const Map m = make_map(CS);
  {
    const long long go = (long long)blockIdx.z * R * C;
    x += go; dy += go;
    if (y) y += go;
    stats += (size_t)blockIdx.z * 4 * C;
    sums += (size_t)blockIdx.z * 2 * C;
  }
  float a[8], b[8], sc[8], sh[8];   // b accumulates sum g*x; the xhat form follows from (b - mean*a) * rstd at the end
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = m.c0 + j;
    a[j] = 0.f; b[j] = 0.f;
    sc[j] = stats[2 * C + c]; sh[j] = stats[3 * C + c];
  }
  for (long long r = m.r0; r < R; r += U * m.stride) {
    typename V8<T>::Raw rx[U], rg[U], ry_[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * m.stride < R ? r + u * m.stride : m.r0;
      rx[u] = V8<T>::ldraw(x + ru * C + m.c0);
      rg[u] = V8<T>::ldraw(dy + ru * C + m.c0);
      if (relu && y) ry_[u] = V8<T>::ldraw(y + ru * C + m.c0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * m.stride < R) {
        float xv[8], g[8], yv[8];
        V8<T>::unpack(rx[u], xv);
        V8<T>::unpack(rg[u], g);
        if (relu && y) V8<T>::unpack(ry_[u], yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = g[j];
          if (relu) {
            const float act = y ? yv[j] : fmaf(xv[j], sc[j], sh[j]);
            gg = act > 0.f ? gg : 0.f;
          }
          a[j] += gg;
          b[j] = fmaf(gg, xv[j], b[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = (b[j] - stats[m.c0 + j] * a[j]) * stats[C + m.c0 + j];
  block_accumulate(m, C, CS, a, b, sums);
}

// backward apply: dx = ca*(g - s1/n - xhat*s2/n) = ca*g + A*x + B  with  ca = gamma*rstd, A = -ca*rstd*s2/n,
// B = -ca*s1/n - A*mean  (eval / frozen statistics: A = B = 0);  dres = g.  CTA column 0 writes dgamma = s2, dbeta = s1.
template <typename T, int U>
__global__ void __launch_bounds__(kT, U == 2 ? 3 : 2) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                          const T* __restrict__ y, const T* __restrict__ gamma,
                                                          const float* __restrict__ stats, const float* __restrict__ sums,
                                                          T* __restrict__ dx, T* __restrict__ dres, T* __restrict__ dgamma,
                                                          T* __restrict__ dbeta, int accumulate, long long R, int C, int CS,
                                                          int batch_stats, int relu) {
  pdl_sync();
  const Map m = make_map(CS);
  const float* sums0 = sums;
  {
    const long long go = (long long)blockIdx.z * R * C;
    x += go; dy += go; dx += go;
    if (y) y += go;
    if (dres) dres += go;
    stats += (size_t)blockIdx.z * 4 * C;
    if (sums) sums += (size_t)blockIdx.z * 2 * C;
  }
  float ca[8], cA[8], cB[8], sc[8], sh[8];
  {
    const float inv_n = 1.f / (float)R;
    const bool publish = blockIdx.x == 0 && blockIdx.z == 0 && m.ry == 0 && dgamma != nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = m.c0 + j;
      const float mean = stats[c], rstd = stats[C + c];
      sc[j] = stats[2 * C + c]; sh[j] = stats[3 * C + c];
      const float s1 = sums ? sums[c] : 0.f, s2 = sums ? sums[C + c] : 0.f;
      ca[j] = (float)gamma[c] * rstd;
      cA[j] = batch_stats ? -ca[j] * rstd * s2 * inv_n : 0.f;
      cB[j] = batch_stats ? -ca[j] * s1 * inv_n - cA[j] * mean : 0.f;
      if (publish) {          // parameter gradients: summed over the groups
        float t1 = 0.f, t2 = 0.f;
        for (int g = 0; g < (int)gridDim.z; ++g) { t1 += sums0[(size_t)g * 2 * C + c]; t2 += sums0[(size_t)g * 2 * C + C + c]; }
        dgamma[c] = (T)(t2 + (accumulate ? (float)dgamma[c] : 0.f));
        dbeta[c] = (T)(t1 + (accumulate ? (float)dbeta[c] : 0.f));
      }
    }
  }
  for (long long r = m.r0; r < R; r += U * m.stride) {
    typename V8<T>::Raw rx[U], rg[U], ry_[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * m.stride < R ? r + u * m.stride : m.r0;
      rx[u] = V8<T>::ldraw(x + ru * C + m.c0);
      rg[u] = V8<T>::ldraw(dy + ru * C + m.c0);
      if (relu && y) ry_[u] = V8<T>::ldraw(y + ru * C + m.c0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * m.stride;
      if (ru < R) {
        float xv[8], g[8], yv[8];
        V8<T>::unpack(rx[u], xv);
        V8<T>::unpack(rg[u], g);
        if (relu && y) V8<T>::unpack(ry_[u], yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = g[j];
          if (relu) {
            const float act = y ? yv[j] : fmaf(xv[j], sc[j], sh[j]);
            gg = act > 0.f ? gg : 0.f;
          }
          g[j] = gg;
          xv[j] = fmaf(ca[j], gg, fmaf(cA[j], xv[j], cB[j]));
        }
        if (dres) V8<T>::store(dres + ru * C + m.c0, g);
        V8<T>::store(dx + ru * C + m.c0, xv);
      }
    }
  }
  if (sums) release_sums(const_cast<float*>(sums0), C);   // after the stream: nothing waits on it
}

struct Grid {
  int CS;
  dim3 g;
};
// tuning switches (ofa_batchnorm_set_tuning): resident-CTA waves of the grid-stride kernels and rows in flight per thread
// of the backward kernels
int g_bn_waves = 0 /* auto */, g_bn_bwd_unroll = 4;

Grid grid_for(long long R, int C, int unroll, int groups, int ctas_per_sm) {
  Grid r;
  r.CS = C < 256 ? C : 256;
  const int slabs = C / r.CS;
  const int rpi = kT / (r.CS / 8);
  long long gx = (R + (long long)rpi * unroll - 1) / ((long long)rpi * unroll);
  // measured (tools/rowwise_bench.py): one resident wave is best for single-slab tensors, two waves for C > 256
  const int waves = g_bn_waves > 0 ? g_bn_waves : (slabs > 1 ? 2 : 1);
  const int total = 148 * ctas_per_sm * waves;
  const long long cap = (total + slabs * groups - 1) / (slabs * groups);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  r.g = dim3((unsigned)gx, (unsigned)slabs, (unsigned)groups);
  return r;
}

template <typename T>
int fwd_impl(const void* x, const void* res, void* y, const void* gamma, const void* beta, void* rm, void* rv,
             long long R, int C, float eps, float momentum, int training, int relu, float* stats, float* ws, int groups,
             cudaStream_t st) {
  float* sums = ws;
  if (training) {
    const Grid gs = grid_for(R, C, kUnroll, groups, 4);
    unsigned long long* limbs = reinterpret_cast<unsigned long long*>(ws + kSumsN + 32);
    OFA_CUDA(ofa_launch_pdl(bn_stats_kernel<T>, gs.g, kT, 0, st, (const T*)x, R, C, gs.CS, sums, limbs));
  }
  const Grid ga = grid_for(R, C, kUnroll, groups, 3);
  OFA_CUDA(ofa_launch_pdl(bn_apply_kernel<T>, ga.g, kT, 0, st, (const T*)x, (const T*)res, (T*)y, (const T*)gamma, (const T*)beta, (T*)rm, (T*)rv,
                                          sums, stats, R, C, ga.CS, eps, momentum, training, relu));
  OFA_LAUNCH_CHECK("batchnorm forward");
  return 0;
}

template <typename T>
int bwd_impl(const void* x, const void* dy, const void* y, const void* gamma, const float* stats, void* dx, void* dres,
             void* dgamma, void* dbeta, int accumulate, long long R, int C, int batch_stats, int relu, float* ws, int groups,
             cudaStream_t st) {
  float* sums = nullptr;
  const int U = g_bn_bwd_unroll, occ = U == 2 ? 3 : 2;
  if (batch_stats || dgamma) {   // frozen statistics without parameter gradients need no reduction at all
    sums = ws;
    const Grid gs = grid_for(R, C, U, groups, occ);
    auto kern = U == 2 ? bn_bwd_stats_kernel<T, 2> : U == 4 ? bn_bwd_stats_kernel<T, 4> : bn_bwd_stats_kernel<T, 6>;
    OFA_CUDA(ofa_launch_pdl(kern, gs.g, kT, 0, st, (const T*)x, (const T*)dy, (const T*)y, stats, R, C, gs.CS, relu, sums));
  }
  const Grid ga = grid_for(R, C, U, groups, occ);
  auto kern = U == 2 ? bn_bwd_apply_kernel<T, 2> : U == 4 ? bn_bwd_apply_kernel<T, 4> : bn_bwd_apply_kernel<T, 6>;
  OFA_CUDA(ofa_launch_pdl(kern, ga.g, kT, 0, st, (const T*)x, (const T*)dy, (const T*)y, (const T*)gamma, stats, sums, (T*)dx,
                          (T*)dres, (T*)dgamma, (T*)dbeta, accumulate, R, C, ga.CS, batch_stats, relu));
  OFA_LAUNCH_CHECK("batchnorm backward");
  return 0;
}

}  // namespace

// scratch floats for either direction: ZERO ON ENTRY, left zero on exit (allocate once with zeros and reuse it for every
// call in the stream); `stats` of the forward is 4*C floats
// (mean | rstd | scale | shift), of which mean and rstd are the backward's inputs
extern "C" int ofa_batchnorm_set_tuning(int waves, int bwd_unroll) {
  if (waves >= 0 && waves <= 8) g_bn_waves = waves;   // 0 = automatic
  if (bwd_unroll == 2 || bwd_unroll == 4 || bwd_unroll == 6) g_bn_bwd_unroll = bwd_unroll;
  return 0;
}

// scratch (floats): [0, kSumsN) fp32 sums | counter | pad to 32 | kLimbs * kSumsN 64-bit limbs | kMaxTickets tickets
extern "C" long long ofa_batchnorm_workspace_floats(int C) {
  (void)C;
  static_assert(kSumsN == 2 * kMaxC * kMaxGroups, "scratch layout");
  return (long long)kSumsN + 32 + 2LL * kLimbs * kSumsN + kMaxTickets;     // one ticket per (channel slab, group)
}

extern "C" int ofa_batchnorm_fwd(const void* x, const void* res, void* y, const void* gamma, const void* beta,
                                 void* running_mean, void* running_var, long long R, int C, float eps, float momentum,
                                 int training, int relu, float* stats, float* workspace, int groups, int dtype,
                                 void* stream) {
  OFA_CHECK(groups >= 1 && groups <= kMaxGroups && R % groups == 0, "ofa_batchnorm_fwd: groups=%d must divide R and be <= 8", groups);
  R /= groups;
  OFA_CHECK(R > 0 && C >= 8 && C <= 2048 && (C & (C - 1)) == 0, "ofa_batchnorm_fwd: C=%d must be a power of two in [8, 2048]", C);
  OFA_CHECK(training || (running_mean && running_var), "ofa_batchnorm_fwd: eval mode needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16) return fwd_impl<__nv_bfloat16>(x, res, y, gamma, beta, running_mean, running_var, R, C, eps, momentum, training, relu, stats, workspace, groups, st);
  if (dtype == OFA_F32) return fwd_impl<float>(x, res, y, gamma, beta, running_mean, running_var, R, C, eps, momentum, training, relu, stats, workspace, groups, st);
  return ofa_set_error("ofa_batchnorm_fwd: bad dtype %d", dtype);
}

extern "C" int ofa_batchnorm_bwd(const void* x, const void* dy, const void* y, const void* gamma, const float* stats,
                                 void* dx, void* dres, void* dgamma, void* dbeta, int accumulate,
                                 long long R, int C, int batch_stats, int relu, float* workspace, int groups, int dtype,
                                 void* stream) {
  OFA_CHECK(groups >= 1 && groups <= kMaxGroups && R % groups == 0, "ofa_batchnorm_bwd: groups=%d must divide R and be <= 8", groups);
  R /= groups;
  OFA_CHECK(R > 0 && C >= 8 && C <= 2048 && (C & (C - 1)) == 0, "ofa_batchnorm_bwd: C=%d must be a power of two in [8, 2048]", C);
  cudaStream_t st = (cudaStream_t)stream;
  OFA_CHECK(!relu || y || stats, "ofa_batchnorm_bwd: relu needs y or the forward stats");
  if (dtype == OFA_BF16) return bwd_impl<__nv_bfloat16>(x, dy, y, gamma, stats, dx, dres, dgamma, dbeta, accumulate, R, C, batch_stats, relu, workspace, groups, st);
  if (dtype == OFA_F32) return bwd_impl<float>(x, dy, y, gamma, stats, dx, dres, dgamma, dbeta, accumulate, R, C, batch_stats, relu, workspace, groups, st);
  return ofa_set_error("ofa_batchnorm_bwd: bad dtype %d", dtype);
}
