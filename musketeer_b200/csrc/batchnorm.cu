// BatchNorm2d (+ReLU, +residual) for the NHWC ResNet patch embedder: training-mode batch statistics with running-stat
// update, eval / frozen mode, forward and backward.   Tensors are viewed as [R = N*H*W, C] row-major (channels-last).
//   restates nn.BatchNorm2d / FrozenBatchNorm2d as used by models/ofa/resnet.py:113-133,211-220 and frozen_bn.py:36-57,
//   with the ReLU and the residual add of the bottleneck fused into the same passes:
//     forward : stats (1 read)              -> apply: y = relu(x*scale_c + shift_c + res)      (1-2 reads, 1 write)
//     backward: reduce (dy, x, y -> s1, s2) -> apply: dx = a_c*(g - s1/n - xhat*s2/n), dres = g (g = dy masked by y > 0)
// HBM-bound: every thread owns 8 consecutive channels (one 16-byte vector of bf16, two of fp32).
#include "common.cuh"

namespace {

template <typename T>
struct V8 {};
template <>
struct V8<float> {
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

// per-channel partial sums.  MODE 0: (sum x, sum x^2).  MODE 1: (sum g, sum g*xhat) with g = dy * (relu ? y > 0 : 1).
// block = (C/8 threads per row) x (256 / (C/8) rows); partials [gridDim.x][2][C]
template <typename T, int MODE>
__global__ void __launch_bounds__(kT) bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                       const T* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, long long R, int C, int relu,
                                                       float* __restrict__ part) {
  const int tpr = C / 8;                 // threads per row
  const int rpi = kT / tpr;              // rows per block iteration
  const int cx = threadIdx.x % tpr, ry = threadIdx.x / tpr;
  const int c0 = cx * 8;
  float a[8], b[8], mu[8], rs[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; mu[j] = 0.f; rs[j] = 1.f; sc[j] = 0.f; sh[j] = 0.f; }
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[c0 + j]; rs[j] = rstd[c0 + j]; }
    if (relu && !y) {   // no residual: the ReLU mask is recomputed from x, saving the read of y
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
    }
  }
  if (ry < rpi) {
    for (long long r = (long long)blockIdx.x * rpi + ry; r < R; r += (long long)gridDim.x * rpi) {
      float xv[8];
      V8<T>::load(x + r * C + c0, xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += xv[j]; b[j] += xv[j] * xv[j]; }
      } else {
        float g[8];
        V8<T>::load(dy + r * C + c0, g);
        if (relu) {
          if (y) {
            float yv[8];
            V8<T>::load(y + r * C + c0, yv);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaf(xv[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += g[j]; b[j] += g[j] * (xv[j] - mu[j]) * rs[j]; }
      }
    }
  }
  __shared__ float sa[kT * 8], sb[kT * 8];   // 16 KB
#pragma unroll
  for (int j = 0; j < 8; ++j) { sa[threadIdx.x * 8 + j] = a[j]; sb[threadIdx.x * 8 + j] = b[j]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kT) {
    const int tx = c / 8, j = c % 8;
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < rpi; ++q) { s0 += sa[(q * tpr + tx) * 8 + j]; s1 += sb[(q * tpr + tx) * 8 + j]; }
    part[((size_t)blockIdx.x * 2 + 0) * C + c] = s0;
    part[((size_t)blockIdx.x * 2 + 1) * C + c] = s1;
  }
}

// sum the [nparts][2][C] partials for channel c: 8 lanes per channel (threadIdx.x / 32), combined through shared memory
__device__ __forceinline__ void sum_partials(const float* __restrict__ part, int nparts, int C, int c, float& s0, float& s1) {
  __shared__ float r0[8][33], r1[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (int p = py; p < nparts; p += 8) { a += part[((size_t)p * 2) * C + c]; b += part[((size_t)p * 2 + 1) * C + c]; }
  r0[py][cx] = a; r1[py][cx] = b;
  __syncthreads();
  s0 = 0.f; s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0 += r0[i][cx]; s1 += r1[i][cx]; }
}

// forward finalize: mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, running-stat update
template <typename T>
__global__ void bn_fwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, long long R,
                                       const T* __restrict__ gamma, const T* __restrict__ beta, float eps,
                                       float momentum, T* __restrict__ running_mean, T* __restrict__ running_var,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                       float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  float s0, s1;
  sum_partials(part, nparts, C, c, s0, s1);
  if (c >= C || threadIdx.x >= 32) return;
  const float n = (float)R;
  const float mean = s0 / n;
  const float var = fmaxf(s1 / n - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  const float sc = (float)gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = (float)beta[c] - mean * sc;
  if (running_mean) {
    running_mean[c] = (T)((1.f - momentum) * (float)running_mean[c] + momentum * mean);
    running_var[c] = (T)((1.f - momentum) * (float)running_var[c] + momentum * var * n / fmaxf(n - 1.f, 1.f));
  }
}

// eval / frozen mode: scale / shift (and mean / rstd for the backward) from the running statistics
template <typename T>
__global__ void bn_eval_coeff_kernel(int C, const T* __restrict__ gamma, const T* __restrict__ beta,
                                     const T* __restrict__ running_mean, const T* __restrict__ running_var, float eps,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = (float)running_mean[c];
  const float rstd = rsqrtf((float)running_var[c] + eps);
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  const float sc = (float)gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = (float)beta[c] - mean * sc;
}

template <typename T>
__global__ void __launch_bounds__(kT) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      T* __restrict__ y, long long n8, int C, int relu) {
  const long long v = (long long)blockIdx.x * kT + threadIdx.x;
  if (v >= n8) return;
  const int c0 = (int)((v * 8) % C);
  float xv[8];
  V8<T>::load(x + v * 8, xv);
#pragma unroll
  for (int j = 0; j < 8; ++j) xv[j] = xv[j] * scale[c0 + j] + shift[c0 + j];
  if (res) {
    float rv[8];
    V8<T>::load(res + v * 8, rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] += rv[j];
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = fmaxf(xv[j], 0.f);
  }
  V8<T>::store(y + v * 8, xv);
}

// backward finalize: dgamma = s2, dbeta = s1 (optionally accumulated), coefficients for the apply pass
template <typename T>
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, long long R,
                                       const T* __restrict__ gamma, const float* __restrict__ rstd, int batch_stats,
                                       T* __restrict__ dgamma, T* __restrict__ dbeta, int accumulate,
                                       float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cc) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  float s1, s2;
  sum_partials(part, nparts, C, c, s1, s2);
  if (c >= C || threadIdx.x >= 32) return;
  if (dgamma) {
    dgamma[c] = (T)(s2 + (accumulate ? (float)dgamma[c] : 0.f));
    dbeta[c] = (T)(s1 + (accumulate ? (float)dbeta[c] : 0.f));
  }
  ca[c] = (float)gamma[c] * rstd[c];
  cb[c] = batch_stats ? s1 / (float)R : 0.f;
  cc[c] = batch_stats ? s2 / (float)R : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(kT) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                          const T* __restrict__ y, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, const float* __restrict__ ca,
                                                          const float* __restrict__ cb, const float* __restrict__ cc,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          T* __restrict__ dx, T* __restrict__ dres, long long n8, int C,
                                                          int relu) {
  const long long v = (long long)blockIdx.x * kT + threadIdx.x;
  if (v >= n8) return;
  const int c0 = (int)((v * 8) % C);
  float xv[8], g[8];
  V8<T>::load(x + v * 8, xv);
  V8<T>::load(dy + v * 8, g);
  if (relu) {
    if (y) {
      float yv[8];
      V8<T>::load(y + v * 8, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(xv[j], scale[c0 + j], shift[c0 + j]) > 0.f ? g[j] : 0.f;
    }
  }
  if (dres) V8<T>::store(dres + v * 8, g);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xh = (xv[j] - mean[c0 + j]) * rstd[c0 + j];
    o[j] = ca[c0 + j] * (g[j] - cb[c0 + j] - xh * cc[c0 + j]);
  }
  V8<T>::store(dx + v * 8, o);
}

int nparts_for(long long R, int C) {
  const int rpi = kT / (C / 8);
  long long n = (R + rpi - 1) / rpi;
  return (int)(n < 296 ? n : 296);   // 2 CTAs per SM
}

template <typename T>
int fwd_impl(const void* x, const void* res, void* y, const void* gamma, const void* beta, void* rm, void* rv,
             long long R, int C, float eps, float momentum, int training, int relu, float* stats, float* ws,
             cudaStream_t st) {
  float* mean = stats; float* rstd = stats + C; float* scale = stats + 2 * C; float* shift = stats + 3 * C; float* part = ws;
  if (training) {
    const int np = nparts_for(R, C);
    bn_reduce_kernel<T, 0><<<np, kT, 0, st>>>((const T*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, R, C, 0, part);
    bn_fwd_finalize_kernel<T><<<(C + 31) / 32, 256, 0, st>>>(part, np, C, R, (const T*)gamma, (const T*)beta, eps, momentum,
                                                             (T*)rm, (T*)rv, mean, rstd, scale, shift);
  } else {
    bn_eval_coeff_kernel<T><<<(C + 127) / 128, 128, 0, st>>>(C, (const T*)gamma, (const T*)beta, (const T*)rm, (const T*)rv,
                                                           eps, mean, rstd, scale, shift);
  }
  const long long n8 = R * C / 8;
  bn_apply_kernel<T><<<(unsigned)((n8 + kT - 1) / kT), kT, 0, st>>>((const T*)x, (const T*)res, scale, shift, (T*)y, n8, C, relu);
  OFA_LAUNCH_CHECK("batchnorm forward");
  return 0;
}

template <typename T>
int bwd_impl(const void* x, const void* dy, const void* y, const void* gamma, const float* mean, const float* rstd,
             const float* scale, const float* shift, void* dx, void* dres, void* dgamma, void* dbeta, int accumulate, long long R, int C, int batch_stats,
             int relu, float* ws, cudaStream_t st) {
  float* ca = ws; float* cb = ws + C; float* cc = ws + 2 * C; float* part = ws + 3 * C;
  const int np = nparts_for(R, C);
  bn_reduce_kernel<T, 1><<<np, kT, 0, st>>>((const T*)x, (const T*)dy, (const T*)y, mean, rstd, scale, shift, R, C, relu, part);
  bn_bwd_finalize_kernel<T><<<(C + 31) / 32, 256, 0, st>>>(part, np, C, R, (const T*)gamma, rstd, batch_stats, (T*)dgamma,
                                                           (T*)dbeta, accumulate, ca, cb, cc);
  const long long n8 = R * C / 8;
  bn_bwd_apply_kernel<T><<<(unsigned)((n8 + kT - 1) / kT), kT, 0, st>>>((const T*)x, (const T*)dy, (const T*)y, mean, rstd, ca,
                                                                      cb, cc, scale, shift, (T*)dx, (T*)dres, n8, C, relu);
  OFA_LAUNCH_CHECK("batchnorm backward");
  return 0;
}

}  // namespace

// scratch floats for either direction (partials + coefficients); `stats` of the forward is 4*C floats
// (mean | rstd | scale | shift), of which mean and rstd are the backward's inputs
extern "C" long long ofa_batchnorm_workspace_floats(int C) { return (long long)C * (3 + 2 * 592); }

extern "C" int ofa_batchnorm_fwd(const void* x, const void* res, void* y, const void* gamma, const void* beta,
                                 void* running_mean, void* running_var, long long R, int C, float eps, float momentum,
                                 int training, int relu, float* stats, float* workspace, int dtype, void* stream) {
  OFA_CHECK(R > 0 && C >= 8 && C <= 2048 && (C & (C - 1)) == 0, "ofa_batchnorm_fwd: C=%d must be a power of two in [8, 2048]", C);
  OFA_CHECK(training || (running_mean && running_var), "ofa_batchnorm_fwd: eval mode needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16) return fwd_impl<__nv_bfloat16>(x, res, y, gamma, beta, running_mean, running_var, R, C, eps, momentum, training, relu, stats, workspace, st);
  if (dtype == OFA_F32) return fwd_impl<float>(x, res, y, gamma, beta, running_mean, running_var, R, C, eps, momentum, training, relu, stats, workspace, st);
  return ofa_set_error("ofa_batchnorm_fwd: bad dtype %d", dtype);
}

extern "C" int ofa_batchnorm_bwd(const void* x, const void* dy, const void* y, const void* gamma, const float* stats,
                                 void* dx, void* dres, void* dgamma, void* dbeta, int accumulate,
                                 long long R, int C, int batch_stats, int relu, float* workspace, int dtype,
                                 void* stream) {
  OFA_CHECK(R > 0 && C >= 8 && C <= 2048 && (C & (C - 1)) == 0, "ofa_batchnorm_bwd: C=%d must be a power of two in [8, 2048]", C);
  cudaStream_t st = (cudaStream_t)stream;
  OFA_CHECK(!relu || y || stats, "ofa_batchnorm_bwd: relu needs y or the forward stats");
  const float *mean = stats, *rstd = stats + C, *scale = stats + 2 * C, *shift = stats + 3 * C;
  if (dtype == OFA_BF16) return bwd_impl<__nv_bfloat16>(x, dy, y, gamma, mean, rstd, scale, shift, dx, dres, dgamma, dbeta, accumulate, R, C, batch_stats, relu, workspace, st);
  if (dtype == OFA_F32) return bwd_impl<float>(x, dy, y, gamma, mean, rstd, scale, shift, dx, dres, dgamma, dbeta, accumulate, R, C, batch_stats, relu, workspace, st);
  return ofa_set_error("ofa_batchnorm_bwd: bad dtype %d", dtype);
}
