// Fused multi-tensor optimizer step for the data-parallel training loop: gradient scaling,
// global-norm clipping and the Adam update with fp32 master weights in two launches over all parameters.
//   restates trainer.py:863-898 (multiply_grads -> clip_grad_norm -> optimizer.step) and the un-vendored fairseq
//   optim/adam.py + fp16_optimizer.py arithmetic it drives (train_musketeer.sh:136: adam, betas (0.9, 0.999), eps 1e-8,
//   weight decay 0.01 applied decoupled as p -= wd*lr*p, clip-norm 0.1):
//     g      = grad * grad_scale * min(1, clip / (||grad * grad_scale|| + 1e-6))
//     m      = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g
//     master = master - wd*lr*master - lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps) ;  param = cast(master)
// HBM-bound: 2 (grad, norm pass) + 2 + 12 (read) + 12 + 2 (write) = 30 bytes per bf16 parameter.
// The parameter list is cut into chunks of <= 65536 elements described by a device-resident table (one CTA per chunk), so
// one launch covers every tensor of the model whatever its size.
#include "common.cuh"

namespace {

struct Chunk {          // 48 bytes, mirrored by musketeer_b200/optim.py (struct format "PPPPPq")
  void* param;          // model parameter (param_dtype)
  const void* grad;     // its gradient (param_dtype)
  float* master;        // fp32 master copy
  float* m;
  float* v;
  long long n;          // elements in this chunk
};

constexpr int kT = 256;

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return (float)*p; }

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[kT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kT / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}

// partial[chunk] = sum g^2 (unscaled); deterministic: a fixed tree per chunk, the chunk sums are added in order later
template <typename T>
__global__ void __launch_bounds__(kT) sqnorm_kernel(const Chunk* __restrict__ table, float* __restrict__ partial) {
  pdl_sync();
  const Chunk c = table[blockIdx.x];
  const T* g = reinterpret_cast<const T*>(c.grad);
  float s = 0.f;
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  const long long nv = vec ? c.n / 8 : 0;
  if (sizeof(T) == 2) {
    for (long long i = threadIdx.x; i < nv; i += kT) {
      const uint4 u = reinterpret_cast<const uint4*>(g)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); s = fmaf(f.x, f.x, s); s = fmaf(f.y, f.y, s); }
    }
    for (long long i = nv * 8 + threadIdx.x; i < c.n; i += kT) { const float f = ldf(g + i); s = fmaf(f, f, s); }
  } else {
    for (long long i = threadIdx.x; i < c.n; i += kT) { const float f = ldf(g + i); s = fmaf(f, f, s); }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, step_size, grad_scale, clip_norm;
  int n_chunks;
};

template <typename T>
__global__ void __launch_bounds__(kT) adam_kernel(const Chunk* __restrict__ table, const float* __restrict__ partial,
                                                  float* __restrict__ grad_norm_out, AdamArgs a) {
  pdl_sync();
  // every CTA re-derives the global norm from the per-chunk partials (a few thousand floats out of L2): same order in
  // every CTA and every run, no extra launch, no host round trip
  float s = 0.f;
  for (int i = threadIdx.x; i < a.n_chunks; i += kT) s += partial[i];
  const float norm = sqrtf(block_sum(s)) * a.grad_scale;
  if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
  float coef = a.grad_scale;
  if (a.clip_norm > 0.f) coef *= fminf(1.f, a.clip_norm / (norm + 1e-6f));
  const Chunk c = table[blockIdx.x];
  T* p = reinterpret_cast<T*>(c.param);
  const T* g = reinterpret_cast<const T*>(c.grad);
  const float decay = 1.f - a.weight_decay * a.lr;
  for (long long i = threadIdx.x; i < c.n; i += kT) {
    const float gr = ldf(g + i) * coef;
    const float m = a.beta1 * c.m[i] + (1.f - a.beta1) * gr;
    const float v = a.beta2 * c.v[i] + (1.f - a.beta2) * gr * gr;
    const float w = c.master[i] * decay - a.step_size * m / (sqrtf(v) + a.eps);
    c.m[i] = m;
    c.v[i] = v;
    c.master[i] = w;
    p[i] = (T)w;
  }
}

}  // namespace

// see include/ofa_b200.h
extern "C" int ofa_adam_step(const void* chunk_table, int n_chunks, float* partial_sqnorm, float* grad_norm_out, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             float clip_norm, int param_dtype, void* stream) {
  OFA_CHECK(chunk_table && partial_sqnorm && n_chunks > 0 && step >= 1, "ofa_adam_step: bad arguments (n_chunks=%d step=%d)",
            n_chunks, step);
  cudaStream_t st = (cudaStream_t)stream;
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr * sqrt(bc2) / bc1);
  a.grad_scale = grad_scale; a.clip_norm = clip_norm; a.n_chunks = n_chunks;
  const Chunk* t = reinterpret_cast<const Chunk*>(chunk_table);
  if (param_dtype == OFA_BF16) {
    OFA_CUDA(ofa_launch_pdl(sqnorm_kernel<__nv_bfloat16>, n_chunks, kT, 0, st, t, partial_sqnorm));
    OFA_CUDA(ofa_launch_pdl(adam_kernel<__nv_bfloat16>, n_chunks, kT, 0, st, t, partial_sqnorm, grad_norm_out, a));
  } else if (param_dtype == OFA_F32) {
    OFA_CUDA(ofa_launch_pdl(sqnorm_kernel<float>, n_chunks, kT, 0, st, t, partial_sqnorm));
    OFA_CUDA(ofa_launch_pdl(adam_kernel<float>, n_chunks, kT, 0, st, t, partial_sqnorm, grad_norm_out, a));
  } else {
    return ofa_set_error("ofa_adam_step: bad dtype %d", param_dtype);
  }
  OFA_LAUNCH_CHECK("adam_kernel");
  return 0;
}
