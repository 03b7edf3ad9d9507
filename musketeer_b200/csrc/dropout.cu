// Fused dropout + drop-path + residual add (forward and backward share one kernel: the Bernoulli mask is regenerated
// from the Philox4x32-10 counter, nothing is stored).
//   y[r, :] = resid[r, :] + x[r, :] * mask(seed, element) / (1 - p) * row_scale[r / rows_per_sample]
// restates FairseqDropout (un-vendored; F.dropout semantics) + drop_path (models/ofa/unify_transformer_layer.py:19-35)
// + residual_connection (:197-198,429-430) as used at :272-273,285-290,515-516,548-549,562-567 and the embedding
// dropouts of unify_transformer.py:733,747,1495.  `row_scale` is the per-sample drop-path factor floor(keep + U)/keep
// (or null); `seed` is a device pointer so the op is CUDA-graph capturable.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_residual_kernel(const T* __restrict__ x, const T* __restrict__ resid,
                                                               T* __restrict__ y, long long n, int C, int rows_per_sample,
                                                               float p, const float* __restrict__ row_scale,
                                                               const unsigned long long* __restrict__ seed) {
  pdl_sync();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
  const long long i0 = v * 4;
  if (i0 >= n) return;
  const unsigned long long s = *seed;
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)(v >> 32), 0u, 0u),
                                  make_uint2((uint32_t)s, (uint32_t)(s >> 32)));
  const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
  const float inv_keep = p < 1.f ? 1.f / (1.f - p) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = i0 + j;
    if (i < n) {
      const float u = (float)(rr[j] >> 8) * (1.0f / 16777216.0f);
      float f = (u >= p) ? inv_keep : 0.f;
      if (row_scale) f *= row_scale[(i / C) / rows_per_sample];
      float o = (float)x[i] * f;
      if (resid) o += (float)resid[i];
      y[i] = (T)o;
    }
  }
}

}  // namespace

extern "C" int ofa_dropout_residual(const void* x, const void* resid, void* y, long long n, int C, int rows_per_sample,
                                    float p, const float* row_scale, const unsigned long long* seed, int dtype,
                                    void* stream) {
  OFA_CHECK(n > 0 && C > 0 && rows_per_sample > 0 && p >= 0.f && p < 1.f && seed, "ofa_dropout_residual: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(((n + 3) / 4 + 255) / 256);
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(dropout_residual_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)resid, (__nv_bfloat16*)y, n, C, rows_per_sample, p, row_scale, seed));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(dropout_residual_kernel<float>, grid, 256, 0, st, (const float*)x, (const float*)resid, (float*)y, n, C, rows_per_sample, p, row_scale, seed));
  else
    return ofa_set_error("ofa_dropout_residual: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("dropout_residual_kernel");
  return 0;
}
