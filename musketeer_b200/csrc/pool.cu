// 3x3 / stride 2 / padding 1 max-pool of the ResNet stem on NHWC activations (models/ofa/resnet.py:179,216: nn.MaxPool2d),
// forward and backward.  HBM-bound: one thread per output pixel and 8 channels (16-byte vectors); the forward stores the
// window position of the maximum (first maximum in (kh, kw) scan order, like ATen) as one byte per element, the backward
// GATHERS: every input pixel lies in at most four windows and sums the dy of those whose recorded maximum it is -- no
// atomics, no zero-fill pass, deterministic.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          unsigned char* __restrict__ idx, int N, int H, int W, int C, int OH,
                                                          int OW) {
  pdl_sync();
  const int c8 = C / 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * OH * OW * c8) return;
  const int cg = (int)(t % c8);
  long long p = t / c8;
  const int ow = (int)(p % OW); p /= OW;
  const int oh = (int)(p % OH);
  const int n = (int)(p / OH);
  float best[8];
  int bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { best[j] = -3.4e38f; bi[j] = 4; }   // the centre tap is always inside the image
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int h = oh * 2 - 1 + kh;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int w = ow * 2 - 1 + kw;
      if (w < 0 || w >= W) continue;
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + cg * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > best[j]) { best[j] = v[j]; bi[j] = kh * 3 + kw; }
    }
  }
  const long long o = (((long long)n * OH + oh) * OW + ow) * C + cg * 8;
  *reinterpret_cast<uint4*>(y + o) = make_uint4(pack_bf16(best[0], best[1]), pack_bf16(best[2], best[3]),
                                                pack_bf16(best[4], best[5]), pack_bf16(best[6], best[7]));
  *reinterpret_cast<uint2*>(idx + o) = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                                                  bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                          const unsigned char* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                                                          int N, int H, int W, int C, int OH, int OW) {
  pdl_sync();
  const int c8 = C / 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * H * W * c8) return;
  const int cg = (int)(t % c8);
  long long p = t / c8;
  const int w = (int)(p % W); p /= W;
  const int h = (int)(p % H);
  const int n = (int)(p / H);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // windows covering input row h are oh = floor(h/2) (tap kh = h - 2*oh + 1) and, for odd h, oh = floor(h/2) + 1 (tap 0)
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    // candidate output rows: floor(h/2) and floor(h/2)+1 (the latter only covers h when h is odd)
    const int oh = h / 2 + a;
    const int kh = h - (oh * 2 - 1);
    if (oh >= OH || kh < 0 || kh > 2) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ow = w / 2 + b;
      const int kw = w - (ow * 2 - 1);
      if (ow >= OW || kw < 0 || kw > 2) continue;
      const long long o = (((long long)n * OH + oh) * OW + ow) * C + cg * 8;
      const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
      float g[8];
      unpack8(*reinterpret_cast<const uint4*>(dy + o), g);
      const int want = kh * 3 + kw;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sel = ((j < 4 ? iv.x >> (8 * j) : iv.y >> (8 * (j - 4))) & 0xff);
        if (sel == want) acc[j] += g[j];
      }
    }
  }
  *reinterpret_cast<uint4*>(dx + (((long long)n * H + h) * W + w) * C + cg * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
}

}  // namespace

// see include/ofa_b200.h
extern "C" int ofa_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* idx, int N, int H, int W, int C, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C % 8 == 0, "ofa_maxpool3x3s2_fwd: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long n = (long long)N * OH * OW * (C / 8);
  OFA_CUDA(ofa_launch_pdl(maxpool_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
                          (const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, N, H, W, C, OH, OW));
  return 0;
}

extern "C" int ofa_maxpool3x3s2_bwd(const void* dy, const unsigned char* idx, void* dx, int N, int H, int W, int C,
                                    void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C % 8 == 0, "ofa_maxpool3x3s2_bwd: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long n = (long long)N * H * W * (C / 8);
  OFA_CUDA(ofa_launch_pdl(maxpool_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
                          (const __nv_bfloat16*)dy, idx, (__nv_bfloat16*)dx, N, H, W, C, OH, OW));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// 7x7 / stride 2 / padding 3 stem convolution (models/ofa/resnet.py:176,214: conv1, 3 -> 64 channels) as patch matrix +
// GEMM: with 3 input channels there is no 64-channel K block for the implicit-GEMM kernel, so the patches are written out
// once, bf16 [N*OH*OW][152] (147 = 3*7*7 taps in (c, kh, kw) order -- the order of weight.view(64, 147) -- zero-padded to a
// 16-byte row), and both the forward and the weight gradient are plain ofa_gemm_bf16 calls on it.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kStemK = 147, kStemKp = 152;
__global__ void __launch_bounds__(256) stem_patches_kernel(const __nv_bfloat16* __restrict__ x /* NHWC, C = 3 */,
                                                           __nv_bfloat16* __restrict__ col, int N, int H, int W, int OH, int OW) {
  pdl_sync();
  // one thread per (output pixel, 8 consecutive columns): 19 threads write a 304-byte patch row with 16-byte stores; the
  // 8 taps are scalar reads of a 7 x 21-element window that stays in L1
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * OH * OW * 19) return;
  const int cg = (int)(t % 19);
  const long long p = t / 19;
  const int ow = (int)(p % OW);
  const int oh = (int)((p / OW) % OH);
  const int n = (int)(p / ((long long)OW * OH));
  const __nv_bfloat16* img = x + (long long)n * H * W * 3;
  unsigned short v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int j = cg * 8 + e;            // column (c, kh, kw); 147..151 are padding
    const int c = j / 49, rem = j - c * 49, kh = rem / 7, kw = rem - kh * 7;
    const int h = oh * 2 - 3 + kh, w = ow * 2 - 3 + kw;
    const bool in = j < kStemK && h >= 0 && h < H && w >= 0 && w < W;
    v[e] = in ? reinterpret_cast<const unsigned short*>(img)[((long long)h * W + w) * 3 + c] : (unsigned short)0;
  }
  *reinterpret_cast<uint4*>(col + p * kStemKp + cg * 8) =
      make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
}
}  // namespace

extern "C" int ofa_stem_patches(const void* x, void* col, int N, int H, int W, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0, "ofa_stem_patches: bad shape");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long n = (long long)N * OH * OW * 19;
  OFA_CUDA(ofa_launch_pdl(stem_patches_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
                          (const __nv_bfloat16*)x, (__nv_bfloat16*)col, N, H, W, OH, OW));
  return 0;
}

// ---- stride-2 pixel subsampling of an NHWC activation (the input side of the stride-2 1x1 downsample convolutions,
// models/ofa/resnet.py:196-203: conv1x1(inplanes, planes * 4, stride)) and its adjoint.  16-byte vectors; esize = bytes per
// element, C * esize % 16 == 0.   fwd: y[n, h, w, :] = x[n, 2h, 2w, :]     bwd: dx = 0 except dx[n, 2h, 2w, :] = dy[n, h, w, :]
namespace {
__global__ void __launch_bounds__(256) subsample2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H,
                                                             int W, int Ho, int Wo, int CV) {
  pdl_sync();
  const long long total = (long long)N * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long pix = i / CV;
    const int wo = (int)(pix % Wo); pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    y[i] = x[(((long long)n * H + 2 * ho) * W + 2 * wo) * CV + cv];
  }
}
__global__ void __launch_bounds__(256) subsample2_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H,
                                                             int W, int Ho, int Wo, int CV) {
  pdl_sync();
  const long long total = (long long)N * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long pix = i / CV;
    const int w = (int)(pix % W); pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (((h | w) & 1) == 0) v = dy[(((long long)n * Ho + (h >> 1)) * Wo + (w >> 1)) * CV + cv];
    dx[i] = v;
  }
}
}  // namespace

extern "C" int ofa_subsample2(const void* src, void* dst, int N, int H, int W, int C, int esize, int backward, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && (esize == 2 || esize == 4) && (C * esize) % 16 == 0,
            "ofa_subsample2: N=%d H=%d W=%d C=%d esize=%d (C * esize must be a multiple of 16)", N, H, W, C, esize);
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, CV = C * esize / 16;
  const long long total = (long long)N * (backward ? (long long)H * W : (long long)Ho * Wo) * CV;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (backward) OFA_CUDA(ofa_launch_pdl(subsample2_bwd_kernel, (unsigned)blocks, 256, 0, st, (const uint4*)src, (uint4*)dst, N, H, W, Ho, Wo, CV));
  else OFA_CUDA(ofa_launch_pdl(subsample2_fwd_kernel, (unsigned)blocks, 256, 0, st, (const uint4*)src, (uint4*)dst, N, H, W, Ho, Wo, CV));
  OFA_LAUNCH_CHECK("subsample2_kernel");
  return 0;
}

// ---- input hand-off: uint8 HWC images -> normalised NCHW activations on the device --------------------------
// The reference normalises on the host (data/mm_data/*_dataset.py: transforms.ToTensor() = x / 255, then
// transforms.Normalize(mean, std) = (x - mean) / std, all fp32) and ships 4 bytes per sample value through PCIe
// (trainer.py:1246-1284 move_to_cuda, then apply_bfloat16).  Here the loader ships the decoded uint8 pixels (1 byte) and this
// kernel applies the same three IEEE fp32 operations in the same order (bit-equal to the host result before the final cast).
// HBM-bound: 3 bytes read, 3 * sizeof(T) written per pixel; four pixels per thread (three 32-bit loads, 8-byte plane stores).
namespace {

template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const unsigned char* __restrict__ x, T* __restrict__ y, long long npix,
                                                           long long plane, float m0, float m1, float m2, float s0, float s1,
                                                           float s2) {
  pdl_sync();
  // npix = N * H * W pixels in all; plane = H * W; pixel p of image n -> y[(n * 3 + c) * plane + p]
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= npix) return;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  unsigned char b[12];
  const bool full = q + 4 <= npix && (plane & 3) == 0;      // four pixels of one image, 12 bytes at a 4-byte aligned address
  if (full) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(x + q * 3);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = (w0 >> (8 * i)) & 255; b[4 + i] = (w1 >> (8 * i)) & 255; b[8 + i] = (w2 >> (8 * i)) & 255; }
  } else {
    for (int i = 0; i < 12; ++i) b[i] = q * 3 + i < npix * 3 ? x[q * 3 + i] : 0;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    T o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = cvt_out<T>(__fdiv_rn(__fsub_rn(__fdiv_rn((float)b[i * 3 + c], 255.f), mean[c]), sd[c]));
    if (full) {
      const long long n = q / plane, p = q - n * plane;
      T* dst = y + (n * 3 + c) * plane + p;
      if constexpr (sizeof(T) == 2) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(o);
      else *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
    } else {
      for (int i = 0; i < 4 && q + i < npix; ++i) {
        const long long n = (q + i) / plane, p = (q + i) - n * plane;
        y[(n * 3 + c) * plane + p] = o[i];
      }
    }
  }
}

}  // namespace

extern "C" int ofa_normalize_u8(const void* x, void* y, int N, int H, int W, const float* mean3, const float* std3, int dtype,
                                void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && x && y && mean3 && std3, "ofa_normalize_u8: null operand or empty problem");
  OFA_CHECK(((uintptr_t)x & 3) == 0 && ((uintptr_t)y & 15) == 0, "ofa_normalize_u8: unaligned buffers");
  const long long plane = (long long)H * W, npix = plane * N;
  const long long threads = (npix + 3) / 4;
  dim3 grid((unsigned)((threads + 255) / 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(normalize_u8_kernel<__nv_bfloat16>, grid, 256, 0, st, (const unsigned char*)x, (__nv_bfloat16*)y, npix, plane,
                            mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(normalize_u8_kernel<float>, grid, 256, 0, st, (const unsigned char*)x, (float*)y, npix, plane,
                            mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]));
  else
    return ofa_set_error("ofa_normalize_u8: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("normalize_u8_kernel");
  return 0;
}
