// Patch matrix (im2col) and its adjoint for NHWC activations, fp32 and bf16: the convolutions that are not covered by the
// implicit-GEMM kernel of conv.cu -- the two stride-2 3x3 convolutions of the ResNet stem (models/ofa/resnet.py:34-37,
// 107-121: conv2 of the first bottleneck of layer2 / layer3) and every k > 1 convolution of the fp32 parity mode (3x3 of all
// bottlenecks, the 7x7 / stride 2 stem convolution :176,214) -- become `patch matrix x weight^T` on the tcgen05 GEMM
// (ofa_gemm_bf16; fp32 operands through the three-way bf16 split), forward, dgrad and wgrad.  No library convolution is
// left on the path.  Also the fp32 / bf16 generic 3x3 / stride 2 / padding 1 max-pool of the parity mode.
//   col[((n*OH + oh)*OW + ow)][(kh*KW + kw)*C + c] = x[n][oh*s - p + kh][ow*s - p + kw][c]   (0 outside the image)
//   dx[n][h][w][c] = sum over (kh, kw) with (h + p - kh) = oh*s, (w + p - kw) = ow*s of dcol[(n, oh, ow)][(kh, kw, c)]
// Pure data movement (the adjoint adds at most KH*KW terms per element, in a fixed order: no atomics).
#include "common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C,
                                                     int KH, int KW, int S, int P, int OH, int OW, long long ldcol) {
  pdl_sync();
  const long long total = (long long)N * OH * OW * KH * KW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int kw = (int)(r % KW); r /= KW;
    const int kh = (int)(r % KH); r /= KH;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    const int h = oh * S - P + kh, w = ow * S - P + kw;
    T v = T();
    if (h >= 0 && h < H && w >= 0 && w < W) v = x[(((long long)n * H + h) * W + w) * C + c];
    col[(((long long)n * OH + oh) * OW + ow) * ldcol + (long long)(kh * KW + kw) * C + c] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int N, int H, int W, int C,
                                                     int KH, int KW, int S, int P, int OH, int OW, long long ldcol) {
  pdl_sync();
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float acc = 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int hh = h + P - kh;
      if (hh < 0 || hh % S != 0 || hh / S >= OH) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int ww = w + P - kw;
        if (ww < 0 || ww % S != 0 || ww / S >= OW) continue;
        acc += (float)dcol[(((long long)n * OH + hh / S) * OW + ww / S) * ldcol + (long long)(kh * KW + kw) * C + c];
      }
    }
    dx[i] = (T)acc;
  }
}

// bf16, C % 8 == 0: eight channels per thread (16-byte loads of the <= 4 taps that reach an input pixel at stride 2, fp32 sums,
// one 16-byte store): the scalar kernel above spent 975 us on the 96x96x128 gradient of layer2.0.conv2 (per-element index
// arithmetic and 2-byte accesses), this one is bound by the 2.3 x |dx| bytes it moves.
__global__ void __launch_bounds__(256) col2im_vec8_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx,
                                                          int N, int H, int W, int C, int KH, int KW, int S, int P, int OH, int OW,
                                                          long long ldcol) {
  pdl_sync();
  const int c8 = C / 8;
  const long long total = (long long)N * H * W * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    long long r = i / c8;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int hh = h + P - kh;
      if (hh < 0 || hh % S != 0 || hh / S >= OH) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int ww = w + P - kw;
        if (ww < 0 || ww % S != 0 || ww / S >= OW) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(
            dcol + (((long long)n * OH + hh / S) * OW + ww / S) * ldcol + (long long)(kh * KW + kw) * C + c);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

// generic max-pool 3x3 / stride 2 / padding 1 (nn.MaxPool2d(3, 2, 1): first maximum in window order wins, as ATen)
template <typename T>
__global__ void __launch_bounds__(256) maxpool_any_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                              unsigned char* __restrict__ idx, int N, int H, int W, int C, int OH,
                                                              int OW) {
  pdl_sync();
  const long long total = (long long)N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    float best = -INFINITY;
    int bi = 0;
    for (int k = 0; k < 9; ++k) {
      const int h = oh * 2 - 1 + k / 3, w = ow * 2 - 1 + k % 3;
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      const float v = (float)x[(((long long)n * H + h) * W + w) * C + c];
      if (v > best || (v != v && best == best)) { best = v; bi = k; }     // NaN propagates like ATen's max-pool
    }
    y[i] = (T)best;
    idx[i] = (unsigned char)bi;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_any_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                              T* __restrict__ dx, int N, int H, int W, int C, int OH, int OW) {
  pdl_sync();
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float acc = 0.f;
    for (int k = 0; k < 9; ++k) {              // output windows (oh, ow) that contain (h, w) at window position k
      const int hh = h + 1 - k / 3, ww = w + 1 - k % 3;
      if (hh < 0 || (hh & 1) || hh / 2 >= OH || ww < 0 || (ww & 1) || ww / 2 >= OW) continue;
      const long long o = (((long long)n * OH + hh / 2) * OW + ww / 2) * C + c;
      if (idx[o] == k) acc += (float)dy[o];
    }
    dx[i] = (T)acc;
  }
}

unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  return (unsigned)(b > 148LL * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int ofa_im2col(const void* x, void* col, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                          long long ldcol, int dtype, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && ldcol >= (long long)KH * KW * C,
            "ofa_im2col: bad arguments");
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  OFA_CHECK(OH > 0 && OW > 0, "ofa_im2col: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  const int es = dtype == OFA_F32 ? 4 : 2;
  if ((C * es) % 16 == 0 && (ldcol * es) % 16 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0) {
    // pure data movement: 16-byte vectors of channels
    const int CV = C * es / 16;
    const long long totv = (long long)N * OH * OW * KH * KW * CV;
    OFA_CUDA(ofa_launch_pdl(im2col_kernel<uint4>, grid_for(totv), 256, 0, st, (const uint4*)x, (uint4*)col, N, H, W, CV, KH, KW, stride, pad, OH, OW, ldcol * es / 16));
    OFA_LAUNCH_CHECK("im2col_kernel");
    return 0;
  }
  const long long total = (long long)N * OH * OW * KH * KW * C;
  if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(im2col_kernel<float>, grid_for(total), 256, 0, st, (const float*)x, (float*)col, N, H, W, C, KH, KW, stride, pad, OH, OW, ldcol));
  else
    OFA_CUDA(ofa_launch_pdl(im2col_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, N, H, W, C, KH, KW, stride, pad, OH, OW, ldcol));
  OFA_LAUNCH_CHECK("im2col_kernel");
  return 0;
}

extern "C" int ofa_col2im(const void* dcol, void* dx, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                          long long ldcol, int dtype, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && ldcol >= (long long)KH * KW * C,
            "ofa_col2im: bad arguments");
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  OFA_CHECK(OH > 0 && OW > 0, "ofa_col2im: empty output");
  const long long total = (long long)N * H * W * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(col2im_kernel<float>, grid_for(total), 256, 0, st, (const float*)dcol, (float*)dx, N, H, W, C, KH, KW, stride, pad, OH, OW, ldcol));
  else if (C % 8 == 0 && ldcol % 8 == 0 && (((uintptr_t)dcol | (uintptr_t)dx) & 15) == 0)
    OFA_CUDA(ofa_launch_pdl(col2im_vec8_kernel, grid_for(total / 8), 256, 0, st, (const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, N, H, W, C, KH, KW, stride, pad, OH, OW, ldcol));
  else
    OFA_CUDA(ofa_launch_pdl(col2im_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, N, H, W, C, KH, KW, stride, pad, OH, OW, ldcol));
  OFA_LAUNCH_CHECK("col2im_kernel");
  return 0;
}

extern "C" int ofa_maxpool3x3s2_any(const void* in, unsigned char* idx, void* out, int N, int H, int W, int C, int backward,
                                    int dtype, void* stream) {
  OFA_CHECK(N > 0 && H > 0 && W > 0 && C > 0, "ofa_maxpool3x3s2_any: bad shape");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * (backward ? (long long)H * W : (long long)OH * OW) * C;
  if (dtype == OFA_F32) {
    if (backward) OFA_CUDA(ofa_launch_pdl(maxpool_any_bwd_kernel<float>, grid_for(total), 256, 0, st, (const float*)in, (const unsigned char*)idx, (float*)out, N, H, W, C, OH, OW));
    else OFA_CUDA(ofa_launch_pdl(maxpool_any_fwd_kernel<float>, grid_for(total), 256, 0, st, (const float*)in, (float*)out, idx, N, H, W, C, OH, OW));
  } else {
    if (backward) OFA_CUDA(ofa_launch_pdl(maxpool_any_bwd_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)in, (const unsigned char*)idx, (__nv_bfloat16*)out, N, H, W, C, OH, OW));
    else OFA_CUDA(ofa_launch_pdl(maxpool_any_fwd_kernel<__nv_bfloat16>, grid_for(total), 256, 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, idx, N, H, W, C, OH, OW));
  }
  OFA_LAUNCH_CHECK("maxpool_any_kernel");
  return 0;
}
