// 3x3 (stride 1, padding 1) convolution of the ResNet patch embedder as an implicit GEMM on the tcgen05 tensor cores:
// forward, data gradient and weight gradient, bf16 NHWC activations, [Cout][3][3][Cin] (channels-last) weights.
//   replaces conv2 of every stride-1 bottleneck, models/ofa/resnet.py:107-108,119-121 (27 of the 30 3x3 convolutions of
//   ResNet-101's three stages; the two stride-2 ones, the 7x7 stem and the max-pool stay on library kernels).
// No im2col buffer exists anywhere: the A operand of filter tap (kh, kw) is the activation tensor itself, fetched by a 4-D
// TMA box {64 channels, 8, 8, 2 images} whose start coordinate is shifted by (kw-1, kh-1); the out-of-image halo is the
// TMA unit's zero fill.  One CTA computes a 128-pixel x BN-channel output tile (128 pixels = 8 x 8 x 2 images) over
// K = 9 taps x Cin, with the same producer / single-thread MMA / 4-warp epilogue structure and TMEM double buffering as
// gemm.cu; the output tile leaves through shared memory and a 4-D TMA store (which also clips partial tiles).
//   fprop : y[n,h,w,co]  = sum_{kh,kw,ci} x[n,h+kh-1,w+kw-1,ci] W[co,kh,kw,ci]      B = W tap slice, K-major
//   dgrad : dx[n,h,w,ci] = sum_{kh,kw,co} dy[n,h-kh+1,w-kw+1,co] W[co,kh,kw,ci]     B = the same bytes, MN-major
//   wgrad : dW[co,kh,kw,ci] = sum_{n,h,w} dy[n,h,w,co] x[n,h+kh-1,w+kw-1,ci]        A = dy^T, B = shifted x, both MN-major,
//           contraction over pixels in 8 x 8 blocks, split along K into an fp32 workspace + a small reduce kernel.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;
constexpr int kNumSMs = 148;
constexpr int TW = 8, TH = 8, TN = 2;   // spatial tile of the 128 output pixels

struct ConvParams {
  int NI, H, W;            // images, height, width (input == output size)
  int Cin, Cout;           // fprop naming; dgrad swaps the roles of the GEMM N / K channel axes
  int tiles_w, tiles_h, tiles_i, tiles_n, total;
  int dgrad;
  // wgrad
  float* ws;
  int kb_total, kb_per_split, splits, tiles_m;
  int reduce_f32;          // wgrad: add the tile into an fp32 [Cout][9][Cin] buffer by TMA reduce instead of writing partials
};

template <int BN>
struct Smem {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kStageBytes = (BM * BK + BN * BK) * 2;
  uint8_t tiles[kStages][kStageBytes];
  uint8_t stage_out[4][2][4096];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_addr;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct Tile {
  int n0, h0, w0, c0;   // first image / row / column of the pixel tile, first output channel
};
__device__ __forceinline__ Tile decode(int w, const ConvParams& p, int BN) {
  Tile t;
  const int nt = w % p.tiles_n;   // channel tiles fastest: neighbours share the activation tile through L2
  w /= p.tiles_n;
  t.w0 = (w % p.tiles_w) * TW;
  w /= p.tiles_w;
  t.h0 = (w % p.tiles_h) * TH;
  t.n0 = (w / p.tiles_h) * TN;
  t.c0 = nt * BN;
  return t;
}

// B_MN = 0: fprop, 1: dgrad.  tmA: activations {C, W, H, N} box {64, 8, 8, 2}; tmB: weights {Cin, 9, Cout} box {64, 1, 64};
// tmD: output {C, W, H, N} box {64, 8, 4, 1} (one epilogue warp's 32 pixels).
template <int BN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmD, ConvParams p) {
  using S = Smem<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kc = (B_MN ? p.Cout : p.Cin) / BK;   // 64-channel blocks of the contraction per filter tap
  const int nkb = 9 * kc;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmD);
    for (int s = 0; s < S::kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(&sm.tmem_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // on-chip set-up done; from here on the kernel touches global memory
  const uint32_t tmem_base = sm.tmem_addr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        const Tile t = decode(w, p, BN);
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / kc, cb = kb - tap * kc;
          const int kh = tap / 3, kw = tap - kh * 3;
          const int dh = B_MN ? 1 - kh : kh - 1, dw = B_MN ? 1 - kw : kw - 1;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], S::kStageBytes);
          uint8_t* sa = sm.tiles[s];
          uint8_t* sb = sa + BM * BK * 2;
          tma_load_4d(sa, &tmA, &sm.full[s], cb * BK, t.w0 + dw, t.h0 + dh, t.n0);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) {
            if (B_MN) tma_load_3d(sb + c * BK * 128, &tmB, &sm.full[s], t.c0 + 64 * c, tap, cb * BK);   // [64 co][64 ci]
            else tma_load_3d(sb + c * 64 * 128, &tmB, &sm.full[s], cb * BK, tap, t.c0 + 64 * c);        // [64 co][64 ci]
          }
          if (++s == S::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0, ph = 0, it = 0;
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, B_MN);
      for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&sm.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(sm.tiles[s]);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&sm.empty[s]);
          if (++s == S::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;   // TMEM lanes / tile rows [32q, 32q+32): image n0 + (q >> 1), rows h0 + 4*(q & 1) .. +3
    int it = 0, sbuf = 0;
    const int nout = B_MN ? p.Cin : p.Cout;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const Tile t = decode(w, p, BN);
      const int acc = it & 1;
      mbar_wait(&sm.tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const int img = t.n0 + (q >> 1), hh = t.h0 + 4 * (q & 1);
      if (img < p.NI && hh < p.H) {
#pragma unroll 1
        for (int c2 = 0; c2 < BN / 32; c2 += 2) {
          if (t.c0 + c2 * 32 >= nout) break;
          uint8_t* sb = sm.stage_out[q][sbuf];
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32(tmem_row + (c2 + half) * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int chunk = half * 4 + q4;
              *reinterpret_cast<uint4*>(sb + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = make_uint4(
                  pack_bf16(__uint_as_float(r[8 * q4]), __uint_as_float(r[8 * q4 + 1])),
                  pack_bf16(__uint_as_float(r[8 * q4 + 2]), __uint_as_float(r[8 * q4 + 3])),
                  pack_bf16(__uint_as_float(r[8 * q4 + 4]), __uint_as_float(r[8 * q4 + 5])),
                  pack_bf16(__uint_as_float(r[8 * q4 + 6]), __uint_as_float(r[8 * q4 + 7])));
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmD, sb, t.c0 + c2 * 32, t.w0, hh, img);
            tma_store_commit();
          }
          sbuf ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
    }
    if (lane == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// weight gradient: work item = (tap, 128-row block of Cout, BN-column block of Cin, K slice); K-blocks are 8 x 8 pixel
// blocks of one image.  tmDY / tmX: {C, W, H, N} box {64, 8, 8, 1}.  Partials: ws[split][9][Cout][Cin] fp32.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                    const __grid_constant__ CUtensorMap tmX,
                                                                    const __grid_constant__ CUtensorMap tmDW, ConvParams p) {
  using S = Smem<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmX);
    for (int s = 0; s < S::kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(&sm.tmem_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // on-chip set-up done; from here on the kernel touches global memory
  const uint32_t tmem_base = sm.tmem_addr;
  const int per_img = p.tiles_w * p.tiles_h;

  // work decode: split fastest, then n-tile, m-tile, tap
  auto dec = [&](int w, int& tap, int& m0, int& n0, int& sp) {
    sp = w % p.splits; w /= p.splits;
    n0 = (w % p.tiles_n) * BN; w /= p.tiles_n;
    m0 = (w % p.tiles_m) * BM;
    tap = w / p.tiles_m;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int w = blockIdx.x; w < p.total; w += gridDim.x) {
        int tap, m0, n0, sp;
        dec(w, tap, m0, n0, sp);
        const int kh = tap / 3, kw = tap - kh * 3;
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int img = kb / per_img, rem = kb - img * per_img;
          const int h0 = (rem / p.tiles_w) * TH, w0 = (rem % p.tiles_w) * TW;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], S::kStageBytes);
          uint8_t* sa = sm.tiles[s];
          uint8_t* sb = sa + BM * BK * 2;
          tma_load_4d(sa, &tmDY, &sm.full[s], m0, w0, h0, img);                  // [64 pixels][64 co]  (MN-major A)
          tma_load_4d(sa + BK * 128, &tmDY, &sm.full[s], m0 + 64, w0, h0, img);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_4d(sb + c * BK * 128, &tmX, &sm.full[s], n0 + 64 * c, w0 + kw - 1, h0 + kh - 1, img);   // [64 px][64 ci]
          if (++s == S::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0, ph = 0, it = 0;
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 1, 1);
      for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
        int tap, m0, n0, sp;
        dec(w, tap, m0, n0, sp);
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int acc = it & 1;
        mbar_wait(&sm.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(sm.tiles[s]);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(tmem_d, umma_smem_desc(sa + k * 2048, BK * 128, 1024), umma_smem_desc(sb + k * 2048, BK * 128, 1024),
                     idesc, (kb != kb0) | (k != 0));
          umma_commit(&sm.empty[s]);
          if (++s == S::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    int it = 0, sbuf = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      int tap, m0, n0, sp;
      dec(w, tap, m0, n0, sp);
      const int acc = it & 1;
      const int row = m0 + q * 32 + lane;   // output channel
      mbar_wait(&sm.tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      if (p.reduce_f32) {
        if (m0 + q * 32 < p.Cout) {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            if (n0 + c * 32 >= p.Cin) break;
            uint8_t* sb = sm.stage_out[q][sbuf];
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            uint32_t r[32];
            tmem_ld32(tmem_row + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              *reinterpret_cast<uint4*>(sb + lane * 128 + ((q4 ^ (lane & 7)) << 4)) =
                  make_uint4(r[4 * q4], r[4 * q4 + 1], r[4 * q4 + 2], r[4 * q4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(&tmDW, sb, n0 + c * 32, tap, m0 + q * 32);
              tma_store_commit();
            }
            sbuf ^= 1;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
        continue;
      }
      float* W = p.ws + (((size_t)sp * 9 + tap) * p.Cout + (row < p.Cout ? row : 0)) * p.Cin + n0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_row + c * 32, r);
        tmem_ld_wait();
        if (row < p.Cout && n0 + c * 32 < p.Cin) {
          float4* d4 = reinterpret_cast<float4*>(W + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
    }
    if (lane == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// dW[co][tap][ci] (bf16, channels-last weight layout) = (accumulate ? dW : 0) + sum_sp ws[sp][tap][co][ci]
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ dw,
                                                                int splits, int Cout, int Cin, int accumulate) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [tap][co][ci]
  const long long per = 9LL * Cout * Cin;
  if (idx >= per) return;
  const int ci = (int)(idx % Cin), co = (int)((idx / Cin) % Cout), tap = (int)(idx / ((long long)Cin * Cout));
  float v = 0.f;
  for (int sp = 0; sp < splits; ++sp) v += ws[(size_t)sp * per + idx];
  __nv_bfloat16* d = dw + ((size_t)co * 9 + tap) * Cin + ci;
  if (accumulate) v += __bfloat162float(*d);
  *d = __float2bfloat16(v);
}

int act_tmap(CUtensorMap* tm, const void* ptr, int C, int W, int H, int NI, int bw, int bh, int bn) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NI};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  return ofa_make_tmap(tm, ptr, 4, dims, strides, box, 1, 2);
}

template <int BN, int B_MN>
int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const ConvParams& p, cudaStream_t st) {
  auto kern = conv3x3_kernel<BN, B_MN>;
  static bool configured = false;
  const int smem = (int)sizeof(Smem<BN>) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  OFA_CUDA(ofa_launch_pdl(kern, p.total < kNumSMs ? p.total : kNumSMs, kThreads, smem, st, ta, tb, td, p));
  OFA_LAUNCH_CHECK("conv3x3_kernel");
  return 0;
}

template <int BN>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const CUtensorMap& tdw, const ConvParams& p, cudaStream_t st) {
  auto kern = conv3x3_wgrad_kernel<BN>;
  static bool configured = false;
  const int smem = (int)sizeof(Smem<BN>) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  OFA_CUDA(ofa_launch_pdl(kern, p.total < kNumSMs ? p.total : kNumSMs, kThreads, smem, st, tdy, tx, tdw, p));
  OFA_LAUNCH_CHECK("conv3x3_wgrad_kernel");
  return 0;
}

void wgrad_plan(int NI, int H, int W, int Cin, int Cout, int* bn, int* splits, int* kb_total, int reduce = 0) {
  *bn = Cin >= 256 ? 256 : (Cin >= 128 ? 128 : 64);
  const int tiles = 9 * ((Cout + BM - 1) / BM) * ((Cin + *bn - 1) / *bn);
  *kb_total = NI * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
  // fp32 partials are the cost of a split: about one work item per SM; two when the tiles are reduce-added in place
  int s = ((reduce ? 2 : 1) * kNumSMs + tiles - 1) / tiles;
  if (s > *kb_total / 8) s = *kb_total / 8;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  *splits = s;
}

}  // namespace

// see include/ofa_b200.h
extern "C" int ofa_conv3x3_bf16(const void* in, const void* weight, void* out, int NI, int H, int W, int Cin, int Cout,
                                int dgrad, void* stream) {
  OFA_CHECK(NI > 0 && H > 0 && W > 0, "ofa_conv3x3_bf16: empty problem");
  OFA_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "ofa_conv3x3_bf16: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
  OFA_CHECK((((uintptr_t)in | (uintptr_t)weight | (uintptr_t)out) & 15) == 0, "ofa_conv3x3_bf16: pointers must be 16B aligned");
  // fprop: in = x [N,H,W,Cin] -> out = y [N,H,W,Cout].   dgrad: in = dy [N,H,W,Cout] -> out = dx [N,H,W,Cin].
  const int c_in_act = dgrad ? Cout : Cin, c_out_act = dgrad ? Cin : Cout;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.NI = NI; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.dgrad = dgrad;
  p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + TH - 1) / TH; p.tiles_i = (NI + TN - 1) / TN;
  const int bn = c_out_act >= 256 ? 256 : (c_out_act >= 128 ? 128 : 64);   // measured: wider beats more, narrower tiles
  p.tiles_n = (c_out_act + bn - 1) / bn;
  p.total = p.tiles_w * p.tiles_h * p.tiles_i * p.tiles_n;
  CUtensorMap ta, tb, td;
  if (int e = act_tmap(&ta, in, c_in_act, W, H, NI, TW, TH, TN)) return e;
  if (int e = act_tmap(&td, out, c_out_act, W, H, NI, TW, 4, 1)) return e;
  {
    uint64_t dims[3] = {(uint64_t)Cin, 9, (uint64_t)Cout};
    uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)9 * Cin * 2};
    uint32_t box[3] = {64, 1, 64};
    if (int e = ofa_make_tmap(&tb, weight, 3, dims, strides, box, 1, 2)) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dgrad) {
    if (bn == 256) return launch_conv<256, 1>(ta, tb, td, p, st);
    if (bn == 128) return launch_conv<128, 1>(ta, tb, td, p, st);
    return launch_conv<64, 1>(ta, tb, td, p, st);
  }
  if (bn == 256) return launch_conv<256, 0>(ta, tb, td, p, st);
  if (bn == 128) return launch_conv<128, 0>(ta, tb, td, p, st);
  return launch_conv<64, 0>(ta, tb, td, p, st);
}

extern "C" long long ofa_conv3x3_wgrad_workspace_bytes(int NI, int H, int W, int Cin, int Cout) {
  int bn, splits, kbt;
  wgrad_plan(NI, H, W, Cin, Cout, &bn, &splits, &kbt);
  return (long long)splits * 9 * Cout * Cin * (long long)sizeof(float);
}

extern "C" int ofa_conv3x3_wgrad_bf16(const void* x, const void* dy, void* dw, int NI, int H, int W, int Cin, int Cout,
                                      int accumulate, void* workspace, long long workspace_bytes, void* stream) {
  OFA_CHECK(NI > 0 && H > 0 && W > 0, "ofa_conv3x3_wgrad_bf16: empty problem");
  OFA_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "ofa_conv3x3_wgrad_bf16: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.NI = NI; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  int bn, splits, kbt;
  const int reduce = accumulate == 2;   // dw is an fp32 [Cout][9][Cin] accumulation buffer
  wgrad_plan(NI, H, W, Cin, Cout, &bn, &splits, &kbt, reduce);
  OFA_CHECK(reduce || (workspace && workspace_bytes >= (long long)splits * 9 * Cout * Cin * (long long)sizeof(float)),
            "ofa_conv3x3_wgrad_bf16: workspace too small (see ofa_conv3x3_wgrad_workspace_bytes)");
  p.ws = (float*)workspace;
  p.reduce_f32 = reduce;
  p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + TH - 1) / TH;
  p.kb_total = kbt;
  p.kb_per_split = (kbt + splits - 1) / splits;
  p.splits = (kbt + p.kb_per_split - 1) / p.kb_per_split;
  p.tiles_m = (Cout + BM - 1) / BM;
  p.tiles_n = (Cin + bn - 1) / bn;
  p.total = 9 * p.tiles_m * p.tiles_n * p.splits;
  CUtensorMap tdy, tx, tdw;
  memset(&tdw, 0, sizeof(tdw));
  if (reduce) {
    uint64_t dims[3] = {(uint64_t)Cin, 9, (uint64_t)Cout};
    uint64_t strides[2] = {(uint64_t)Cin * 4, (uint64_t)9 * Cin * 4};
    uint32_t box[3] = {32, 1, 32};
    if (int e = ofa_make_tmap(&tdw, dw, 3, dims, strides, box, 1, 4)) return e;
  }
  if (int e = act_tmap(&tdy, dy, Cout, W, H, NI, TW, TH, 1)) return e;
  if (int e = act_tmap(&tx, x, Cin, W, H, NI, TW, TH, 1)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (bn == 256) rc = launch_wgrad<256>(tdy, tx, tdw, p, st);
  else if (bn == 128) rc = launch_wgrad<128>(tdy, tx, tdw, p, st);
  else rc = launch_wgrad<64>(tdy, tx, tdw, p, st);
  if (rc || reduce) return rc;
  const long long n = 9LL * Cout * Cin;
  OFA_CUDA(ofa_launch_pdl(conv_wgrad_reduce_kernel, (unsigned)((n + 255) / 256), 256, 0, st, p.ws, (__nv_bfloat16*)dw, p.splits, Cout, Cin, accumulate));
  OFA_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
  return 0;
}
