// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM), operands staged by TMA
// into 128B-swizzled shared memory, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer, warps2-5 = epilogue.
//
//   D[b][m][n] = epi( sum_k A[b](m,k) * B[b](n,k) )          epi(v) = act((v + bias[n]) * alpha) + resid[m][n]
//
// Operand storage ("major"):  K-major  = row-major [rows=M|N][K]   (Linear forward: X[M,K], W[N,K])
//                             MN-major = row-major [K][rows=M|N]   (dgrad: W as [N'=K][..]; wgrad: dY^T, X^T views)
// so forward / dgrad / wgrad of every nn.Linear on the OFA path (models/ofa/unify_multihead_attention.py:213-232,399,
// unify_transformer_layer.py:280-284,557-561, unify_transformer.py:739,906-911,1303-1316,1577-1583) run through this one
// kernel without transposed copies.  fp32 "parity mode" feeds it 3-way bf16 splits concatenated along K (ops.py).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 3;
constexpr int kStageBytes = (BM * BK + BN * BK) * 2;  // 32 KiB
constexpr int kTmemCols = 128;
constexpr int kThreads = 192;

struct GemmParams {
  void* D;
  const void* bias;    // [N] (OutT) or null
  const void* resid;   // [M, ldr] (OutT) or null
  long long ldd, ldr;  // elements
  long long batch_stride_d, batch_stride_r;
  int M, N, K;
  float alpha;
  int act;  // 0 none, 1 gelu(erf)
};

struct SmemLayout {
  // tiles first (1024B aligned for SWIZZLE_128B), then barriers
  uint8_t tiles[kStages][kStageBytes];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint32_t tmem_addr;
};

template <typename OutT>
__device__ __forceinline__ float ld_as_float(const OutT* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <int A_MN, int B_MN, typename OutT>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, bz = blockIdx.z;
  const int nkb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(&sm.tmem_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sm.tmem_addr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages, it = kb / kStages;
        mbar_wait(&sm.empty[s], (it & 1) ^ 1);
        mbar_expect_tx(&sm.full[s], kStageBytes);
        uint8_t* sa = sm.tiles[s];
        uint8_t* sb = sa + BM * BK * 2;
        if (A_MN) {
          tma_load_3d(sa, &tmA, &sm.full[s], m0, kb * BK, bz);
          tma_load_3d(sa + BK * 128, &tmA, &sm.full[s], m0 + 64, kb * BK, bz);
        } else {
          tma_load_3d(sa, &tmA, &sm.full[s], kb * BK, m0, bz);
        }
        if (B_MN) {
          tma_load_3d(sb, &tmB, &sm.full[s], n0, kb * BK, bz);
          tma_load_3d(sb + BK * 128, &tmB, &sm.full[s], n0 + 64, kb * BK, bz);
        } else {
          tma_load_3d(sb, &tmB, &sm.full[s], kb * BK, n0, bz);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages, it = kb / kStages;
        mbar_wait(&sm.full[s], it & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(sm.tiles[s]);
        const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major: 16 bf16 = 32 B inside the 128B swizzle span; MN-major: 16 k-rows = 2 swizzle atoms = 2048 B
          const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
          umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&sm.empty[s]);  // frees the smem stage when these MMAs retire
      }
      umma_commit(&sm.tmem_full);
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(&sm.tmem_full, 0);
    tc_fence_after();
    OutT* D = reinterpret_cast<OutT*>(p.D) + (long long)bz * p.batch_stride_d;
    const OutT* R = p.resid ? reinterpret_cast<const OutT*>(p.resid) + (long long)bz * p.batch_stride_r : nullptr;
    const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (row < p.M && nb < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int nvalid = min(32, p.N - nb);
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += ld_as_float<OutT>(bias + nb + j);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (R) {
          const OutT* rr = R + (long long)row * p.ldr + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += ld_as_float<OutT>(rr + j);
        }
        OutT* dp = D + (long long)row * p.ldd + nb;
        const bool vec_ok = (nvalid == 32) && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0);
        if (sizeof(OutT) == 2) {
          if (vec_ok) {
            uint4* d4 = reinterpret_cast<uint4*>(dp);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d4[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dp);
            for (int j = 0; j < nvalid; ++j) d[j] = __float2bfloat16(v[j]);
          }
        } else {
          if (vec_ok) {
            float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
            for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            float* d = reinterpret_cast<float*>(dp);
            for (int j = 0; j < nvalid; ++j) d[j] = v[j];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_d);
  }
}

template <int A_MN, int B_MN, typename OutT>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int batch, cudaStream_t st) {
  auto kern = gemm_tc_kernel<A_MN, B_MN, OutT>;
  static bool configured = false;  // per template instantiation
  const int smem = (int)sizeof(SmemLayout) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch);
  kern<<<grid, kThreads, smem, st>>>(ta, tb, p);
  OFA_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

}  // namespace

// see include/ofa_b200.h
extern "C" int ofa_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int batch, long long lda,
                             long long ldb, long long ldd, long long stride_a, long long stride_b, long long stride_d,
                             int a_mn_major, int b_mn_major, int out_dtype, const void* bias, float alpha, int act,
                             const void* resid, long long ldr, long long stride_r, void* stream) {
  OFA_CHECK(M > 0 && N > 0 && K > 0 && batch > 0, "ofa_gemm_bf16: empty problem M=%d N=%d K=%d batch=%d", M, N, K, batch);
  OFA_CHECK(lda % 8 == 0 && ldb % 8 == 0, "ofa_gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16B stride)");
  OFA_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "ofa_gemm_bf16: A/B must be 16B aligned");
  OFA_CHECK(stride_a % 8 == 0 && stride_b % 8 == 0, "ofa_gemm_bf16: batch strides must be multiples of 8 elements");
  CUtensorMap ta, tb;
  {
    // K-major: dims {K, rows, batch}, box {64, 128, 1};  MN-major: dims {rows, K, batch}, box {64, 64, 1}
    uint64_t dims[3], strides[2];
    uint32_t box[3];
    if (a_mn_major) { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = BM; }
    dims[2] = batch; box[2] = 1;
    strides[0] = (uint64_t)lda * 2;
    strides[1] = (uint64_t)(batch > 1 ? stride_a : (long long)dims[1] * lda) * 2;
    if (int e = ofa_make_tmap(&ta, A, 3, dims, strides, box, 1, 2)) return e;
    if (b_mn_major) { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = BN; }
    strides[0] = (uint64_t)ldb * 2;
    strides[1] = (uint64_t)(batch > 1 ? stride_b : (long long)dims[1] * ldb) * 2;
    if (int e = ofa_make_tmap(&tb, B, 3, dims, strides, box, 1, 2)) return e;
  }
  GemmParams p;
  p.D = D; p.bias = bias; p.resid = resid; p.ldd = ldd; p.ldr = ldr;
  p.batch_stride_d = stride_d; p.batch_stride_r = stride_r;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  const int sel = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  if (out_dtype == OFA_BF16) {
    switch (sel) {
      case 0: return launch<0, 0, __nv_bfloat16>(ta, tb, p, batch, st);
      case 1: return launch<0, 1, __nv_bfloat16>(ta, tb, p, batch, st);
      case 2: return launch<1, 0, __nv_bfloat16>(ta, tb, p, batch, st);
      default: return launch<1, 1, __nv_bfloat16>(ta, tb, p, batch, st);
    }
  } else if (out_dtype == OFA_F32) {
    switch (sel) {
      case 0: return launch<0, 0, float>(ta, tb, p, batch, st);
      case 1: return launch<0, 1, float>(ta, tb, p, batch, st);
      case 2: return launch<1, 0, float>(ta, tb, p, batch, st);
      default: return launch<1, 1, float>(ta, tb, p, batch, st);
    }
  }
  return ofa_set_error("ofa_gemm_bf16: bad out_dtype %d", out_dtype);
}
