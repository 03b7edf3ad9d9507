// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores: persistent, warp-specialised, TMA + tcgen05 + TMEM.
//
//   D[b][m][n] = epi( sum_k A[b](m,k) * B[b](n,k) )          epi(v) = act((v + bias[n]) * alpha) + resid[m][n]
//
// One CTA per SM loops over 128 x BN output tiles (BN = 128 or 256).  warp0 = TMA producer (kStages-deep ring of
// 128B-swizzled A/B stages), warp1 = single-thread tcgen05.mma issuer, warps2-5 = epilogue.  The fp32 accumulator is
// double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i+1.  Problems with
// few tiles and a long contraction (weight gradients: K = tokens) are split along K; the slices land in an fp32
// workspace and a small kernel reduces them and applies the epilogue.
//
// Operand storage ("major"):  K-major  = row-major [rows=M|N][K]   (Linear forward: X[M,K], W[N,K])
//                             MN-major = row-major [K][rows=M|N]   (dgrad: W as [N'=K][..]; wgrad: dY^T, X^T views)
// so forward / dgrad / wgrad of every nn.Linear on the OFA path (models/ofa/unify_multihead_attention.py:213-232,399,
// unify_transformer_layer.py:280-284,557-561, unify_transformer.py:739,906-911,1303-1316,1577-1583) run through this one
// kernel without transposed copies.  fp32 "parity mode" feeds it 3-way bf16 splits concatenated along K (ops.py).
#include <string.h>

#include "common.cuh"

// a 256 x 256 pair tile costs each SM about what one 128 x 128 tile costs (measured ~1.05 vs ~0.55 PFLOP/s): the pair kernel
// wins as soon as the single-CTA kernel would need a second wave, i.e. from about a quarter of the machine's pairs
static int g_ofa_gemm_pair_min_tiles = 38;
static int g_ofa_gemm_small64 = 1;        // 128 x 64 tiles for small-M forward / dgrad problems (A/B switch)
static int g_ofa_gemm_wgrad_bn256 = 1;    // weight-gradient (fp32 accumulate) problems prefer 128 x 256 tiles
static int g_ofa_gemm_tma_store = 1;      // bf16 epilogue through shared memory + TMA store (0: per-thread row stores)
static int g_ofa_gemm_pair_enabled = 2;   // 0: single-CTA tiles, 1: pair with B multicast, 2: cta_group::2 MMA

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;
constexpr int kNumSMs = 148;

struct GemmParams {
  void* D;
  const void* bias;    // [N] (OutT) or null
  const void* resid;   // [M, ldr] (OutT) or null
  float* ws;           // split-K workspace [splits][batch][M][N] fp32 (splits > 1)
  long long ldd, ldr;  // elements
  long long batch_stride_d, batch_stride_r;
  int M, N, K;
  int tiles_m, tiles_n, batch, splits, kb_per_split;
  int n_big;     // work items [0, n_big) are full 128 x BN tiles; the rest are 128 x 64 sub-tiles of the last tiles
  int total;     // total work items
  float alpha;
  int alpha_cols;  // alpha applies to output columns < alpha_cols only (fused q|k|v projection: q is pre-scaled); 0 = all
  int act;  // 0 none, 1 gelu(erf)
  float* rowsum;  // fp32 [M]: rowsum[m] += alpha * sum_k A(m,k) (bias gradient of a weight-gradient GEMM) or null
  int reduce_f32; // fp32 output ADDED into D by TMA reduce (gradient accumulation; K slices need no workspace)
  int tma_store;  // bf16 output staged through shared memory and written with TMA (needs 16B-aligned D rows)
};

template <int BN>
struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kStageBytes = (BM * BK + BN * BK) * 2;
  static constexpr int kTmemCols = 2 * BN;
};

template <int BN>
struct SmemLayout {
  uint8_t tiles[Cfg<BN>::kStages][Cfg<BN>::kStageBytes];  // 1024B aligned (SWIZZLE_128B)
  uint8_t stage_out[4][2][4096];  // per epilogue warp: two 32-row x 64-column bf16 slabs (SWIZZLE_128B) for the TMA store
  uint8_t ones[1024];             // one 8-row swizzle atom of bf16 1.0 that both 8-row groups of the N = 16 operand alias (SBO = 0): B operand of the row-sum MMA (bias gradients inside the wgrad GEMM)
  uint64_t full[Cfg<BN>::kStages];
  uint64_t empty[Cfg<BN>::kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_addr;
};

template <typename OutT>
__device__ __forceinline__ float ld_as_float(const OutT* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

struct Work {
  int m0, n0, bn, bz, sp;
};
// Tail splitting: with T tiles on P persistent CTAs the last T mod P tiles would occupy a whole extra round; they are
// issued as 128 x 64 sub-tiles instead so the tail spreads over all SMs.
template <int BN, int PAIR>
__device__ __forceinline__ Work decode(int w, const GemmParams& p, int crank) {
  Work r;
  int sub = 0;
  r.bn = BN;
  constexpr int SUB = PAIR ? 128 : 64;   // a pair shares B in two halves, so its sub-tiles stay a multiple of 128 wide
  if (w >= p.n_big) {
    const int q = BN / SUB;
    const int u = w - p.n_big;
    sub = u % q;
    w = p.n_big + u / q;
    r.bn = SUB;
  }
  const int nt = w % p.tiles_n;  // n fastest: CTAs running concurrently share the A row-block through L2
  w /= p.tiles_n;
  const int mt = w % p.tiles_m;  // PAIR: tiles_m counts row-tile PAIRS
  w /= p.tiles_m;
  r.sp = w % p.splits;
  r.bz = w / p.splits;
  r.m0 = (PAIR ? 2 * mt + crank : mt) * BM;
  r.n0 = nt * BN + sub * SUB;
  return r;
}

template <typename OutT>
__device__ __forceinline__ void store_row32(OutT* dp, const float (&v)[32], int nvalid) {
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
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) d[j] = __float2bfloat16(v[j]);
    }
  } else {
    if (vec_ok) {
      float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      float* d = reinterpret_cast<float*>(dp);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) d[j] = v[j];
    }
  }
}


// Epilogue of one 32-row slab (one warp) of a bf16 tile through shared memory + TMA store: every 64-column group is
// packed into a 128B-swizzled [32][64] slab (lane = row, conflict-free 16-byte stores) and written with one
// cp.async.bulk.tensor store, which also clips rows >= M and columns >= N.  Replaces 32 scattered 64-byte row segments
// per warp instruction (2x the L2 write transactions) with full-line writes, and takes the stores off the LSU.
template <bool RESID>
__device__ __forceinline__ void epilogue_tma_impl(const CUtensorMap* tmD, uint8_t (*stage)[4096], int& sbuf, uint32_t tmem_row,
                                             int nch, int row, int n0, int bz, const GemmParams& p, int lane) {
  const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  const __nv_bfloat16* R =
      RESID && p.resid ? reinterpret_cast<const __nv_bfloat16*>(p.resid) + (long long)bz * p.batch_stride_r : nullptr;
  const int row0 = row - lane;
  // residual rows are fetched one 64-column step ahead: each lane reads its own row (32 different lines per load
  // instruction), so an inline load would expose a full memory latency per step and the epilogue of short-K GEMMs
  // (the 1x1-convolution dgrads with the identity-branch gradient as residual) would be bound by it
  const bool rfast = RESID && R != nullptr && row < p.M && ((reinterpret_cast<uintptr_t>(R) & 15) == 0) && (p.ldr % 8 == 0) && (n0 % 8 == 0);
  uint4 rn[8];
  auto fetch = [&](int c2n) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int nb = n0 + (c2n + half) * 32;
      const bool ok = c2n + half < nch && nb + 32 <= p.N;
      const uint4* src = reinterpret_cast<const uint4*>(R + (long long)row * p.ldr + (ok ? nb : 0));
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) rn[half * 4 + q4] = ok ? src[q4] : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (rfast) fetch(0);
#pragma unroll 1
  for (int c2 = 0; c2 < nch; c2 += 2) {
    const int nb0 = n0 + c2 * 32;
    if (nb0 >= p.N) break;
    uint4 rc[8];
    if (RESID) {
#pragma unroll
      for (int q = 0; q < 8; ++q) rc[q] = rn[q];
      if (rfast && c2 + 2 < nch) fetch(c2 + 2);
    }
    uint8_t* sb = stage[sbuf];
    if (lane == 0) tma_store_wait_read<1>();   // the store issued two groups ago has finished reading this slab
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (c2 + half >= nch) break;
      uint32_t r[32];
      tmem_ld32(tmem_row + (c2 + half) * 32, r);
      tmem_ld_wait();
      const int nb = nb0 + half * 32;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int nvalid = min(32, p.N - nb);
      if (bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] += __bfloat162float(bias[nb + j]);
      }
      const float al = (p.alpha_cols == 0 || nb < p.alpha_cols) ? p.alpha : 1.f;   // 32-column groups never straddle
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= al;
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      }
      if (RESID && rfast && nvalid == 32) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rc[half * 4 + q4]);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[q4 * 8 + 2 * k] += f.x; v[q4 * 8 + 2 * k + 1] += f.y; }
        }
      } else if (RESID && R && row < p.M) {
        const __nv_bfloat16* rr = R + (long long)row * p.ldr + nb;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(rr) & 15) == 0)) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint4 u = reinterpret_cast<const uint4*>(rr)[q4];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[q4 * 8 + 2 * k] += f.x; v[q4 * 8 + 2 * k + 1] += f.y; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += __bfloat162float(rr[j]);
        }
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = half * 4 + q4;   // 16-byte chunk inside the 128-byte slab row
        *reinterpret_cast<uint4*>(sb + lane * 128 + ((chunk ^ (lane & 7)) << 4)) =
            make_uint4(pack_bf16(v[8 * q4], v[8 * q4 + 1]), pack_bf16(v[8 * q4 + 2], v[8 * q4 + 3]),
                       pack_bf16(v[8 * q4 + 4], v[8 * q4 + 5]), pack_bf16(v[8 * q4 + 6], v[8 * q4 + 7]));
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmD, sb, nb0, row0, bz);
      tma_store_commit();
    }
    sbuf ^= 1;
  }
}

// the residual-free instantiation carries none of the residual prefetch state (most GEMMs of the step)
__device__ __forceinline__ void epilogue_tma(const CUtensorMap* tmD, uint8_t (*stage)[4096], int& sbuf, uint32_t tmem_row,
                                             int nch, int row, int n0, int bz, const GemmParams& p, int lane) {
  if (p.resid) epilogue_tma_impl<true>(tmD, stage, sbuf, tmem_row, nch, row, n0, bz, p, lane);
  else epilogue_tma_impl<false>(tmD, stage, sbuf, tmem_row, nch, row, n0, bz, p, lane);
}


// fp32 accumulate epilogue of one 32-row slab: D[rows, 32-column group] += alpha * acc, by TMA reduce-add from a
// 128B-swizzled [32][32] fp32 slab.  Used for weight gradients: every K slice and every task of a micro-step adds into
// the same fp32 buffer, so split-K needs neither a partial workspace nor a reduce kernel.
__device__ __forceinline__ void epilogue_reduce(const CUtensorMap* tmD, uint8_t (*stage)[4096], int& sbuf, uint32_t tmem_row,
                                                int nch, int row0, int n0, int bz, const GemmParams& p, int lane) {
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    const int nb = n0 + c * 32;
    if (nb >= p.N) break;
    uint8_t* sb = stage[sbuf];
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
    uint32_t r[32];
    tmem_ld32(tmem_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      float4 v = make_float4(__uint_as_float(r[4 * q4]) * p.alpha, __uint_as_float(r[4 * q4 + 1]) * p.alpha,
                             __uint_as_float(r[4 * q4 + 2]) * p.alpha, __uint_as_float(r[4 * q4 + 3]) * p.alpha);
      *reinterpret_cast<float4*>(sb + lane * 128 + ((q4 ^ (lane & 7)) << 4)) = v;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_3d(tmD, sb, nb, row0, bz);
      tma_store_commit();
    }
    sbuf ^= 1;
  }
}

// PAIR = 1: launched as clusters of two CTAs that own vertically adjacent 128-row tiles of the same column tile.  The B
// (weight) stage is shared: each CTA fetches half of it and TMA-multicasts it into both CTAs' shared memory, which halves
// the L2 -> SM traffic of the B operand (the 128 x 256 single-CTA kernel saturates the ~12 TB/s L2 fabric at ~750
// TFLOP/s).  A stage may be refilled only after BOTH CTAs' MMAs have consumed it: tcgen05.commit multicasts the release.
template <int A_MN, int B_MN, typename OutT, int BN, int PAIR>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmD, GemmParams p) {
  using C = Cfg<BN>;
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int wstart = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout<BN>& sm =
      *reinterpret_cast<SmemLayout<BN>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.total;
  const int nkb_all = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], PAIR ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tmem_full[i], 1);
      mbar_init(&sm.tmem_empty[i], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  // TMEM footprint matters beyond this kernel: a smaller allocation lets the next kernel's CTAs start while ours drain
  // Row-sum calls (weight gradients: about one long tile per CTA, nothing for a second accumulator to overlap) run
  // single-buffered -- 128 accumulator + 16 row-sum columns -- so that the allocation stays at 256 columns.
  const bool rs_mode = p.rowsum != nullptr;
  const uint32_t tmem_cols = (uint32_t)C::kTmemCols;
  if (warp == 2) tmem_alloc_n(&sm.tmem_addr, tmem_cols);
  if (p.rowsum) {
    for (int e = threadIdx.x; e < 256; e += kThreads) reinterpret_cast<uint32_t*>(sm.ones)[e] = 0x3F803F80u;   // bf16 1.0 pairs
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  pdl_sync();   // on-chip set-up done; from here on the kernel touches global memory
  const uint32_t tmem_base = sm.tmem_addr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int w = wstart; w < total; w += wstep) {
        const Work wk = decode<BN, PAIR>(w, p, crank);
        const int m0 = wk.m0, n0 = wk.n0;
        const int kb0 = wk.sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
        const int nchunk = wk.bn / 64;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], (BM + wk.bn) * BK * 2);
          uint8_t* sa = sm.tiles[s];
          uint8_t* sb = sa + BM * BK * 2;
          if (A_MN) {
            tma_load_3d(sa, &tmA, &sm.full[s], m0, kb * BK, wk.bz);
            tma_load_3d(sa + BK * 128, &tmA, &sm.full[s], m0 + 64, kb * BK, wk.bz);
          } else {
            tma_load_3d(sa, &tmA, &sm.full[s], kb * BK, m0, wk.bz);
          }
          if (PAIR) {   // this CTA fetches chunks [crank*nchunk/2, (crank+1)*nchunk/2) for both CTAs
            const int c_lo = crank * (nchunk >> 1), c_hi = c_lo + (nchunk >> 1);
            if (B_MN) {
              for (int c = c_lo; c < c_hi; ++c) tma_load_3d_mc(sb + c * BK * 128, &tmB, &sm.full[s], n0 + 64 * c, kb * BK, wk.bz, 3);
            } else {
              for (int c = c_lo; c < c_hi; ++c) tma_load_3d_mc(sb + c * 64 * 128, &tmB, &sm.full[s], kb * BK, n0 + 64 * c, wk.bz, 3);
            }
          } else if (B_MN) {
            for (int c = 0; c < nchunk; ++c) tma_load_3d(sb + c * BK * 128, &tmB, &sm.full[s], n0 + 64 * c, kb * BK, wk.bz);
          } else {   // 64-row boxes stacked at the 1024 B / 8-row pitch
            for (int c = 0; c < nchunk; ++c) tma_load_3d(sb + c * 64 * 128, &tmB, &sm.full[s], kb * BK, n0 + 64 * c, wk.bz);
          }
          if (++s == C::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0, ph = 0, it = 0;
      for (int w = wstart; w < total; w += wstep, ++it) {
        const Work wk = decode<BN, PAIR>(w, p, crank);
        const uint32_t idesc = umma_idesc_bf16(BM, wk.bn, A_MN, B_MN);
        const int kb0 = wk.sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
        const int acc = rs_mode ? 0 : (it & 1);
        mbar_wait(&sm.tmem_empty[acc], ((rs_mode ? it : (it >> 1)) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const bool rs_tile = p.rowsum && (wk.n0 % BN) == 0;
        const int rs_nt = wk.n0 / BN;
        uint32_t rs_acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(sm.tiles[s]);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 bf16 = 32 B inside the 128B swizzle span; MN-major: 16 k-rows = 2 swizzle atoms = 2048 B.
            // K-major B is a stack of 64-row boxes: rows continue at the same 1024 B / 8-row pitch.
            const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_f16(tmem_d, da, db, idesc, (kb != kb0) | (k != 0));
          }
          // row sums of A on the tensor core: N = 16 MMAs against the all-ones tile.  They re-read the A stage (the 128 x 128
          // kernel is shared-memory-bandwidth bound), so the k-blocks are dealt round-robin to the column tiles of a row block
          if (rs_tile && kb % p.tiles_n == rs_nt) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
              umma_f16(tmem_base + BN, da, umma_smem_desc(smem_u32(sm.ones) + k * 32, 16, 0),
                       umma_idesc_bf16(BM, 16, A_MN, 0), rs_acc | (k != 0));
            }
            rs_acc = 1;
          }
          if (PAIR) umma_commit_mc(&sm.empty[s], 3);  // both CTAs' producers write into this stage of both CTAs
          else umma_commit(&sm.empty[s]);             // frees the smem stage when these MMAs retire
          if (++s == C::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&sm.tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    int it = 0, sbuf = 0;
    for (int w = wstart; w < total; w += wstep, ++it) {
      const Work wk = decode<BN, PAIR>(w, p, crank);
      const int m0 = wk.m0, n0 = wk.n0;
      const int nch = wk.bn / 32;
      const int acc = rs_mode ? 0 : (it & 1);
      const int row = m0 + q * 32 + lane;
      mbar_wait(&sm.tmem_full[acc], (rs_mode ? it : (it >> 1)) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      if (sizeof(OutT) == 4 && p.reduce_f32) {
        if (m0 + q * 32 < p.M) epilogue_reduce(&tmD, sm.stage_out[q], sbuf, tmem_d, nch, m0 + q * 32, n0, wk.bz, p, lane);
        const int rs_kb0 = wk.sp * p.kb_per_split, rs_tn = p.tiles_n;
        const int rs_first = rs_kb0 + (((n0 / BN) - rs_kb0 % rs_tn) + rs_tn) % rs_tn;      // first k-block dealt to this tile
        if (p.rowsum && (n0 % BN) == 0 && rs_first < min(nkb_all, rs_kb0 + p.kb_per_split)) {
          uint32_t r[16];
          tmem_ld16(tmem_base + BN + ((uint32_t)(q * 32) << 16), r);
          tmem_ld_wait();
          if (row < p.M) atomicAdd(p.rowsum + row, __uint_as_float(r[0]) * p.alpha);
        }
      } else if (sizeof(OutT) == 2 && p.tma_store) {
        if (m0 + q * 32 < p.M) epilogue_tma(&tmD, sm.stage_out[q], sbuf, tmem_d, nch, row, n0, wk.bz, p, lane);
      } else if (p.splits > 1) {
        float* W = p.ws + ((size_t)(wk.sp * p.batch + wk.bz) * p.M) * p.N;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_d + c * 32, r);
          tmem_ld_wait();
          const int nb = n0 + c * 32;
          if (row < p.M && nb < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            store_row32<float>(W + (size_t)row * p.N + nb, v, min(32, p.N - nb));
          }
        }
      } else {
        OutT* D = reinterpret_cast<OutT*>(p.D) + (long long)wk.bz * p.batch_stride_d;
        const OutT* R = p.resid ? reinterpret_cast<const OutT*>(p.resid) + (long long)wk.bz * p.batch_stride_r : nullptr;
        const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_d + c * 32, r);
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
            const float al = (p.alpha_cols == 0 || nb < p.alpha_cols) ? p.alpha : 1.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= al;
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
            store_row32<OutT>(D + (long long)row * p.ldd + nb, v, nvalid);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);  // accumulator buffer may be overwritten
    }
    if (lane == 0) tma_store_wait_read<0>();   // shared memory must outlive the last bulk stores
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // neither CTA may retire while the peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_n(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// cta_group::2 variant: a cluster of two CTAs (one SM pair) computes a 256 x 256 tile with ONE tcgen05.mma stream issued by
// the leader CTA.  Each CTA stages its own 128 A rows and only HALF of the B tile (128 of the 256 n-rows); the tensor
// cores read the other half from the peer SM's shared memory.  Per SM and 64-deep k-block that is 32 KB of L2 -> SM
// traffic for 2 M MACs (64 B/clk at full MMA rate) instead of 48 KB (96 B/clk) for the single-CTA 128 x 256 tile, which
// ncu showed pinned at ~52 % tensor-pipe activity by the SM's L2 read port (profiles/r01_ncu_gemm_*).
// Barriers: TMA of both CTAs signals the LEADER's full[s]; tcgen05.commit multicasts to both CTAs' empty[s] / tmem_full;
// the peer's epilogue warps arrive remotely on the leader's tmem_empty.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStages2 = 6;
constexpr int kStageBytes2 = (BM * BK + 128 * BK) * 2;   // A rows of this CTA + this CTA's half of B

struct SmemLayout2 {
  uint8_t tiles[kStages2][kStageBytes2];
  uint8_t stage_out[4][2][4096];
  uint64_t full[kStages2];      // used on the leader: its arrive.expect_tx + the TMA bytes of both CTAs
  uint64_t empty[kStages2];     // per CTA, released by the leader's multicast commit
  uint64_t tmem_full[2];        // per CTA, multicast commit after the last k-block
  uint64_t tmem_empty[2];       // leader only: 4 local + 4 remote epilogue warps
  uint32_t tmem_addr;
};

struct Work2 {
  int m0, n0, bn, sp;   // m0: first row of the PAIR tile (256 rows); bn: 256 or 128 (tail sub-tile); sp: K slice
};
__device__ __forceinline__ Work2 decode2(int w, const GemmParams& p) {
  Work2 r;
  int sub = 0;
  r.bn = 256;
  r.sp = 0;
  if (p.splits > 1) {          // split-K: slices of one tile are adjacent work items (no tail sub-tiles)
    r.sp = w % p.splits;
    w /= p.splits;
  } else if (w >= p.n_big) {
    const int u = w - p.n_big;
    sub = u & 1;
    w = p.n_big + (u >> 1);
    r.bn = 128;
  }
  const int nt = w % p.tiles_n;
  const int mt = w / p.tiles_n;
  r.m0 = mt * 256;
  r.n0 = nt * 256 + sub * 128;
  return r;
}

template <int A_MN, int B_MN, typename OutT>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmD, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemLayout2& sm = *reinterpret_cast<SmemLayout2*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  const int wstart = blockIdx.x >> 1, wstep = gridDim.x >> 1;
  const int total = p.total;
  const int nkb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&sm.full[s], 1);     // the leader's arrive.expect_tx; the peer contributes bytes only
      mbar_init(&sm.empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tmem_full[i], 1);
      mbar_init(&sm.tmem_empty[i], 8);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(&sm.tmem_addr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = sm.tmem_addr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      for (int w = wstart; w < total; w += wstep) {
        const Work2 wk = decode2(w, p);
        const int m0 = wk.m0 + crank * 128;             // this CTA's A rows
        const int nh = wk.bn >> 1;                      // n-rows of B held by each CTA
        const int n0 = wk.n0 + crank * nh;
        const int kb0 = wk.sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.empty[s], ph ^ 1);
          uint8_t* sa = sm.tiles[s];
          uint8_t* sb = sa + BM * BK * 2;
          // The leader arms its barrier with the bytes of BOTH CTAs.  The peer only issues its loads: a remote arrive here
          // would cost a MEMBAR.ALL.GPU per k-block (release.cluster) and serialise the peer's ring (seen in ncu).
          if (leader) mbar_expect_tx(&sm.full[s], 2 * (BM + nh) * BK * 2);
          if (A_MN) {
            tma_load_3d_2sm(sa, &tmA, &sm.full[s], m0, kb * BK, 0);
            tma_load_3d_2sm(sa + BK * 128, &tmA, &sm.full[s], m0 + 64, kb * BK, 0);
          } else {
            tma_load_3d_2sm(sa, &tmA, &sm.full[s], kb * BK, m0, 0);
          }
          for (int c = 0; c < nh / 64; ++c) {
            if (B_MN) tma_load_3d_2sm(sb + c * BK * 128, &tmB, &sm.full[s], n0 + 64 * c, kb * BK, 0);
            else tma_load_3d_2sm(sb + c * 64 * 128, &tmB, &sm.full[s], kb * BK, n0 + 64 * c, 0);
          }
          if (++s == kStages2) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int s = 0, ph = 0, it = 0;
      for (int w = wstart; w < total; w += wstep, ++it) {
        const Work2 wk = decode2(w, p);
        const uint32_t idesc = umma_idesc_bf16(256, wk.bn, A_MN, B_MN);
        const int acc = it & 1;
        mbar_wait(&sm.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        const int kb0 = wk.sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(sm.tiles[s]);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, BK * 128, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, BK * 128, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_f16_2sm(tmem_d, da, db, idesc, (kb != kb0) | (k != 0));
          }
          umma_commit_2sm_mc(&sm.empty[s], 3);
          if (++s == kStages2) { s = 0; ph ^= 1; }
        }
        umma_commit_2sm_mc(&sm.tmem_full[acc], 3);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    int it = 0, sbuf = 0;
    OutT* D = reinterpret_cast<OutT*>(p.D);
    const OutT* R = reinterpret_cast<const OutT*>(p.resid);
    const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
    for (int w = wstart; w < total; w += wstep, ++it) {
      const Work2 wk = decode2(w, p);
      const int acc = it & 1;
      const int row = wk.m0 + crank * 128 + q * 32 + lane;
      mbar_wait(&sm.tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      const int nch = wk.bn / 32;
      if (sizeof(OutT) == 4 && p.reduce_f32) {
        if (row - lane < p.M) epilogue_reduce(&tmD, sm.stage_out[q], sbuf, tmem_d, nch, row - lane, wk.n0, 0, p, lane);
      } else if (p.splits > 1) {
        float* W = p.ws + (size_t)wk.sp * p.M * p.N;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_d + c * 32, r);
          tmem_ld_wait();
          const int nb = wk.n0 + c * 32;
          if (row < p.M && nb < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            store_row32<float>(W + (size_t)row * p.N + nb, v, min(32, p.N - nb));
          }
        }
      } else if (sizeof(OutT) == 2 && p.tma_store) {
        if (row - lane < p.M) epilogue_tma(&tmD, sm.stage_out[q], sbuf, tmem_d, nch, row, wk.n0, 0, p, lane);
      } else {
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_d + c * 32, r);
        tmem_ld_wait();
        const int nb = wk.n0 + c * 32;
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
          const float al = (p.alpha_cols == 0 || nb < p.alpha_cols) ? p.alpha : 1.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= al;
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
          store_row32<OutT>(D + (long long)row * p.ldd + nb, v, nvalid);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&sm.tmem_empty[acc]);
        else mbar_arrive_remote(&sm.tmem_empty[acc], 0);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// out[b][m][n] = epi(sum_sp ws[sp][b][m][n])
template <typename OutT>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmParams p) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)p.M * p.N;
  if (idx >= per * p.batch) return;
  const int bz = (int)(idx / per);
  const long long mn = idx % per;
  const int m = (int)(mn / p.N), n = (int)(mn % p.N);
  float v = 0.f;
  for (int sp = 0; sp < p.splits; ++sp) v += p.ws[((size_t)(sp * p.batch + bz)) * per + mn];
  if (p.bias) v += ld_as_float<OutT>(reinterpret_cast<const OutT*>(p.bias) + n);
  v *= (p.alpha_cols == 0 || n < p.alpha_cols) ? p.alpha : 1.f;
  if (p.act == 1) v = gelu_erf(v);
  if (p.resid) v += ld_as_float<OutT>(reinterpret_cast<const OutT*>(p.resid) + (long long)bz * p.batch_stride_r + (long long)m * p.ldr + n);
  reinterpret_cast<OutT*>(p.D)[(long long)bz * p.batch_stride_d + (long long)m * p.ldd + n] = (OutT)v;
}

template <int A_MN, int B_MN, typename OutT>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_tc2_kernel<A_MN, B_MN, OutT>;
  static bool configured = false;
  const int smem = (int)sizeof(SmemLayout2) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int clusters = p.total < kNumSMs / 2 ? p.total : kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_ofa_pdl ? 2 : 1;
  OFA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, p));
  OFA_LAUNCH_CHECK("gemm_tc2_kernel");
  if (p.splits > 1 && !p.reduce_f32) {
    const long long n = (long long)p.M * p.N;
    OFA_CUDA(ofa_launch_pdl(splitk_reduce_kernel<OutT>, (unsigned)((n + 255) / 256), 256, 0, st, p));
    OFA_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return 0;
}

template <int A_MN, int B_MN, typename OutT, int BN, int PAIR>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_tc_kernel<A_MN, B_MN, OutT, BN, PAIR>;
  static bool configured = false;  // per template instantiation
  const int smem = (int)sizeof(SmemLayout<BN>) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int total = p.total;
  if (PAIR) {
    const int clusters = total < kNumSMs / 2 ? total : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_ofa_pdl ? 2 : 1;
    OFA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, p));
  } else {
    OFA_CUDA(ofa_launch_pdl(kern, total < kNumSMs ? total : kNumSMs, kThreads, smem, st, ta, tb, td, p));
  }
  OFA_LAUNCH_CHECK("gemm_tc_kernel");
  if (p.splits > 1 && !p.reduce_f32) {
    const long long n = (long long)p.M * p.N * p.batch;
    OFA_CUDA(ofa_launch_pdl(splitk_reduce_kernel<OutT>, (unsigned)((n + 255) / 256), 256, 0, st, p));
    OFA_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return 0;
}

template <int A_MN, int B_MN, typename OutT>
int launch_bn(int bn, int pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, cudaStream_t st) {
  if (pair) return bn == 256 ? launch<A_MN, B_MN, OutT, 256, 1>(ta, tb, td, p, st) : launch<A_MN, B_MN, OutT, 128, 1>(ta, tb, td, p, st);
  if constexpr (A_MN == 0 && sizeof(OutT) == 2) {      // 128 x 64 tiles: forward / dgrad of the small-M problems (decoder steps)
    if (bn == 64) return launch<A_MN, B_MN, OutT, 64, 0>(ta, tb, td, p, st);
  }
  return bn == 256 ? launch<A_MN, B_MN, OutT, 256, 0>(ta, tb, td, p, st) : launch<A_MN, B_MN, OutT, 128, 0>(ta, tb, td, p, st);
}

// split-K plan of the cta_group::2 kernel for problems with fewer than one round of 256 x 256 pair tiles (weight gradients:
// K = tokens): 0 = not applicable
int plan2_splits(int M, int N, int K, int batch) {
  if (batch != 1 || M < 256 || N < 256) return 0;
  const int nkb = (K + BK - 1) / BK;
  const long long t2 = (long long)((M + 255) / 256) * ((N + 255) / 256);
  const int workers = kNumSMs / 2;
  if (t2 >= workers) return 0;
  int s = (int)(workers / t2);
  if (s > nkb / 64) s = nkb / 64;      // the pair pipeline needs a long main loop per slice to beat 128 x 128 tiles
  if (s < 2 || t2 * s < (workers * 3) / 4) return 0;
  return s;
}

void plan(int M, int N, int K, int batch, int* bn, int* splits) {
  const int tm = (M + BM - 1) / BM;
  const int nkb = (K + BK - 1) / BK;
  // Tile width by a wave-count model: a 128 x 256 tile takes ~1.5x the time of a 128 x 128 tile (measured ~750 vs ~550
  // TFLOP/s when the machine is full), so it wins whenever it saves enough waves
  const long long t128 = (long long)tm * ((N + 127) / 128) * batch, t256 = (long long)tm * ((N + 255) / 256) * batch;
  const long long w128 = (t128 + kNumSMs - 1) / kNumSMs, w256 = (t256 + kNumSMs - 1) / kNumSMs;
  *bn = (N >= 256 && 3 * w256 < 2 * w128) ? 256 : 128;
  const long long tiles = (long long)tm * ((N + *bn - 1) / *bn) * batch;
  int s = 1;
  if (tiles * 2 <= kNumSMs && nkb >= 8) {
    s = (int)(kNumSMs / tiles);
    if (s > nkb / 4) s = nkb / 4;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
  }
  *splits = s;
}

}  // namespace

// debugging / A-B switch for the CTA-pair (TMA multicast) variant; returns the previous setting
extern "C" int ofa_gemm_set_pair_min_tiles(int n) {
  const int old = g_ofa_gemm_pair_min_tiles;
  g_ofa_gemm_pair_min_tiles = n;
  return old;
}
extern "C" int ofa_gemm_set_small64(int enabled) {
  const int old = g_ofa_gemm_small64;
  g_ofa_gemm_small64 = enabled;
  return old;
}
extern "C" int ofa_gemm_set_wgrad_bn256(int enabled) {
  const int old = g_ofa_gemm_wgrad_bn256;
  g_ofa_gemm_wgrad_bn256 = enabled;
  return old;
}
extern "C" int ofa_gemm_set_tma_store(int enabled) {
  const int old = g_ofa_gemm_tma_store;
  g_ofa_gemm_tma_store = enabled;
  return old;
}
extern "C" int ofa_gemm_set_pair_mode(int enabled) {
  const int old = g_ofa_gemm_pair_enabled;
  g_ofa_gemm_pair_enabled = enabled;
  return old;
}

// host helper: bytes of fp32 split-K workspace ofa_gemm_bf16 wants for this problem (0 = none)
extern "C" long long ofa_gemm_workspace_bytes(int M, int N, int K, int batch) {
  int bn, splits;
  plan(M, N, K, batch, &bn, &splits);
  const int s2 = plan2_splits(M, N, K, batch);
  if (s2 > splits) splits = s2;
  return splits > 1 ? (long long)splits * batch * M * N * (long long)sizeof(float) : 0;
}

// see include/ofa_b200.h
extern "C" int ofa_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int batch, long long lda,
                             long long ldb, long long ldd, long long stride_a, long long stride_b, long long stride_d,
                             int a_mn_major, int b_mn_major, int out_dtype, const void* bias, float alpha, int act,
                             const void* resid, long long ldr, long long stride_r, void* workspace,
                             long long workspace_bytes, int alpha_cols, void* stream) {
  OFA_CHECK(alpha_cols >= 0 && alpha_cols % 32 == 0, "ofa_gemm_bf16: alpha_cols=%d must be a multiple of 32", alpha_cols);
  OFA_CHECK(M > 0 && N > 0 && K > 0 && batch > 0, "ofa_gemm_bf16: empty problem M=%d N=%d K=%d batch=%d", M, N, K, batch);
  OFA_CHECK(lda % 8 == 0 && ldb % 8 == 0, "ofa_gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16B stride)");
  OFA_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "ofa_gemm_bf16: A/B must be 16B aligned");
  OFA_CHECK(stride_a % 8 == 0 && stride_b % 8 == 0, "ofa_gemm_bf16: batch strides must be multiples of 8 elements");
  const int reduce_f32 = out_dtype == OFA_F32_ACC;
  float* rowsum = nullptr;
  if (reduce_f32) {
    rowsum = (float*)bias;    // fp32-accumulate mode: `bias` names an fp32 [M] buffer that receives alpha * row sums of A
    bias = nullptr;
    OFA_CHECK(!resid && act == 0, "ofa_gemm_bf16: the fp32-accumulate output takes no residual / activation");
    OFA_CHECK(!rowsum || batch == 1, "ofa_gemm_bf16: row sums need batch == 1");
    OFA_CHECK(ldd % 4 == 0 && N % 4 == 0 && ((uintptr_t)D & 15) == 0 && stride_d % 4 == 0,
              "ofa_gemm_bf16: fp32-accumulate output needs 16-byte aligned rows (N=%d ldd=%lld)", N, ldd);
    out_dtype = OFA_F32;
  }
  int bn, splits;
  plan(M, N, K, batch, &bn, &splits);
  if (reduce_f32 && g_ofa_gemm_wgrad_bn256 && batch == 1 && N >= 256 && N % 256 == 0) {
    // weight gradients (K slices reduce-add in place, about one long tile per CTA): 128 x 256 tiles halve the shared-memory
    // traffic per MAC of the 128 x 128 tile that the occupancy heuristic of plan() would pick for so few tiles
    const long long tiles = (long long)((M + BM - 1) / BM) * (N / 256);
    const int nkb_r = (K + BK - 1) / BK;
    int s = 1;
    if (tiles * 2 <= kNumSMs && nkb_r >= 8) {
      s = (int)(kNumSMs / tiles);
      if (s > nkb_r / 4) s = nkb_r / 4;
      if (s > 32) s = 32;
      if (s < 1) s = 1;
    }
    if (tiles * s >= kNumSMs / 2) { bn = 256; splits = s; }
  }
  if (!reduce_f32 && g_ofa_gemm_small64 && batch == 1 && !a_mn_major && out_dtype == OFA_BF16 && N % 64 == 0) {
    // small-M problems (one decoder step: M = beams; the 12-token decoder groups of a training step): fewer 128-wide tiles than
    // half the machine.  128 x 64 tiles double the CTA count, and K up to 1984 then runs unsplit -- the fp32 partials and the
    // reduce pass of split-K cost more than the second half of the k-loop (M = 576, N = K = 768: 19.3 us split vs cuBLAS 9.1)
    const int tm = (M + BM - 1) / BM, nkb64 = (K + BK - 1) / BK;
    if ((long long)tm * ((N + 127) / 128) * 2 <= kNumSMs) {
      const int tiles = tm * (N / 64);
      int s = 1;
      if (tiles * 2 <= kNumSMs && nkb64 >= 32) {
        s = kNumSMs / tiles;
        if (s > nkb64 / 8) s = nkb64 / 8;
        if (s < 1) s = 1;
      }
      if (s <= splits || splits == 1) { bn = 64; splits = s; }
    }
  }
  if (!reduce_f32 && splits > 1 &&
      (workspace == nullptr || workspace_bytes < (long long)splits * batch * M * N * (long long)sizeof(float)))
    splits = 1;  // no workspace: run unsplit (still correct, just fewer CTAs)
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
    else            { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = 64; }
    strides[0] = (uint64_t)ldb * 2;
    strides[1] = (uint64_t)(batch > 1 ? stride_b : (long long)dims[1] * ldb) * 2;
    if (int e = ofa_make_tmap(&tb, B, 3, dims, strides, box, 1, 2)) return e;
  }
  CUtensorMap td;
  memset(&td, 0, sizeof(td));
  const int nkb_plan = (K + BK - 1) / BK;
  const int kbps = (nkb_plan + splits - 1) / splits;
  const bool unsplit = (nkb_plan + kbps - 1) / kbps == 1;
  int tma_store = 0;
  // TMA stores clip at 16-byte granularity: with N % 8 != 0 the columns [N, ceil8(N)) of a row are written too (zeros), which
  // is only acceptable when they are the row's own padding
  if (g_ofa_gemm_tma_store && out_dtype == OFA_BF16 && unsplit && ldd % 8 == 0 && ((uintptr_t)D & 15) == 0 &&
      (batch == 1 || stride_d % 8 == 0) && (N % 8 == 0 || ldd == (N + 7) / 8 * 8)) {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)batch}, strides[2];
    uint32_t box[3] = {64, 32, 1};
    strides[0] = (uint64_t)ldd * 2;
    strides[1] = (uint64_t)(batch > 1 ? stride_d : (long long)M * ldd) * 2;
    if (int e = ofa_make_tmap(&td, D, 3, dims, strides, box, 1, 2)) return e;
    tma_store = 1;
  }
  if (reduce_f32) {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)batch}, strides[2];
    uint32_t box[3] = {32, 32, 1};
    strides[0] = (uint64_t)ldd * 4;
    strides[1] = (uint64_t)(batch > 1 ? stride_d : (long long)M * ldd) * 4;
    if (int e = ofa_make_tmap(&td, D, 3, dims, strides, box, 1, 4)) return e;
  }
  GemmParams p;
  p.rowsum = rowsum;
  p.reduce_f32 = reduce_f32;
  p.tma_store = tma_store;
  p.D = D; p.bias = bias; p.resid = resid; p.ws = (float*)workspace; p.ldd = ldd; p.ldr = ldr;
  p.batch_stride_d = stride_d; p.batch_stride_r = stride_r;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.alpha_cols = alpha_cols; p.act = act;
  p.tiles_m = (M + BM - 1) / BM; p.tiles_n = (N + bn - 1) / bn; p.batch = batch; p.splits = splits;
  const int nkb = (K + BK - 1) / BK;
  p.kb_per_split = (nkb + splits - 1) / splits;
  p.splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;  // drop empty trailing slices
  // cta_group::2 (mode 2): 256 x 256 pair tiles for plain problems with at least one full round of pair tiles, or split
  // along K when the output is small and the contraction long (weight gradients)
  int s2 = (g_ofa_gemm_pair_enabled == 2 && !rowsum) ? plan2_splits(M, N, K, batch) : 0;
  if (!reduce_f32 && s2 > 1 && (workspace == nullptr || workspace_bytes < (long long)s2 * M * N * (long long)sizeof(float))) s2 = 0;
  if (g_ofa_gemm_pair_enabled == 2 && batch == 1 && N >= 256 && !rowsum &&
      (s2 > 1 || (p.splits == 1 && (long long)((M + 255) / 256) * ((N + 255) / 256) >= g_ofa_gemm_pair_min_tiles))) {
    p.tiles_m = (M + 255) / 256;
    p.tiles_n = (N + 255) / 256;
    const int tiles = p.tiles_m * p.tiles_n, workers = kNumSMs / 2;
    p.n_big = tiles;
    p.total = tiles;
    p.splits = 1;
    p.kb_per_split = nkb;
    if (s2 > 1) {
      p.kb_per_split = (nkb + s2 - 1) / s2;
      p.splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;
      p.total = tiles * p.splits;
      p.tma_store = 0;
    }
    const int rem = tiles % workers;
    if (p.splits == 1 && tiles > workers && rem > 0 && rem <= (workers * 3) / 4) {
      p.n_big = tiles - rem;
      p.total = p.n_big + rem * 2;
    }
    cudaStream_t st2 = (cudaStream_t)stream;
    const int sel2 = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
    if (out_dtype == OFA_BF16) {
      switch (sel2) {
        case 0: return launch2<0, 0, __nv_bfloat16>(ta, tb, td, p, st2);
        case 1: return launch2<0, 1, __nv_bfloat16>(ta, tb, td, p, st2);
        case 2: return launch2<1, 0, __nv_bfloat16>(ta, tb, td, p, st2);
        default: return launch2<1, 1, __nv_bfloat16>(ta, tb, td, p, st2);
      }
    } else if (out_dtype == OFA_F32) {
      switch (sel2) {
        case 0: return launch2<0, 0, float>(ta, tb, td, p, st2);
        case 1: return launch2<0, 1, float>(ta, tb, td, p, st2);
        case 2: return launch2<1, 0, float>(ta, tb, td, p, st2);
        default: return launch2<1, 1, float>(ta, tb, td, p, st2);
      }
    }
    return ofa_set_error("ofa_gemm_bf16: bad out_dtype %d", out_dtype);
  }
  // CTA pairs with B multicast (mode 1) for plain problems with enough row tiles
  const int pair = (batch == 1 && p.splits == 1 && p.tiles_m >= 2 && (long long)p.tiles_m * p.tiles_n >= kNumSMs &&
                    g_ofa_gemm_pair_enabled == 1) ? 1 : 0;
  if (pair) p.tiles_m = (p.tiles_m + 1) / 2;
  {
    const int workers = pair ? kNumSMs / 2 : kNumSMs;
    const int tiles = p.tiles_m * p.tiles_n * batch * p.splits;
    p.n_big = tiles;
    p.total = tiles;
    // tail splitting (plain problems only): the last (tiles mod workers) tiles are issued as narrower sub-tiles
    const int rem = tiles % workers;
    const int sub = pair ? 128 : 64;
    if (batch == 1 && p.splits == 1 && tiles > workers && rem > 0 && rem <= (workers * 3) / 4 && bn > sub) {
      p.n_big = tiles - rem;
      p.total = p.n_big + rem * (bn / sub);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int sel = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  if (out_dtype == OFA_BF16) {
    switch (sel) {
      case 0: return launch_bn<0, 0, __nv_bfloat16>(bn, pair, ta, tb, td, p, st);
      case 1: return launch_bn<0, 1, __nv_bfloat16>(bn, pair, ta, tb, td, p, st);
      case 2: return launch_bn<1, 0, __nv_bfloat16>(bn, pair, ta, tb, td, p, st);
      default: return launch_bn<1, 1, __nv_bfloat16>(bn, pair, ta, tb, td, p, st);
    }
  } else if (out_dtype == OFA_F32) {
    switch (sel) {
      case 0: return launch_bn<0, 0, float>(bn, pair, ta, tb, td, p, st);
      case 1: return launch_bn<0, 1, float>(bn, pair, ta, tb, td, p, st);
      case 2: return launch_bn<1, 0, float>(bn, pair, ta, tb, td, p, st);
      default: return launch_bn<1, 1, float>(bn, pair, ta, tb, td, p, st);
    }
  }
  return ofa_set_error("ofa_gemm_bf16: bad out_dtype %d", out_dtype);
}
