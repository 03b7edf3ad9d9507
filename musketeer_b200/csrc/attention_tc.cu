// bf16 flash attention forward for OFA on sm_100a: TMA-staged 128B-swizzled tiles, S = Q'K'^T and O += P V on the
// tcgen05 tensor cores with accumulators in TMEM, online softmax in registers (one thread per query row).
//
//   Q' = [q*s ; pos_q*s]  K' = [k ; pos_k]  (d_qk = 128, d_v = 64): the absolute-position term pos_q.pos_k^T of
//   models/ofa/unify_transformer.py:906-912,1297-1318 is a second K-block of the same MMA instead of a [B,H,N,N] tensor;
//   the per-layer relative-position bias (:640-658,923-933,1282-1295,1519-1529) is a LUT lookup on the S tile;
//   key padding (-inf), causal mask, fp32 softmax, c_attn head scale: unify_multihead_attention.py:345-398.
// One CTA = one (batch, head, 128-query tile) sweeping 64-key tiles; 128 threads; 72 KB smem and 128 TMEM columns
// per CTA so several CTAs share an SM and overlap each other's MMA / softmax phases.
#include <math_constants.h>

#include <type_traits>

#include "attention_common.cuh"

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int kThreads = 256;   // 2 threads per query row: 32 key columns / 32 output columns each
constexpr int kTmemCols = 128;  // S: [0,64)  O_partial: [64,128)
constexpr int kTokLut = 1024 + 128;

struct TcSmem {
  uint8_t q[2][BQ * 128];   // [0: q, 1: pos_q][128 rows][128 B]   K-major, SWIZZLE_128B
  uint8_t k[2][BKV * 128];  // [0: k, 1: pos_k][64 keys][128 B]    K-major
  uint8_t v[BKV * 128];     // [64 keys][64 x bf16]                MN-major B operand of P.V
  uint8_t p[BQ * 128];      // [128 rows][64 keys x bf16]          K-major A operand of P.V
  float tok_s[kTokLut];     // token rel-pos LUT staged for this query tile: index (r - j_t) + S_t - 1
  float rowmax[2][BQ];      // per-tile partial row maxima of the two column halves
  float rowsum[BQ];         // final exchange of the partial row sums
  uint64_t bar_q, bar_k, bar_v, bar_s, bar_o;
  uint32_t tmem_addr;
  int kinfo[BKV];  // per key of the current tile: bit31 masked | bit30 image key | [0,16) kr*(2*ibs-1)+kc
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmPQ,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmPK,
                   const __grid_constant__ CUtensorMap tmV, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int t = threadIdx.x, warp = t >> 5;
  const int r = t & 127, hf = t >> 7;   // query row / TMEM lane ; column half
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const AttnBias& bz = a.bias;
  int ntiles = (a.S + BKV - 1) / BKV;
  if (a.causal) {   // key tiles entirely above the diagonal of this query tile are skipped
    const int last = q0 + BQ - 1 + a.q_pos_off;
    ntiles = min(ntiles, last / BKV + 1);
  }
  const int w83 = 2 * bz.ibs - 1;

  if (t == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmPQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmPK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_k, 1); mbar_init(&sm.bar_v, 1); mbar_init(&sm.bar_s, 1);
    mbar_init(&sm.bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(&sm.tmem_addr);
  pdl_sync();   // everything above is on-chip set-up; the LUT staging below reads the predecessor's output
  // token LUT slice for this query tile
  const int S_t = a.S - bz.k_text_off;
  const int i_t0 = q0 + a.q_pos_off - bz.q_text_off;
  if (bz.tok_lut) {
    for (int e = t; e < kTokLut; e += kThreads) {
      const int rel = i_t0 + e - (S_t - 1) + bz.tok_max - 1;
      sm.tok_s[e] = (rel >= 0 && rel < 2 * bz.tok_max - 1) ? bz.tok_lut[(size_t)h * (2 * bz.tok_max - 1) + rel] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = sm.tmem_addr + ((uint32_t)((warp & 3) * 32) << 16) + hf * 32;
  const uint32_t tmem_o = sm.tmem_addr + ((uint32_t)((warp & 3) * 32) << 16) + BKV + hf * 32;

  if (t == 0) {
    mbar_expect_tx(&sm.bar_q, 2 * BQ * 128);
    tma_load_4d(sm.q[0], &tmQ, &sm.bar_q, 0, h, q0, b);
    tma_load_4d(sm.q[1], &tmPQ, &sm.bar_q, 0, h, q0, b);
    mbar_expect_tx(&sm.bar_k, 2 * BKV * 128);
    tma_load_4d(sm.k[0], &tmK, &sm.bar_k, 0, h, 0, b);
    tma_load_4d(sm.k[1], &tmPK, &sm.bar_k, 0, h, 0, b);
    mbar_expect_tx(&sm.bar_v, BKV * 128);
    tma_load_4d(sm.v, &tmV, &sm.bar_v, 0, h, 0, b);
  }

  // per-row state
  const int i = q0 + r;
  const int iabs = i + a.q_pos_off;
  const bool row_ok = i < a.T;
  const bool q_text = bz.tok_lut && iabs >= bz.q_text_off;
  const bool q_img = bz.img_lut && iabs < bz.n_img_q && row_ok;
  int rowbase = 0;
  if (q_img) {
    const int pid = bz.q_pid[(size_t)b * bz.n_img_q + iabs] - 1;
    rowbase = (pid / bz.ibs + bz.ibs - 1) * w83 + (pid % bz.ibs + bz.ibs - 1);
  }
  const float* img_lut = bz.img_lut ? bz.img_lut + (size_t)h * bz.n_img_rel : nullptr;
  constexpr float kLog2e = 1.4426950408889634f;
  float m = -CUDART_INF_F, l = 0.f;
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;

  constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);
  const int col0 = hf * 32;

  for (int jt = 0; jt < ntiles; ++jt) {
    const int k0 = jt * BKV;
    const uint32_t ph = jt & 1;
    // key-side metadata for this tile
    int my_masked = 0;
    if (t < BKV) {
      const int j = k0 + t;
      int info = 0;
      if (j >= a.S || (a.kpm && a.kpm[(size_t)b * a.S + j])) { info |= (int)0x80000000u; my_masked = 1; }
      if (bz.img_lut && j < bz.n_img_k) {
        const int pid = bz.k_pid[(size_t)b * bz.n_img_k + j] - 1;
        info |= 0x40000000 | ((pid / bz.ibs) * w83 + (pid % bz.ibs));
      }
      sm.kinfo[t] = info;
    }
    if (t == 0) {
      if (jt == 0) mbar_wait(&sm.bar_q, 0);
      mbar_wait(&sm.bar_k, ph);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(sm.tmem_addr, umma_smem_desc(smem_u32(sm.q[kb]) + ks * 32, 16, 1024),
                   umma_smem_desc(smem_u32(sm.k[kb]) + ks * 32, 16, 1024), idesc_s, (kb | ks) != 0);
      umma_commit(&sm.bar_s);
    }
    const bool keys_any_masked = __syncthreads_or(my_masked) != 0;   // kinfo visible; also reconverges warp 0
    const bool keys_all_img = bz.img_lut != nullptr && (k0 + BKV <= bz.n_img_k);
    const bool keys_all_txt = bz.tok_lut != nullptr && (k0 >= bz.k_text_off) && (bz.img_lut == nullptr || k0 >= bz.n_img_k);
    // padded / masked keys inside this thread's 32 columns as a bit mask: the fast paths stay usable on such tiles
    uint32_t cm = 0;
    if (keys_any_masked) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) cm |= (sm.kinfo[col0 + jj] < 0 ? 1u : 0u) << jj;
    }
    const bool plain = !(a.causal && k0 + col0 + 31 > iabs);
    int mode = 3;
    if (plain) {
      if (q_text && keys_all_txt) mode = 1;
      else if (q_img && keys_all_img) mode = 2;
      else if ((keys_all_txt && !q_text) || (keys_all_img && !q_img) || (!bz.tok_lut && !bz.img_lut)) mode = 0;
    }
    const int tb = r + S_t - 1 - (k0 - bz.k_text_off) - col0;   // tok_s index of column col0; column jj -> tb - jj
    mbar_wait(&sm.bar_s, ph);
    tc_fence_after();
    if (t == 0 && jt + 1 < ntiles) {  // K' buffer is free: prefetch the next key tile under the softmax
      mbar_expect_tx(&sm.bar_k, 2 * BKV * 128);
      tma_load_4d(sm.k[0], &tmK, &sm.bar_k, 0, h, k0 + BKV, b);
      tma_load_4d(sm.k[1], &tmPK, &sm.bar_k, 0, h, k0 + BKV, b);
    }
    float s[32];
    {
      uint32_t rr[32];
      tmem_ld32(tmem_s, rr);
      tmem_ld_wait();
      if (mode == 0) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) s[jj] = (cm >> jj) & 1u ? -CUDART_INF_F : __uint_as_float(rr[jj]);
      } else if (mode == 1) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) s[jj] = (cm >> jj) & 1u ? -CUDART_INF_F : __uint_as_float(rr[jj]) + sm.tok_s[tb - jj];
      } else if (mode == 2) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          s[jj] = (cm >> jj) & 1u ? -CUDART_INF_F
                                  : __uint_as_float(rr[jj]) + __ldg(img_lut + rowbase - (sm.kinfo[col0 + jj] & 0xffff));
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int jl = col0 + jj, j = k0 + jl;
          const int info = sm.kinfo[jl];
          float x = __uint_as_float(rr[jj]);
          if (q_text && j >= bz.k_text_off) x += sm.tok_s[tb - jj];
          if (q_img && (info & 0x40000000)) x += __ldg(img_lut + rowbase - (info & 0xffff));
          if (info < 0 || (a.causal && j > iabs)) x = -CUDART_INF_F;
          s[jj] = x;
        }
      }
    }
    float mx = s[0];
#pragma unroll
    for (int jj = 1; jj < 32; ++jj) mx = fmaxf(mx, s[jj]);
    sm.rowmax[hf][r] = mx;
    __syncthreads();
    mx = fmaxf(m, fmaxf(sm.rowmax[0][r], sm.rowmax[1][r]));
    const float mu = (mx == -CUDART_INF_F) ? 0.f : mx;
    const float alpha = ex2((m - mu) * kLog2e);
    m = mx;
    float rs = 0.f;
    const float mneg = -mu * kLog2e;
#pragma unroll
    for (int c16 = 0; c16 < 4; ++c16) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        pv[e] = ex2(fmaf(s[c16 * 8 + e], kLog2e, mneg));
        rs += pv[e];
      }
      const uint4 pk = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                                  pack_bf16(pv[6], pv[7]));
      *reinterpret_cast<uint4*>(sm.p + r * 128 + (((hf * 4 + c16) ^ (r & 7)) << 4)) = pk;
    }
    l = l * alpha + rs;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      mbar_wait(&sm.bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BKV / 16; ++ks)
        umma_f16(sm.tmem_addr + BKV, umma_smem_desc(smem_u32(sm.p) + ks * 32, 16, 1024),
                 umma_smem_desc(smem_u32(sm.v) + ks * 2048, 1024, 1024), idesc_o, ks != 0);
      umma_commit(&sm.bar_o);
    }
    __syncwarp();
    mbar_wait(&sm.bar_o, ph);
    tc_fence_after();
    if (t == 0 && jt + 1 < ntiles) {  // V buffer is free
      mbar_expect_tx(&sm.bar_v, BKV * 128);
      tma_load_4d(sm.v, &tmV, &sm.bar_v, 0, h, k0 + BKV, b);
    }
    {
      uint32_t rr[32];
      tmem_ld32(tmem_o, rr);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) o[jj] = fmaf(o[jj], alpha, __uint_as_float(rr[jj]));
    }
    tc_fence_before();
  }

  // combine the two partial row sums, normalise, write this thread's 32 output columns
  if (hf == 1) sm.rowsum[r] = l;
  __syncthreads();
  if (hf == 0) l += sm.rowsum[r];
  __syncthreads();
  if (hf == 0) sm.rowsum[r] = l;
  __syncthreads();
  l = sm.rowsum[r];
  if (row_ok) {
    const float inv = (l > 0.f ? 1.f / l : 0.f) * (a.head_scale ? a.head_scale[h] : 1.f);
    __nv_bfloat16* O = (__nv_bfloat16*)a.o + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD + hf * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      reinterpret_cast<uint4*>(O)[c] =
          make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                     pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    if (hf == 0) a.lse[((size_t)b * a.H + h) * a.T + i] = (m == -CUDART_INF_F ? 0.f : m) + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(sm.tmem_addr);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward, warp-specialised (the default): one CTA = one (batch, head) and TWO 128-row query tiles sweeping 64-key tiles.
//   warps 0-7 / 8-15 : softmax of query tile 0 / 1 -- two threads per query row (TMEM lane), 32 scores each per key tile in
//                      registers (warp w: columns 0-31, warp w+4: columns 32-63 of the same rows; the row maximum is
//                      exchanged through shared memory behind a 256-thread named barrier), online softmax, P (bf16) to
//                      swizzled shared memory, O folded from the per-tile P.V partial (32 output columns per thread)
//   warp 16          : TMA producer -- Q' once, K'/V through a 3-stage ring; starts before the LUT staging of the others
//   warps 17, 18     : tcgen05 issuers, one per query tile -- S(j+1) is issued BEFORE P.V(j), so the tensor pipe computes the
//                      next scores while the softmax warps work on the current ones; every hand-over is an mbarrier
//                      (tcgen05.commit -> s_full / o_full / kv_empty; softmax arrivals -> s_free / p_full), no __syncthreads
//                      in the main loop.
// The tensor cores read an operand that lives in shared memory at ~64 B/clk (measured: a 128x64x16 MMA with both operands
// in shared memory takes ~100 cycles instead of 32, ncu r02_attn_fwd_ws_v0), so with kTS both A operands are moved to tensor
// memory: Q' is copied there once per CTA (bf16 pairs, 64 columns), P is written there by the softmax threads (tcgen05.st)
// instead of to swizzled shared memory; only K' and V are then read from shared memory.  The per-head image LUT (27.6 KB), the token LUT slice
// of the 256 query rows and the per-key metadata of ALL keys are staged in shared memory once per CTA.
// ---------------------------------------------------------------------------------------------------------------------
// -DOFA_WS_DEBUG: clock64() phase accounting of one softmax warp and of the issuer thread of one CTA (tools/attn_ws_phases.py
// reads it through ofa_attn_ws_debug_read).  Not compiled into the shipped library.
#ifdef OFA_WS_DEBUG
__device__ long long g_ws_dbg[64];
#define WSD_DECL(cond) const bool wsd = (cond); long long wsd_last = clock64();
#define WSD(i) if (wsd) { const long long now_ = clock64(); g_ws_dbg[i] += now_ - wsd_last; wsd_last = now_; }
#else
#define WSD_DECL(cond)
#define WSD(i)
#endif
constexpr int kImgLutMax = 83 * 83 + 3 + 1;
constexpr int WS_QT = 2, WS_ST = 3, WS_SOFTMAX = 512, WS_THREADS = WS_SOFTMAX + 32 * (1 + WS_QT), WS_MAXK = 2048;
// TMEM columns per query tile: S [0,64) | O partial [64,128) | (kTS) Q' as bf16 A operand [128,192) | P as bf16 A operand [192,224)
constexpr int WS_TILE_COLS = 256, WS_COL_S = 0, WS_COL_O = 64, WS_COL_Q = 128, WS_COL_P = 192;
constexpr int WS_TOK = 1024 + WS_QT * BQ;

struct WsSmem {
  uint8_t q[WS_QT][2][BQ * 128];    // [tile][q | pos_q][128 rows][128 B]
  uint8_t k[WS_ST][2][BKV * 128];   // [stage][k | pos_k][64 keys][128 B]
  uint8_t v[WS_ST][BKV * 128];      // [stage][64 keys][64 x bf16]
  uint8_t p[WS_QT][BQ * 128];       // [tile][128 rows][64 keys x bf16]
  float img_s[kImgLutMax];          // image rel-pos LUT of this head
  float tok_s[WS_TOK];              // token rel-pos LUT slice: index (row in CTA) + S_t - 1 - j_t
  alignas(16) int koff[WS_MAXK];    // image keys: 4 * (kr*(2*ibs-1)+kc) (a byte offset into img_s), else 0
  uint32_t kmask[WS_MAXK / 32];     // bit set = key masked (padding, beyond S)
  float rowmax[WS_QT][2][2][BQ];    // [tile][iteration parity][column half][row]: row-maximum exchange of the two half-row threads
  float rowsum[WS_QT][2][BQ];       // final exchange of the partial row sums
  uint64_t q_full, k_full[WS_ST], kv_empty[WS_ST], s_full[WS_QT], s_free[WS_QT], p_full[WS_QT], o_full[WS_QT], qt_full[WS_QT];
  uint32_t tmem_addr;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool kTS>   // kTS: the A operands of both GEMMs (Q', P) live in tensor memory instead of shared memory
__global__ void __launch_bounds__(WS_THREADS, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmPQ,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmPK,
                   const __grid_constant__ CUtensorMap tmV, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WsSmem& sm = *reinterpret_cast<WsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
#ifdef OFA_WS_DEBUG
  if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 5 && t == 0) g_ws_dbg[39] += clock64();
#endif
  constexpr int kProducer = WS_SOFTMAX / 32, kIssuer = kProducer + 1;
  const int q0 = blockIdx.x * (WS_QT * BQ), h = blockIdx.y, b = blockIdx.z;
  const AttnBias& bz = a.bias;
  const int ntiles_all = (a.S + BKV - 1) / BKV;
  int nt[WS_QT];
#pragma unroll
  for (int tile = 0; tile < WS_QT; ++tile) {
    const int q0t = q0 + tile * BQ;
    int n = q0t < a.T ? ntiles_all : 0;
    if (a.causal && n) n = min(n, (q0t + BQ - 1 + a.q_pos_off) / BKV + 1);   // key tiles above the diagonal are skipped
    nt[tile] = n;
  }
  const int nmax = max(nt[0], nt[1]);
  const int w83 = 2 * bz.ibs - 1;
#ifdef OFA_WS_DEBUG
  const bool wsd_cta = blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 5;
#endif

  if (warp == kProducer) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmPQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmPK);
      tma_prefetch_desc(&tmV);
      mbar_init(&sm.q_full, 1);
      for (int s = 0; s < WS_ST; ++s) { mbar_init(&sm.k_full[s], 1); mbar_init(&sm.kv_empty[s], (nt[0] > 0) + (nt[1] > 0)); }
      for (int q = 0; q < WS_QT; ++q) {
        mbar_init(&sm.s_full[q], 1); mbar_init(&sm.s_free[q], 2 * BQ); mbar_init(&sm.p_full[q], 2 * BQ);
        mbar_init(&sm.o_full[q], 1); mbar_init(&sm.qt_full[q], 2 * BQ);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<WS_QT * WS_TILE_COLS>(&sm.tmem_addr);
  }
  pdl_sync();   // everything above is on-chip set-up; what follows reads the predecessor's output
  if (warp == kProducer && lane == 0) {
    // Q' and the first K'/V stages are requested before the LUT staging below, so their latency runs under it
    uint32_t qbytes = 0;
#pragma unroll
    for (int tile = 0; tile < WS_QT; ++tile) qbytes += nt[tile] ? 2 * BQ * 128 : 0;
    mbar_expect_tx(&sm.q_full, qbytes);
#pragma unroll
    for (int tile = 0; tile < WS_QT; ++tile)
      if (nt[tile]) {
        tma_load_4d(sm.q[tile][0], &tmQ, &sm.q_full, 0, h, q0 + tile * BQ, b);
        tma_load_4d(sm.q[tile][1], &tmPQ, &sm.q_full, 0, h, q0 + tile * BQ, b);
      }
    for (int j = 0; j < min(nmax, WS_ST); ++j) {
      mbar_expect_tx(&sm.k_full[j], 3 * BKV * 128);
      tma_load_4d(sm.k[j][0], &tmK, &sm.k_full[j], 0, h, j * BKV, b);
      tma_load_4d(sm.k[j][1], &tmPK, &sm.k_full[j], 0, h, j * BKV, b);
      tma_load_4d(sm.v[j], &tmV, &sm.k_full[j], 0, h, j * BKV, b);
    }
  }
  const int S_t = a.S - bz.k_text_off;
  {
    const int Spad = ntiles_all * BKV;
    for (int j0 = 0; j0 < Spad; j0 += WS_THREADS) {
      if (j0 + warp * 32 < Spad) {   // warp-uniform
        const int j = j0 + t;
        const bool masked = j >= a.S || (a.kpm && a.kpm[(size_t)b * a.S + j]);
        int off = 0;
        if (bz.img_lut && j < bz.n_img_k) {
          const int pid = bz.k_pid[(size_t)b * bz.n_img_k + j] - 1;
          off = 4 * ((pid / bz.ibs) * w83 + (pid % bz.ibs));
        }
        sm.koff[j] = off;
        const uint32_t bal = __ballot_sync(0xffffffffu, masked);
        if (lane == 0) sm.kmask[j >> 5] = bal;
      }
    }
    if (bz.tok_lut) {
      const int i_t0 = q0 + a.q_pos_off - bz.q_text_off;
      const float* lut = bz.tok_lut + (size_t)h * (2 * bz.tok_max - 1);
      for (int e = t; e < WS_TOK; e += WS_THREADS) {
        const int rel = i_t0 + e - (S_t - 1) + bz.tok_max - 1;
        sm.tok_s[e] = (rel >= 0 && rel < 2 * bz.tok_max - 1) ? lut[rel] : 0.f;
      }
    }
    if (bz.img_lut) {
      const float4* lut4 = reinterpret_cast<const float4*>(bz.img_lut + (size_t)h * bz.n_img_rel);
      if ((((size_t)h * bz.n_img_rel) & 3) == 0 && (reinterpret_cast<uintptr_t>(bz.img_lut) & 15) == 0) {
        for (int e = t; e < (bz.n_img_rel >> 2); e += WS_THREADS) reinterpret_cast<float4*>(sm.img_s)[e] = lut4[e];
        for (int e = (bz.n_img_rel & ~3) + t; e < bz.n_img_rel; e += WS_THREADS)
          sm.img_s[e] = bz.img_lut[(size_t)h * bz.n_img_rel + e];
      } else {
        for (int e = t; e < bz.n_img_rel; e += WS_THREADS) sm.img_s[e] = bz.img_lut[(size_t)h * bz.n_img_rel + e];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_addr;
#ifdef OFA_WS_DEBUG
  if (wsd_cta && t == 0) g_ws_dbg[40] += clock64();      // (start stamps are subtracted on the host side: single CTA)
#endif

  if (warp == kProducer) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int j = WS_ST; j < nmax; ++j) {
        const int st = j % WS_ST;
        mbar_wait(&sm.kv_empty[st], ((j / WS_ST) - 1) & 1);
        mbar_expect_tx(&sm.k_full[st], 3 * BKV * 128);
        tma_load_4d(sm.k[st][0], &tmK, &sm.k_full[st], 0, h, j * BKV, b);
        tma_load_4d(sm.k[st][1], &tmPK, &sm.k_full[st], 0, h, j * BKV, b);
        tma_load_4d(sm.v[st], &tmV, &sm.k_full[st], 0, h, j * BKV, b);
      }
    }
    __syncwarp();
  } else if (warp >= kIssuer) {
    // ---------------------------------------------------------------------------------------------- tcgen05 issuers
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);
      auto issue_s = [&](int tile, int st) {
        const uint32_t tb = tm + tile * WS_TILE_COLS;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t db = umma_smem_desc(smem_u32(sm.k[st][kb]) + ks * 32, 16, 1024);
            if (kTS) umma_f16_ts(tb + WS_COL_S, tb + WS_COL_Q + kb * 32 + ks * 8, db, idesc_s, (kb | ks) != 0);
            else umma_f16(tb + WS_COL_S, umma_smem_desc(smem_u32(sm.q[tile][kb]) + ks * 32, 16, 1024), db, idesc_s, (kb | ks) != 0);
          }
        umma_commit(&sm.s_full[tile]);
      };
      auto issue_pv = [&](int tile, int st) {
        const uint32_t tb = tm + tile * WS_TILE_COLS;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t db = umma_smem_desc(smem_u32(sm.v[st]) + ks * 2048, 1024, 1024);
          if (kTS) umma_f16_ts(tb + WS_COL_O, tb + WS_COL_P + ks * 8, db, idesc_o, ks != 0);
          else umma_f16(tb + WS_COL_O, umma_smem_desc(smem_u32(sm.p[tile]) + ks * 32, 16, 1024), db, idesc_o, ks != 0);
        }
        umma_commit(&sm.o_full[tile]);
      };
      // one issuer thread per query tile (two warps): a tcgen05.mma costs its issuing thread ~80-110 cycles whatever its
      // size (phase timing, tools/attn_ws_phases.py), so the two tiles' instruction streams are issued concurrently
      const int tile = warp - kIssuer;
      const int n = nt[tile];
      WSD_DECL(wsd_cta && tile == 0)
      if (n > 0) {
        if (kTS) mbar_wait(&sm.qt_full[tile], 0);      // Q' of the tile has been copied into tensor memory
        else mbar_wait(&sm.q_full, 0);
        mbar_wait(&sm.k_full[0], 0);
        tc_fence_after();
        WSD(20)
        issue_s(tile, 0);
        WSD(21)
      }
      for (int j = 0; n > 0 && j < nmax; ++j) {
        const int st = j % WS_ST;
        // a tile with fewer key tiles than its neighbour (causal) keeps releasing the stages, in step with the producer
        if (j >= n) mbar_wait(&sm.k_full[st], (j / WS_ST) & 1);
        if (j + 1 < n) {
          const int st1 = (j + 1) % WS_ST;
          mbar_wait(&sm.k_full[st1], ((j + 1) / WS_ST) & 1);
          WSD(22)
          mbar_wait(&sm.s_free[tile], j & 1);      // the softmax warps hold S(j) in registers
          tc_fence_after();
          WSD(23)
          issue_s(tile, st1);
          WSD(24)
        }
        if (j < n) {
          mbar_wait(&sm.p_full[tile], j & 1);
          tc_fence_after();
          WSD(27)
          issue_pv(tile, st);
          WSD(28)
        }
        umma_commit(&sm.kv_empty[st]);             // (both issuers: the stage is free when both tiles' MMAs have retired)
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax warps (warp < kProducer)
    const int tile = warp >> 3;
    const int hf = (warp >> 2) & 1;         // column half of the key tile: 32 scores / 32 output columns per thread
    const int n = nt[tile];
    const int r = (warp & 3) * 32 + lane;   // row inside the tile = TMEM lane
    const int rc = tile * BQ + r;           // row inside the CTA (token-LUT indexing)
    const int i = q0 + rc, iabs = i + a.q_pos_off;
    const bool row_ok = i < a.T;
    const bool q_text = bz.tok_lut && iabs >= bz.q_text_off;
    const bool q_img = bz.img_lut && iabs < bz.n_img_q && row_ok;
    const char* img_row = reinterpret_cast<const char*>(sm.img_s);
    if (q_img) {
      const int pid = bz.q_pid[(size_t)b * bz.n_img_q + iabs] - 1;
      img_row += 4 * ((pid / bz.ibs + bz.ibs - 1) * w83 + (pid % bz.ibs + bz.ibs - 1));
    }
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_t = tm + lane_off + tile * WS_TILE_COLS;
    const uint32_t tmem_s = tmem_t + WS_COL_S + hf * 32, tmem_o = tmem_t + WS_COL_O + hf * 32, tmem_p = tmem_t + WS_COL_P + hf * 16;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -CUDART_INF_F, l = 0.f, alpha_prev = 1.f;
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    uint8_t* prow = sm.p[tile] + r * 128;
    const int sw = r & 7;
    if (kTS && n > 0) {
      // Q' -> tensor memory, once: this thread moves the 64 bf16 of its row's q (hf = 0) or pos_q (hf = 1) block (8 swizzled
      // 16-byte chunks of the TMA tile) into 32 columns of the A-operand region
      mbar_wait(&sm.q_full, 0);
      uint32_t qa[32];
      const uint8_t* qrow = sm.q[tile][hf] + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(qrow + ((c ^ sw) << 4));
        qa[4 * c] = u.x; qa[4 * c + 1] = u.y; qa[4 * c + 2] = u.z; qa[4 * c + 3] = u.w;
      }
      tmem_st32(tmem_t + WS_COL_Q + hf * 32, qa);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&sm.qt_full[tile]);
    }
    // bias class of a 32-column key block: fully image keys -> image LUT gather (image query rows) or no bias; fully text
    // keys -> token LUT (text query rows) or no bias; blocks that straddle the image / text boundary and blocks cut by the
    // causal diagonal take the generic per-element path (mode 3)
    const bool no_bias = !bz.tok_lut && !bz.img_lut;
    const int blk_img_end = no_bias ? (1 << 30) : (bz.img_lut ? bz.n_img_k / 32 : 0);
    const int blk_txt_beg = bz.tok_lut ? (max(bz.k_text_off, bz.img_lut ? bz.n_img_k : 0) + 31) / 32 : (1 << 30);
    const int mode_img = (q_img && !no_bias) ? 2 : 0, mode_txt = q_text ? 1 : 0;
    auto fold = [&](float al) {             // o = o * al + (P.V partial of the previous key tile)
      uint32_t ro[32];
      tmem_ld32(tmem_o, ro);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) o[jj] = fmaf(o[jj], al, __uint_as_float(ro[jj]));
    };
    WSD_DECL(wsd_cta && warp == 0 && lane == 0)
    for (int j = 0; j < n; ++j) {
      const int kc0 = j * BKV + hf * 32;    // first key column of this thread
      float s[32];
      WSD(0)
      mbar_wait(&sm.s_full[tile], j & 1);
      tc_fence_after();
      WSD(1)
      {
        uint32_t rr[32];
        tmem_ld32(tmem_s, rr);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sm.s_free[tile]);      // S(j) is in registers: the issuer may overwrite it with S(j+1)
        WSD(2)
        const uint32_t cm = sm.kmask[kc0 >> 5];
        const int blk = 2 * j + hf;         // 32-column block index on the key axis
        int mode = blk < blk_img_end ? mode_img : (blk >= blk_txt_beg ? mode_txt : 3);
        if (a.causal && kc0 + 31 > iabs) mode = 3;
        const int tb = rc + S_t - 1 - (kc0 - bz.k_text_off);   // tok_s index of column kc0; column jj -> tb - jj
        if (mode == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) s[jj] = __uint_as_float(rr[jj]);
        } else if (mode == 1) {
          const float* ts = sm.tok_s + tb;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) s[jj] = __uint_as_float(rr[jj]) + ts[-jj];
        } else if (mode == 2) {
          const int4* ko = reinterpret_cast<const int4*>(sm.koff + kc0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int4 kk = ko[j4];
            s[j4 * 4 + 0] = __uint_as_float(rr[j4 * 4 + 0]) + *reinterpret_cast<const float*>(img_row - kk.x);
            s[j4 * 4 + 1] = __uint_as_float(rr[j4 * 4 + 1]) + *reinterpret_cast<const float*>(img_row - kk.y);
            s[j4 * 4 + 2] = __uint_as_float(rr[j4 * 4 + 2]) + *reinterpret_cast<const float*>(img_row - kk.z);
            s[j4 * 4 + 3] = __uint_as_float(rr[j4 * 4 + 3]) + *reinterpret_cast<const float*>(img_row - kk.w);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int jg = kc0 + jj;
            float x = __uint_as_float(rr[jj]);
            if (q_text && jg >= bz.k_text_off) x += sm.tok_s[tb - jj];
            if (q_img && jg < bz.n_img_k) x += *reinterpret_cast<const float*>(img_row - sm.koff[jg]);
            if (a.causal && jg > iabs) x = -CUDART_INF_F;
            s[jj] = x;
          }
        }
        if (cm != 0) {                      // padded keys in these 32 columns (warp-uniform: the mask is per key)
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if ((cm >> jj) & 1u) s[jj] = -CUDART_INF_F;
        }
      }
      float mx4[4] = {s[0], s[1], s[2], s[3]};   // four independent chains: the row maximum is on the critical path
#pragma unroll
      for (int jj = 4; jj < 32; ++jj) mx4[jj & 3] = fmaxf(mx4[jj & 3], s[jj]);
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      sm.rowmax[tile][j & 1][hf][r] = mx;
      WSD(3)
      named_bar_sync(1 + tile, 2 * BQ);
      mx = fmaxf(mx, sm.rowmax[tile][j & 1][hf ^ 1][r]);
      WSD(4)
      const float mn = fmaxf(m, mx);
      const float mu = (mn == -CUDART_INF_F) ? 0.f : mn;
      const float alpha = ex2((m - mu) * kLog2e);
      m = mn;
      const float mneg = -mu * kLog2e;
      uint32_t pk[16];
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float p0 = ex2(fmaf(s[2 * jj], kLog2e, mneg)), p1 = ex2(fmaf(s[2 * jj + 1], kLog2e, mneg));
        rs4[jj & 3] += p0 + p1;
        pk[jj] = pack_bf16(p0, p1);
      }
      l = l * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      WSD(5)
      if (j > 0) {                          // P.V(j-1) has retired: its partial is ready and the P buffer is free
        mbar_wait(&sm.o_full[tile], (j - 1) & 1);
        tc_fence_after();
        WSD(6)
        fold(alpha_prev);
        WSD(7)
      }
      alpha_prev = alpha;
      if (kTS) {
        tmem_st16(tmem_p, pk);
        tmem_st_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(prow + (((hf * 4 + c) ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
      }
      tc_fence_before();
      mbar_arrive(&sm.p_full[tile]);
      WSD(8)
    }
#ifdef OFA_WS_DEBUG
    if (wsd) g_ws_dbg[41] += clock64();
#endif
    if (n > 0) {
      mbar_wait(&sm.o_full[tile], (n - 1) & 1);
      tc_fence_after();
      fold(alpha_prev);
      // total row sum = the two half-row partial sums (same running maximum on both sides)
      sm.rowsum[tile][hf][r] = l;
      named_bar_sync(1 + tile, 2 * BQ);
      l += sm.rowsum[tile][hf ^ 1][r];
    }
    if (row_ok) {
      const float inv = (l > 0.f ? 1.f / l : 0.f) * (a.head_scale ? a.head_scale[h] : 1.f);
      __nv_bfloat16* O = (__nv_bfloat16*)a.o + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD + hf * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        reinterpret_cast<uint4*>(O)[c] =
            make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                       pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
      if (hf == 0) a.lse[((size_t)b * a.H + h) * a.T + i] = (m == -CUDART_INF_F ? 0.f : m) + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef OFA_WS_DEBUG
  if (wsd_cta && t == 0) g_ws_dbg[42] += clock64();
#endif
  if (warp == kProducer) {
    tc_fence_after();
    tmem_dealloc<WS_QT * WS_TILE_COLS>(tm);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
// One CTA = one (batch, head, 128-key tile); K', V stay in shared memory, dK' and dV accumulate in TMEM over the sweep of
// 128-query tiles.  Per (q-tile, k-tile) pair five tcgen05 GEMMs:
//   S = Q'K'^T, dP = dO V^T  ->  P = exp(S + bias - lse), dS = P o (c dP - delta)  (registers, 2 threads per query row)
//   dV += P^T dO, dK' += dS^T Q', dQ'_partial = dS K'   (P / dS staged as bf16 in swizzled smem; the transposed uses
//   read the same bytes through MN-major descriptors).  dQ' partials are staged in shared memory (the idle P / dS
//   buffers) and added into an fp32 accumulator by TMA reduce (cp.reduce.async.bulk.tensor .add: four 16 KB bulk
//   operations per tile pair instead of 4096 red.global.add.v4 per CTA, which saturated the L2 atomic units);
//   relative-position table gradients are privatised in shared-memory histograms.
// -DOFA_ATTN_DEBUG: clock64() stamps of thread 0 of one CTA around every phase of the backward main loop, per query tile
// (tools/attn_phase_debug.py reads them through ofa_attn_debug_read).  Not compiled into the shipped library.
#ifdef OFA_ATTN_DEBUG
__device__ long long g_attn_dbg[128];
#define DBG_T(i) if (dbg && t == 0) { const long long now_ = clock64(); g_attn_dbg[(it & 7) * 12 + (i)] += now_ - dbg_last; dbg_last = now_; }
#else
#define DBG_T(i)
#endif
constexpr int BK2 = 128;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int kBwdSoft = 512;      // softmax / histogram / drain threads: 4 per query row, 32 key columns each
constexpr int kBwdThreads = kBwdSoft + 128;  // + one more warpgroup: its first warp (16) is the TMA producer and tcgen05 issuer
                                             // (one elected lane); a whole warpgroup because setmaxnreg works per warpgroup
constexpr int kTokHist = 1024 + 128;
constexpr int kImgHistMax = 83 * 83 + 3;

struct BwdSmem {
  uint8_t k[2][BK2 * 128];   // [k | pos_k][128 keys][128 B]
  uint8_t v[BK2 * 128];      // [128 keys][64 x bf16]
  uint8_t q[2][BQ * 128];    // [q | pos_q][128 rows][128 B]
  uint8_t dout[BQ * 128];    // [128 rows][64 x bf16]
  uint8_t p[2][BQ * 128];    // [64-key half][128 rows][128 B]
  uint8_t ds[2][BQ * 128];
  // relative-position table gradients as FIXED-POINT sums: value * 2^(kq - E) in int32, E = the CTA's running block
  // exponent.  Shared-memory float atomics are compare-and-swap loops on sm_100 (5.5 us per 128x128 tile, measured:
  // tools/scratch/atom_probe.cu); int32 adds are native ATOMS.ADD (0.9 us) and order-independent.
  int hist_tok[kTokHist];    // d tok_lut, indexed (i_t - jl) + 127
  float tok_s[kTokHist];     // tok_lut staged with the same indexing (valid when all keys of the tile are text)
  int hist_img[kImgHistMax + 1];
  alignas(16) uint32_t wmax[16];   // per-warp max |dS| of the current tile (float bits)
  int kinfo[BK2];            // bit31 masked | bit30 image key | [0,16) kr*(2*ibs-1)+kc
  float img_s[kImgHistMax + 1];   // image rel-pos LUT of this head (the per-element gather went to L1 / L2 before: long-scoreboard
                                  // stalls were the top stall reason of the round-1 capture)
  uint64_t bar_kv, bar_q, bar_sp, bar_dq, bar_qf;
  uint64_t bar_s;        // S of the tile is in tensor memory (bar_sp: dP too)
  uint64_t bar_pds;      // P / dS of the tile are in shared memory (softmax threads -> issuer)
  uint64_t bar_drain;    // dQ' of the tile has left tensor memory (softmax threads -> issuer: dP of the next tile may land there)
  uint64_t bar_slab;     // the dQ' reduce has finished reading the staging slabs (thread 0 -> softmax threads: P / dS may be written)
  uint32_t tmem_addr;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmPQ,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmPK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQ /* dq_acc [B,T,H,128] fp32, box 32 x 128 rows */, AttnArgs a,
                   AttnGrads g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int t = threadIdx.x, warp = t >> 5;
#ifdef OFA_ATTN_DEBUG
  const long long dbg_t0 = clock64();
#endif
  const int r = t & 127, qd = t >> 7;   // TMEM lane / query row inside the tile ; 32-column quarter of the key tile
  const int hf = qd >> 1, ch = qd & 1;  // 64-key half of the P / dS staging tiles, 32-column chunk inside it
  const int k0 = blockIdx.x * BK2, h = blockIdx.y, b = blockIdx.z;
  const AttnBias& bz = a.bias;
  const bool has_tok = bz.tok_lut != nullptr && g.dtok_lut != nullptr;
  const bool has_img = bz.img_lut != nullptr && g.dimg_lut != nullptr;
  const int nq_tiles = (a.T + BQ - 1) / BQ;
  const int w83 = 2 * bz.ibs - 1;
  // causal: query rows i with i + q_pos_off < k0 see none of this tile's keys
  int qt0 = 0;
  if (a.causal) { const int first = k0 - a.q_pos_off; qt0 = first > 0 ? first / BQ : 0; }

  if (t == kBwdSoft) {     // the producer / issuer thread initialises the barriers it is about to use
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmPQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmPK);
    tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ);
    mbar_init(&sm.bar_kv, 1); mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_sp, 1); mbar_init(&sm.bar_dq, 1); mbar_init(&sm.bar_qf, 1);
    mbar_init(&sm.bar_pds, 1); mbar_init(&sm.bar_drain, 1); mbar_init(&sm.bar_slab, 1); mbar_init(&sm.bar_s, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&sm.tmem_addr);
  pdl_sync();
  if (qt0 < nq_tiles && t == kBwdSoft) {     // first loads: in flight while the LUTs and the key metadata are staged below
    mbar_expect_tx(&sm.bar_kv, 3 * BK2 * 128);
    tma_load_4d(sm.k[0], &tmK, &sm.bar_kv, 0, h, k0, b);
    tma_load_4d(sm.k[1], &tmPK, &sm.bar_kv, 0, h, k0, b);
    tma_load_4d(sm.v, &tmV, &sm.bar_kv, 0, h, k0, b);
    mbar_expect_tx(&sm.bar_q, 3 * BQ * 128);
    tma_load_4d(sm.q[0], &tmQ, &sm.bar_q, 0, h, qt0 * BQ, b);
    tma_load_4d(sm.q[1], &tmPQ, &sm.bar_q, 0, h, qt0 * BQ, b);
    tma_load_4d(sm.dout, &tmDO, &sm.bar_q, 0, h, qt0 * BQ, b);
  }
  // tile classification: the keys are stationary per CTA
  const bool keys_all_img = bz.img_lut != nullptr && (k0 + BK2 <= bz.n_img_k);
  const bool keys_all_txt = bz.tok_lut != nullptr && (k0 >= bz.k_text_off) && (bz.img_lut == nullptr || k0 >= bz.n_img_k);
  const int tok_base = k0 - bz.k_text_off + 127;   // (i_t - j_t) = u - tok_base with u = i_t - jl + 127
  for (int e = t; e < kTokHist; e += kBwdThreads) {
    sm.hist_tok[e] = 0;
    float lv = 0.f;
    if (bz.tok_lut) {
      const int rel = e - tok_base + bz.tok_max - 1;
      if (rel >= 0 && rel < 2 * bz.tok_max - 1) lv = bz.tok_lut[(size_t)h * (2 * bz.tok_max - 1) + rel];
    }
    sm.tok_s[e] = lv * kLog2e;         // (the LUTs are staged pre-multiplied by log2(e): p = 2^(S log2e + lut' - lse'))
  }
  for (int e = t; e < kImgHistMax + 1; e += kBwdThreads) sm.hist_img[e] = 0;
  if (bz.img_lut) {
    const float* lut = bz.img_lut + (size_t)h * bz.n_img_rel;
    for (int e = t; e < bz.n_img_rel; e += kBwdThreads) sm.img_s[e] = lut[e] * kLog2e;
  }
  int my_masked = 0;
  if (t < BK2) {
    const int j = k0 + t;
    int info = 0;
    if (j >= a.S || (a.kpm && a.kpm[(size_t)b * a.S + j])) { info |= (int)0x80000000u; my_masked = 1; }
    if (bz.img_lut && j < bz.n_img_k) {
      const int pid = bz.k_pid[(size_t)b * bz.n_img_k + j] - 1;
      info |= 0x40000000 | ((pid / bz.ibs) * w83 + (pid % bz.ibs));
    }
    sm.kinfo[t] = info;
  }
  tc_fence_before();
  const bool keys_any_masked = __syncthreads_or(my_masked) != 0;
  tc_fence_after();
  const uint32_t tm = sm.tmem_addr;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 384;
  // masked keys (padding / beyond S) of this thread's 32 columns: the keys are stationary, so one bit mask serves every
  // query tile and the fast softmax paths stay usable on tiles that contain padded keys
  uint32_t colmask = 0;
  if (keys_any_masked && t < kBwdSoft) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) colmask |= (sm.kinfo[qd * 32 + jj] < 0 ? 1u : 0u) << jj;
  }

  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  const float* img_lut = sm.img_s;
  constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);    // S, dP
  constexpr uint32_t id_dv = umma_idesc_bf16(128, 64, 1, 1);    // dV  = P^T dO
  constexpr uint32_t id_dk = umma_idesc_bf16(128, 128, 1, 1);   // dK' = dS^T Q'
  constexpr uint32_t id_dq = umma_idesc_bf16(128, 128, 0, 1);   // dQ' = dS K'
  const int col0 = qd * 32;

#ifdef OFA_ATTN_DEBUG
  const bool dbg = blockIdx.x == 2 && blockIdx.y == 3 && blockIdx.z == 1;
  long long dbg_last = clock64();
  if (dbg && t == 0) g_attn_dbg[96] = dbg_last - dbg_t0;            // prologue (LUT staging, TMEM allocation, key metadata)
#endif
  // fixed-point histograms: kq fractional bits below the block exponent; a bin receives at most 128 values of magnitude
  // <= 2^kq per tile, so 2^(kq + 7) * tiles stays below 2^30
  const bool has_hist = has_tok || has_img;
  int kq = 23;
  for (int n = 1; n < nq_tiles - qt0; n <<= 1) --kq;
  int e_cur = 0;                                   // biased exponent E + 126 of the current scale; 0 = tables still empty
  int it = 0;
  // register budget: 640 threads launch with 96 registers each; the issuer warpgroup hands most of its share to the four
  // softmax warpgroups (the pool is what the CTA launched with: 640 x 96 = 4 x 128 x 112 + 128 x 32)
  if (t >= kBwdSoft) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp != kBwdSoft / 32) return;
    // ---- warp 16: TMA producer + tcgen05 issuer.  Runs ahead of the 16 softmax warps: S of the next query tile is issued as
    // soon as its Q' has landed (under the histogram / dQ' drain of the current one), MMA issue no longer delays warp 0's share
    // of the softmax and histogram work, every hand-over is an mbarrier.  dQ' accumulates in the dP columns (dP is consumed
    // by then), so S never waits for the drain.
    if ((t & 31) == 0) {
      for (int qt = qt0; qt < nq_tiles; ++qt, ++it) {
        const uint32_t ph = it & 1;
        const int q0 = qt * BQ;
        if (it == 0) mbar_wait(&sm.bar_kv, 0);
        mbar_wait(&sm.bar_q, ph);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tm + COL_S, umma_smem_desc(smem_u32(sm.q[kb]) + ks * 32, 16, 1024),
                     umma_smem_desc(smem_u32(sm.k[kb]) + ks * 32, 16, 1024), id_s, (kb | ks) != 0);
        umma_commit(&sm.bar_s);
        if (it > 0) {                      // dQ' of the previous tile has been read out of the dP columns
          mbar_wait(&sm.bar_drain, (it - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tm + COL_DP, umma_smem_desc(smem_u32(sm.dout) + ks * 32, 16, 1024),
                   umma_smem_desc(smem_u32(sm.v) + ks * 32, 16, 1024), id_s, ks != 0);
        umma_commit(&sm.bar_sp);
        mbar_wait(&sm.bar_pds, ph);        // P / dS staged by the softmax threads
        tc_fence_after();
        const uint32_t acc = it != 0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dV[keys, hd] += P^T dO   (contraction over the 128 query rows, 16 per MMA)
          umma_f16(tm + COL_DV, umma_smem_desc(smem_u32(sm.p[0]) + ks * 2048, BQ * 128, 1024),
                   umma_smem_desc(smem_u32(sm.dout) + ks * 2048, 1024, 1024), id_dv, acc | (ks != 0));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dK'[keys, 128] += dS^T Q'
          umma_f16(tm + COL_DK, umma_smem_desc(smem_u32(sm.ds[0]) + ks * 2048, BQ * 128, 1024),
                   umma_smem_desc(smem_u32(sm.q[0]) + ks * 2048, BQ * 128, 1024), id_dk, acc | (ks != 0));
        umma_commit(&sm.bar_qf);         // dV and dK' retired = Q' and dO are no longer read
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dQ'[rows, 128] = dS K'      (contraction over the 128 keys)
          umma_f16(tm + COL_DP, umma_smem_desc(smem_u32(sm.ds[ks >> 2]) + (ks & 3) * 32, 16, 1024),
                   umma_smem_desc(smem_u32(sm.k[0]) + ks * 2048, BK2 * 128, 1024), id_dq, ks != 0);
        umma_commit(&sm.bar_dq);
        if (qt + 1 < nq_tiles) {
          // reload Q' / dO for the next query tile as soon as dV / dK' are done: the TMA latency (~2 us) runs under the dQ'
          // GEMM, the histogram phase and the drain
          mbar_wait(&sm.bar_qf, ph);
          tc_fence_after();
          mbar_expect_tx(&sm.bar_q, 3 * BQ * 128);
          tma_load_4d(sm.q[0], &tmQ, &sm.bar_q, 0, h, q0 + BQ, b);
          tma_load_4d(sm.q[1], &tmPQ, &sm.bar_q, 0, h, q0 + BQ, b);
          tma_load_4d(sm.dout, &tmDO, &sm.bar_q, 0, h, q0 + BQ, b);
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
  float lse_nx = 0.f, delta_nx = 0.f;
  int pid_nx = 1;
  auto row_meta = [&](int i) {
    const int iabs = i + a.q_pos_off;
    const size_t ridx = ((size_t)b * a.H + h) * a.T + (i < a.T ? i : 0);
    lse_nx = a.lse[ridx];
    delta_nx = g.delta[ridx];
    if (bz.img_lut && iabs < bz.n_img_q && i < a.T) pid_nx = bz.q_pid[(size_t)b * bz.n_img_q + iabs];
  };
  if (qt0 < nq_tiles) row_meta(qt0 * BQ + r);
  for (int qt = qt0; qt < nq_tiles; ++qt, ++it) {
    const uint32_t ph = it & 1;
    const int q0 = qt * BQ;
    DBG_T(0)
    // per-row metadata
    const int i = q0 + r;
    const int iabs = i + a.q_pos_off;
    const bool row_ok = i < a.T;
    // lse / delta / position id of this row were requested one tile ago; the next tile's are requested now
    const float lse = lse_nx, delta = delta_nx;
    const bool q_text = bz.tok_lut && iabs >= bz.q_text_off;
    const bool q_img = bz.img_lut && iabs < bz.n_img_q && row_ok;
    int rowbase = 0;
    if (q_img) {
      const int pid = pid_nx - 1;
      rowbase = (pid / bz.ibs + bz.ibs - 1) * w83 + (pid % bz.ibs + bz.ibs - 1);
    }
    if (qt + 1 < nq_tiles) row_meta(q0 + BQ + r);
    const int i_t = iabs - bz.q_text_off;
    const int tu = i_t + 127 - col0;   // tok_s / hist_tok index of column col0 for this row; column jj -> tu - jj
    // fast paths need: no masked key in the tile, a valid row, and no causal cut inside this thread's 32 columns
    const bool plain = row_ok && !(a.causal && k0 + col0 + 31 > iabs);
    const uint32_t cm = row_ok ? colmask : 0xffffffffu; // rows beyond T: the plain path with every column masked
    int mode = row_ok ? 3 : 0;                          // 3 = generic per-element path
    if (plain) {
      if (q_text && keys_all_txt) mode = 1;             // text x text: token LUT from shared memory
      else if (q_img && keys_all_img) mode = 2;         // image x image: image LUT gather
      else if ((keys_all_txt && !q_text) || (keys_all_img && !q_img) || (!bz.tok_lut && !bz.img_lut)) mode = 0;
    }

    DBG_T(1)
    mbar_wait(&sm.bar_s, ph);          // S is issued ahead of dP (which waits for the dQ' drain): the probabilities -- the exp and
    tc_fence_after();                  // LUT part of this phase -- are computed while dP is still on its way
    DBG_T(2)
    float dsv[32];
    {
      uint32_t rs[32];
      tmem_ld32(tm + lane_off + COL_S + col0, rs);
      tmem_ld_wait();
      float pv[32];
      const float nl2 = -lse * kLog2e;
      auto probs = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
        if (mode == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            float p = ex2_approx(fmaf(__uint_as_float(rs[jj]), kLog2e, nl2));
            if (kMasked && (cm & (1u << jj))) p = 0.f;
            pv[jj] = p;
          }
        } else if (mode == 1) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            float p = ex2_approx(fmaf(__uint_as_float(rs[jj]), kLog2e, sm.tok_s[tu - jj] + nl2));
            if (kMasked && (cm & (1u << jj))) p = 0.f;
            pv[jj] = p;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int idx = rowbase - (sm.kinfo[col0 + jj] & 0xffff);
            float p = ex2_approx(fmaf(__uint_as_float(rs[jj]), kLog2e, img_lut[idx] + nl2));
            if (kMasked && (cm & (1u << jj))) p = 0.f;
            pv[jj] = p;
          }
        }
      };
      if (mode == 3) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int jl = col0 + jj, j = k0 + jl;
          const int info = sm.kinfo[jl];
          float x = nl2;
          const bool tok_el = q_text && j >= bz.k_text_off;
          if (tok_el) x += sm.tok_s[tu - jj];
          if (q_img && (info & 0x40000000)) x += img_lut[rowbase - (info & 0xffff)];
          const bool masked = info < 0 || (a.causal && j > iabs) || !row_ok;
          pv[jj] = masked ? 0.f : ex2_approx(fmaf(__uint_as_float(rs[jj]), kLog2e, x));
        }
      } else if (cm == 0) {
        probs(std::false_type{});
      } else {
        probs(std::true_type{});
      }
      mbar_wait(&sm.bar_sp, ph);       // dP
      tc_fence_after();
      {
        uint32_t rp[32];
        tmem_ld32(tm + lane_off + COL_DP + col0, rp);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) dsv[jj] = pv[jj] * fmaf(__uint_as_float(rp[jj]), cs, -delta);
      }
      if (it > 0) {       // the previous tile's dQ' reduce must have finished READING the P / dS buffers (its staging slabs)
        if (t == 0) { tma_store_wait_read<0>(); mbar_arrive(&sm.bar_slab); }
        mbar_wait(&sm.bar_slab, (it - 1) & 1);
      }
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {
        const int chunk = ch * 4 + c16;   // 16-byte chunk inside the 128-byte row of this half
        const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sm.p[hf] + off) =
            make_uint4(pack_bf16(pv[c16 * 8], pv[c16 * 8 + 1]), pack_bf16(pv[c16 * 8 + 2], pv[c16 * 8 + 3]),
                       pack_bf16(pv[c16 * 8 + 4], pv[c16 * 8 + 5]), pack_bf16(pv[c16 * 8 + 6], pv[c16 * 8 + 7]));
        *reinterpret_cast<uint4*>(sm.ds[hf] + off) =
            make_uint4(pack_bf16(dsv[c16 * 8], dsv[c16 * 8 + 1]), pack_bf16(dsv[c16 * 8 + 2], dsv[c16 * 8 + 3]),
                       pack_bf16(dsv[c16 * 8 + 4], dsv[c16 * 8 + 5]), pack_bf16(dsv[c16 * 8 + 6], dsv[c16 * 8 + 7]));
      }
    }
    if (has_hist) {
      float m = 0.f;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) m = fmaxf(m, fabsf(dsv[jj]));
      uint32_t mb = __float_as_uint(m);
      if (mb >= 0x7f800000u) mb = 0x7f7fffffu;      // inf / nan: saturate the scale, the conversion below saturates too
      const uint32_t wm = __reduce_max_sync(0xffffffffu, mb);   // non-negative floats order like their bit patterns
      if ((t & 31) == 0) sm.wmax[warp] = wm;
    }
    DBG_T(3)
    fence_proxy_async();
    tc_fence_before();
    named_bar_sync(1, kBwdSoft);
    DBG_T(4)
    if (t == 0) mbar_arrive(&sm.bar_pds);      // warp 16 issues dV / dK' / dQ' while the histogram below runs
    DBG_T(9)
    // relative-position table gradients: shared-memory histogram updates run here, under the three tensor-core GEMMs just
    // issued.  Block exponent: every thread derives the tile's max |dS| from the 16 warp maxima; when it outgrows the
    // current scale the tables are shifted down first (rare: the scale only ever grows, by at least 2 bits a time).
    float qscale = 0.f;
    if (has_hist) {
      uint32_t mb = 0;
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const uint4 u = reinterpret_cast<const uint4*>(sm.wmax)[w4];
        mb = max(max(mb, u.x), max(max(u.y, u.z), u.w));
      }
      int e_t = (int)(mb >> 23);                    // |dS| < 2^(e_t - 126) for every element of the tile
      if (e_t < 40) e_t = 40;                       // all-zero / denormal-scale tiles: any scale will do
      if (e_t > e_cur) {
        const int e_new = e_t + 1;
        if (e_cur != 0) {                           // CTA-uniform branch
          const int sh = e_new - e_cur;
          for (int e = t; e < kTokHist + kImgHistMax + 1; e += kBwdSoft) {
            int* bin = e < kTokHist ? &sm.hist_tok[e] : &sm.hist_img[e - kTokHist];
            const int v = *bin;
            if (v != 0) *bin = sh >= 31 ? 0 : (v + (1 << (sh - 1))) >> sh;
          }
          named_bar_sync(1, kBwdSoft);
        }
        e_cur = e_new;
      }
      qscale = __uint_as_float((uint32_t)(253 + kq - e_cur) << 23);      // 2^(kq - (e_cur - 126))
    }
    if (mode == 1) {
      if (has_tok) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) atomicAdd(&sm.hist_tok[tu - jj], __float2int_rn(dsv[jj] * qscale));
      }
    } else if (mode == 2) {
      if (has_img) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          atomicAdd(&sm.hist_img[rowbase - (sm.kinfo[col0 + jj] & 0xffff)], __float2int_rn(dsv[jj] * qscale));
      }
    } else if (mode == 3 && has_hist) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int jl = col0 + jj;
        const int info = sm.kinfo[jl];
        if (dsv[jj] != 0.f) {       // masked elements carry p = 0
          const int qv = __float2int_rn(dsv[jj] * qscale);
          if (has_tok && q_text && k0 + jl >= bz.k_text_off) atomicAdd(&sm.hist_tok[tu - jj], qv);
          if (has_img && q_img && (info & 0x40000000)) atomicAdd(&sm.hist_img[rowbase - (info & 0xffff)], qv);
        }
      }
    }
    DBG_T(5)
    mbar_wait(&sm.bar_dq, ph);
    tc_fence_after();
    DBG_T(6)
    {
      // dQ' partial [128 rows][32 columns of this quarter] -> fp32 slab qd (128B-swizzled rows) in the P / dS buffers, which
      // are idle until the next tile's softmax phase; one TMA reduce per slab adds it into dq_acc (rows >= T are clipped)
      uint32_t rq[32];
      tmem_ld32(tm + lane_off + COL_DP + col0, rq);
      tmem_ld_wait();
      uint8_t* slab = sm.p[0] + qd * (BQ * 128);
#pragma unroll
      for (int v4 = 0; v4 < 8; ++v4)
        *reinterpret_cast<uint4*>(slab + r * 128 + ((v4 ^ (r & 7)) << 4)) =
            make_uint4(rq[v4 * 4], rq[v4 * 4 + 1], rq[v4 * 4 + 2], rq[v4 * 4 + 3]);
    }
    DBG_T(7)
    fence_proxy_async();
    tc_fence_before();
    named_bar_sync(1, kBwdSoft);
    DBG_T(8)
    if (t == 0) {
      mbar_arrive(&sm.bar_drain);
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) tma_reduce_add_4d(&tmDQ, sm.p[0] + s4 * (BQ * 128), s4 * 32, h, q0, b);
      tma_store_commit();
    }
  }
  if (t == 0) tma_store_wait_read<0>();
#ifdef OFA_ATTN_DEBUG
  const long long dbg_t1 = clock64();
  if (dbg && t == 0) g_attn_dbg[97] = dbg_t1 - dbg_t0;              // prologue + main loop
#endif

  // epilogue: dK' (quarters 0-1: k part, 2-3: pos_k part), dV (16 columns per quarter), histograms
  const int j = k0 + r;
  if (it > 0) {
    tc_fence_after();
    {
      uint32_t rk[32];
      tmem_ld32(tm + lane_off + COL_DK + col0, rk);
      tmem_ld_wait();
      if (j < a.S) {
        __nv_bfloat16* dst = qd < 2 ? (__nv_bfloat16*)g.dk + (size_t)b * g.bsdk + (size_t)j * g.lddk + h * HD + qd * 32
                                    : (__nv_bfloat16*)g.dpk + (size_t)b * g.bsdpk + (size_t)j * g.lddpk + h * HD + (qd - 2) * 32;
        const bool addto = g.acc_pos && qd >= 2;         // pos_k part: summed over the layers that share pos_k
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(rk[8 * v4 + e]);
          if (addto) {
            const uint4 old = reinterpret_cast<const uint4*>(dst)[v4];
            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 of = __bfloat1622float2(oh[e]); f[2 * e] += of.x; f[2 * e + 1] += of.y; }
          }
          reinterpret_cast<uint4*>(dst)[v4] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                         pack_bf16(f[6], f[7]));
        }
      }
    }
    {
      uint32_t rv[16];
      tmem_ld16(tm + lane_off + COL_DV + qd * 16, rv);
      tmem_ld_wait();
      if (j < a.S) {
        __nv_bfloat16* dst = (__nv_bfloat16*)g.dv + (size_t)b * g.bsdv + (size_t)j * g.lddv + h * HD + qd * 16;
#pragma unroll
        for (int v4 = 0; v4 < 2; ++v4)
          reinterpret_cast<uint4*>(dst)[v4] = make_uint4(
              pack_bf16(__uint_as_float(rv[8 * v4]) * cs, __uint_as_float(rv[8 * v4 + 1]) * cs),
              pack_bf16(__uint_as_float(rv[8 * v4 + 2]) * cs, __uint_as_float(rv[8 * v4 + 3]) * cs),
              pack_bf16(__uint_as_float(rv[8 * v4 + 4]) * cs, __uint_as_float(rv[8 * v4 + 5]) * cs),
              pack_bf16(__uint_as_float(rv[8 * v4 + 6]) * cs, __uint_as_float(rv[8 * v4 + 7]) * cs));
      }
    }
  } else if (j < a.S) {   // causal tile with no visible query rows: zero gradients
    const uint4 z = make_uint4(0, 0, 0, 0);
    __nv_bfloat16* d0 = qd < 2 ? (__nv_bfloat16*)g.dk + (size_t)b * g.bsdk + (size_t)j * g.lddk + h * HD + qd * 32
                               : (__nv_bfloat16*)g.dpk + (size_t)b * g.bsdpk + (size_t)j * g.lddpk + h * HD + (qd - 2) * 32;
    if (!(g.acc_pos && qd >= 2))          // (an accumulated pos_k gradient receives nothing from this tile)
      for (int v4 = 0; v4 < 4; ++v4) reinterpret_cast<uint4*>(d0)[v4] = z;
    __nv_bfloat16* d1 = (__nv_bfloat16*)g.dv + (size_t)b * g.bsdv + (size_t)j * g.lddv + h * HD + qd * 16;
    for (int v4 = 0; v4 < 2; ++v4) reinterpret_cast<uint4*>(d1)[v4] = z;
  }
  named_bar_sync(1, kBwdSoft);
  const float unq = e_cur != 0 ? __uint_as_float((uint32_t)(e_cur + 1 - kq) << 23) : 0.f;      // 2^((e_cur - 126) - kq)
  if (has_tok) {
    float* gt = g.dtok_lut + (size_t)h * (2 * bz.tok_max - 1);
    for (int e = t; e < kTokHist; e += kBwdSoft) {
      const int v = sm.hist_tok[e];
      const int rel = e - tok_base + bz.tok_max - 1;
      if (v != 0 && rel >= 0 && rel < 2 * bz.tok_max - 1) atomicAdd(gt + rel, (float)v * unq);
    }
  }
  if (has_img) {
    float* gi = g.dimg_lut + (size_t)h * bz.n_img_rel;
    for (int e = t; e < bz.n_img_rel && e < kImgHistMax; e += kBwdSoft) {
      const int v = sm.hist_img[e];
      if (v != 0) atomicAdd(gi + e, (float)v * unq);
    }
  }
  tc_fence_before();
  named_bar_sync(1, kBwdSoft);
#ifdef OFA_ATTN_DEBUG
  if (dbg && t == 0) g_attn_dbg[98] = clock64() - dbg_t1;           // epilogue (dK' / dV stores, histogram flush)
#endif
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

// delta[b,h,i] = sum_d dOut . Out   (one warp per (b, i, h))
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                                             long long ldo, long long bso, int B, int T, int H,
                                                             float* __restrict__ delta) {
  // one thread per 16-byte chunk (8 of the 64 head dims) of a row: a warp reads 512 contiguous bytes of each tensor (4 heads),
  // the 8 lanes of a head add up by shuffles.  (One warp per (row, head) with 4-byte loads ran at 0.76 TB/s: 216 us per launch
  // at the merged encoder shape.)
  pdl_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * H * 8;
  const bool ok = idx < total;
  const long long c = ok ? idx : 0;
  const int ch = (int)(c % (H * 8));
  const long long row = c / (H * 8);
  const int i = (int)(row % T), b = (int)(row / T);
  const size_t off = (size_t)b * bso + (size_t)i * ldo + (size_t)ch * 8;
  const uint4 x = *reinterpret_cast<const uint4*>(dout + off), y = *reinterpret_cast<const uint4*>(out + off);
  const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
  const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a2 = __bfloat1622float2(xh[e]), b2 = __bfloat1622float2(yh[e]);
    s = fmaf(a2.x, b2.x, s);
    s = fmaf(a2.y, b2.y, s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && (threadIdx.x & 7) == 0) delta[((size_t)b * H + ch / 8) * T + i] = s;
}

// dq_acc [B,T,H,128] fp32 -> dq, dpq [B,T,H*64] bf16
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq,
                                           __nv_bfloat16* __restrict__ dpq, long long lddq, long long bsdq,
                                           long long lddpq, long long bsdpq, int B, int T, int H, float dq_scale, int acc_pos) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 elements
  const long long total = (long long)B * T * H * 16;
  if (idx >= total) return;
  const int c8 = idx % 16, h = (idx / 16) % H, i = (idx / (16 * H)) % T, b = idx / (16LL * H * T);
  float4 x = reinterpret_cast<const float4*>(acc)[idx * 2], y = reinterpret_cast<const float4*>(acc)[idx * 2 + 1];
  if (c8 < 8) { x.x *= dq_scale; x.y *= dq_scale; x.z *= dq_scale; x.w *= dq_scale; y.x *= dq_scale; y.y *= dq_scale; y.z *= dq_scale; y.w *= dq_scale; }
  if (c8 >= 8 && acc_pos) {               // pos_q part: summed over the layers that share pos_q
    const uint4 old = *reinterpret_cast<const uint4*>(dpq + (size_t)b * bsdpq + (size_t)i * lddpq + h * HD + (c8 - 8) * 8);
    const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
    const float2 o0 = __bfloat1622float2(oh[0]), o1 = __bfloat1622float2(oh[1]), o2 = __bfloat1622float2(oh[2]), o3 = __bfloat1622float2(oh[3]);
    x.x += o0.x; x.y += o0.y; x.z += o1.x; x.w += o1.y; y.x += o2.x; y.y += o2.y; y.z += o3.x; y.w += o3.y;
  }
  const uint4 pk = make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
  if (c8 < 8) *reinterpret_cast<uint4*>(dq + (size_t)b * bsdq + (size_t)i * lddq + h * HD + c8 * 8) = pk;
  else *reinterpret_cast<uint4*>(dpq + (size_t)b * bsdpq + (size_t)i * lddpq + h * HD + (c8 - 8) * 8) = pk;
}

int make_qkv_tmap(CUtensorMap* tm, const void* p, int L, int H, int B, long long ld, long long bs, int box_rows) {
  // dims {64 (head dim), H, L, B}; box {64, 1, box_rows, 1}
  uint64_t dims[4] = {(uint64_t)HD, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)ld * 2, (uint64_t)bs * 2};
  uint32_t box[4] = {HD, 1, (uint32_t)box_rows, 1};
  return ofa_make_tmap(tm, p, 4, dims, strides, box, 1, 2);
}

}  // namespace

static int g_attn_bwd_small = 1;      // A/B switch: short-query backward kernel (ofa_attn_set_bwd_small)
extern "C" int ofa_attn_set_bwd_small(int enabled) {
  const int old = g_attn_bwd_small;
  g_attn_bwd_small = enabled;
  return old;
}
static int g_attn_fwd_ws = 1;
/* A/B switch: 1 = warp-specialised forward with the A operands in tensor memory (default), 2 = warp-specialised with every
 * operand in shared memory, 0 = the single-role round-1 kernel; returns the previous setting */
extern "C" int ofa_attn_set_fwd_ws(int enabled) {
  const int old = g_attn_fwd_ws;
  g_attn_fwd_ws = enabled;
  return old;
}
#ifdef OFA_WS_DEBUG
extern "C" int ofa_attn_ws_debug_read(long long* out) {
  cudaMemcpyFromSymbol(out, g_ws_dbg, sizeof(long long) * 64);
  long long z[64] = {0};
  cudaMemcpyToSymbol(g_ws_dbg, z, sizeof(z));
  return 0;
}
#endif
#ifdef OFA_ATTN_DEBUG
extern "C" int ofa_attn_debug_read(long long* out) {
  cudaMemcpyFromSymbol(out, g_attn_dbg, sizeof(long long) * 128);
  long long z[128] = {0};
  cudaMemcpyToSymbol(g_attn_dbg, z, sizeof(z));
  return 0;
}
#endif
extern "C" int ofa_attn_fwd_tc(const AttnArgs* a, void* stream) {
  OFA_CHECK(a->T > 0 && a->S > 0 && a->B > 0 && a->H > 0, "ofa_attn_fwd_tc: empty problem");
  OFA_CHECK(a->pq && a->pk, "ofa_attn_fwd_tc: the absolute-position operands pq/pk are required");
  OFA_CHECK(a->ldq % 8 == 0 && a->ldpq % 8 == 0 && a->ldk % 8 == 0 && a->ldpk % 8 == 0 && a->ldv % 8 == 0 &&
                a->ldo % 8 == 0 && a->bso % 8 == 0,
            "ofa_attn_fwd_tc: strides must be multiples of 8 elements");
  OFA_CHECK(a->bias.ibs < 256, "ofa_attn_fwd_tc: image bucket size must be < 256");
  if (g_attn_bwd_small && ofa_attn_fwd_small_applicable(a))      // short targets: csrc/attention_small.cu
    return ofa_attn_fwd_small_launch(a, (cudaStream_t)stream);
  CUtensorMap tq, tpq, tk, tpk, tv;
  if (int e = make_qkv_tmap(&tq, a->q, a->T, a->H, a->B, a->ldq, a->bsq, BQ)) return e;
  if (int e = make_qkv_tmap(&tpq, a->pq, a->T, a->H, a->B, a->ldpq, a->bspq, BQ)) return e;
  if (int e = make_qkv_tmap(&tk, a->k, a->S, a->H, a->B, a->ldk, a->bsk, BKV)) return e;
  if (int e = make_qkv_tmap(&tpk, a->pk, a->S, a->H, a->B, a->ldpk, a->bspk, BKV)) return e;
  if (int e = make_qkv_tmap(&tv, a->v, a->S, a->H, a->B, a->ldv, a->bsv, BKV)) return e;
  if (g_attn_fwd_ws && a->S <= WS_MAXK && a->bias.n_img_rel <= kImgLutMax && a->S - a->bias.k_text_off <= 1024) {
    static bool configured_ws = false;
    const int smem_ws = (int)sizeof(WsSmem) + 1024;
    if (!configured_ws) {
      OFA_CUDA(cudaFuncSetAttribute(attn_fwd_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ws));
      OFA_CUDA(cudaFuncSetAttribute(attn_fwd_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ws));
      configured_ws = true;
    }
    dim3 grid_ws((a->T + WS_QT * BQ - 1) / (WS_QT * BQ), a->H, a->B);
    if (g_attn_fwd_ws == 2)
      OFA_CUDA(ofa_launch_pdl(attn_fwd_ws_kernel<false>, grid_ws, WS_THREADS, smem_ws, (cudaStream_t)stream, tq, tpq, tk, tpk, tv, *a));
    else
      OFA_CUDA(ofa_launch_pdl(attn_fwd_ws_kernel<true>, grid_ws, WS_THREADS, smem_ws, (cudaStream_t)stream, tq, tpq, tk, tpk, tv, *a));
    OFA_LAUNCH_CHECK("attn_fwd_ws_kernel");
    return 0;
  }
  static bool configured = false;
  const int smem = (int)sizeof(TcSmem) + 1024;  // 1 KB slack for the 1024B align-up
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((a->T + BQ - 1) / BQ, a->H, a->B);
  OFA_CUDA(ofa_launch_pdl(attn_fwd_tc_kernel, grid, kThreads, smem, (cudaStream_t)stream, tq, tpq, tk, tpk, tv, *a));
  OFA_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return 0;
}

// workspace: dq_acc = B*T*H*128 floats (zero-filled here), delta = B*H*T floats (g->delta)
extern "C" int ofa_attn_bwd_tc(const AttnArgs* a, const AttnGrads* g, float* dq_acc, void* stream) {
  OFA_CHECK(a->T > 0 && a->S > 0 && a->B > 0 && a->H > 0, "ofa_attn_bwd_tc: empty problem");
  OFA_CHECK(a->pq && a->pk && g->delta && dq_acc, "ofa_attn_bwd_tc: pq/pk, delta and dq_acc are required");
  OFA_CHECK(a->bias.ibs < 256 && a->bias.n_img_rel <= kImgHistMax, "ofa_attn_bwd_tc: image bucket table too large");
  OFA_CHECK(a->T <= 1024, "ofa_attn_bwd_tc: T=%d exceeds max positions 1024", a->T);
  OFA_CHECK(a->ldo % 8 == 0 && a->bso % 8 == 0 && g->lddk % 8 == 0 && g->lddpk % 8 == 0 && g->lddv % 8 == 0 &&
                g->lddq % 8 == 0 && g->lddpq % 8 == 0, "ofa_attn_bwd_tc: strides must be multiples of 8 elements");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq, tpq, tk, tpk, tv, tdo, tdq;
  {
    OFA_CHECK(((uintptr_t)dq_acc & 15) == 0, "ofa_attn_bwd_tc: dq_acc must be 16-byte aligned");
    uint64_t dims[4] = {128, (uint64_t)a->H, (uint64_t)a->T, (uint64_t)a->B};
    uint64_t strides[3] = {128 * 4, (uint64_t)a->H * 128 * 4, (uint64_t)a->T * a->H * 128 * 4};
    uint32_t box[4] = {32, 1, (uint32_t)BQ, 1};
    if (int e = ofa_make_tmap(&tdq, dq_acc, 4, dims, strides, box, 1, 4)) return e;
  }
  if (int e = make_qkv_tmap(&tq, a->q, a->T, a->H, a->B, a->ldq, a->bsq, BQ)) return e;
  if (int e = make_qkv_tmap(&tpq, a->pq, a->T, a->H, a->B, a->ldpq, a->bspq, BQ)) return e;
  if (int e = make_qkv_tmap(&tk, a->k, a->S, a->H, a->B, a->ldk, a->bsk, BK2)) return e;
  if (int e = make_qkv_tmap(&tpk, a->pk, a->S, a->H, a->B, a->ldpk, a->bspk, BK2)) return e;
  if (int e = make_qkv_tmap(&tv, a->v, a->S, a->H, a->B, a->ldv, a->bsv, BK2)) return e;
  if (int e = make_qkv_tmap(&tdo, g->dout, a->T, a->H, a->B, a->ldo, a->bso, BQ)) return e;
  static bool configured = false;
  const int smem = (int)sizeof(BwdSmem) + 1024;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const long long nrow = (long long)a->B * a->T * a->H;
  if (g_attn_bwd_small && ofa_attn_bwd_small_applicable(a, g)) {
    // short targets against a long source (decoder cross-attention of the 5 / 12-token tasks): one 128-thread CTA per (batch,
    // head) on warp-level MMA tiles, dQ' in registers -- no accumulator memset, no convert pass (csrc/attention_small.cu)
    OFA_CUDA(ofa_launch_pdl(attn_bwd_delta_kernel, (unsigned)((nrow * 8 + 255) / 256), 256, 0, st, (const __nv_bfloat16*)g->dout, (const __nv_bfloat16*)a->o, a->ldo, a->bso, a->B, a->T, a->H, g->delta));
    OFA_LAUNCH_CHECK("attn_bwd_delta_kernel");
    return ofa_attn_bwd_small_launch(a, g, st);
  }
  OFA_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)nrow * 128 * sizeof(float), st));
  OFA_CUDA(ofa_launch_pdl(attn_bwd_delta_kernel, (unsigned)((nrow * 8 + 255) / 256), 256, 0, st, (const __nv_bfloat16*)g->dout, (const __nv_bfloat16*)a->o, a->ldo, a->bso, a->B, a->T, a->H, g->delta));
  OFA_LAUNCH_CHECK("attn_bwd_delta_kernel");
  dim3 grid((a->S + BK2 - 1) / BK2, a->H, a->B);
  OFA_CUDA(ofa_launch_pdl(attn_bwd_tc_kernel, grid, kBwdThreads, smem, st, tq, tpq, tk, tpk, tv, tdo, tdq, *a, *g));
  OFA_LAUNCH_CHECK("attn_bwd_tc_kernel");
  OFA_CUDA(ofa_launch_pdl(attn_bwd_dq_convert_kernel, (unsigned)((nrow * 16 + 255) / 256), 256, 0, st, dq_acc, (__nv_bfloat16*)g->dq, (__nv_bfloat16*)g->dpq, g->lddq, g->bsdq, g->lddpq, g->bsdpq, a->B, a->T, a->H, g->dq_scale == 0.f ? 1.f : g->dq_scale, g->acc_pos));
  OFA_LAUNCH_CHECK("attn_bwd_dq_convert_kernel");
  return 0;
}
