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

#include "attention_common.cuh"

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int kThreads = 128;
constexpr int kTmemCols = 128;  // S: [0,64)  O_partial: [64,128)

struct TcSmem {
  uint8_t q[2][BQ * 128];   // [0: q, 1: pos_q][128 rows][128 B]   K-major, SWIZZLE_128B
  uint8_t k[2][BKV * 128];  // [0: k, 1: pos_k][64 keys][128 B]    K-major
  uint8_t v[BKV * 128];     // [64 keys][64 x bf16]                MN-major B operand of P.V
  uint8_t p[BQ * 128];      // [128 rows][64 keys x bf16]          K-major A operand of P.V
  uint64_t bar_q, bar_k, bar_v, bar_s, bar_o;
  uint32_t tmem_addr;
  int kinfo[BKV];  // per key of the current tile: bit31 masked | bit30 image key | [8,16) col | [0,8) row
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
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (a.S + BKV - 1) / BKV;
  const AttnBias& bz = a.bias;

  if (t == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmPQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmPK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_k, 1); mbar_init(&sm.bar_v, 1); mbar_init(&sm.bar_s, 1);
    mbar_init(&sm.bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(&sm.tmem_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = sm.tmem_addr + ((uint32_t)(warp * 32) << 16);
  const uint32_t tmem_o = tmem_s + BKV;

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
  const int i = q0 + t;
  const int iabs = i + a.q_pos_off;
  const bool q_text = bz.tok_lut && iabs >= bz.q_text_off;
  const bool q_img = bz.img_lut && iabs < bz.n_img_q && i < a.T;
  int qr = 0, qc = 0;
  if (q_img) {
    const int pid = bz.q_pid[(size_t)b * bz.n_img_q + iabs] - 1;
    qr = pid / bz.ibs; qc = pid % bz.ibs;
  }
  const float* tok_lut = bz.tok_lut ? bz.tok_lut + (size_t)h * (2 * bz.tok_max - 1) + (iabs - bz.q_text_off) + bz.tok_max - 1 +
                                          bz.k_text_off : nullptr;  // index with -j
  const float* img_lut = bz.img_lut ? bz.img_lut + (size_t)h * bz.n_img_rel : nullptr;
  const int w83 = 2 * bz.ibs - 1;
  constexpr float kLog2e = 1.4426950408889634f;
  float m = -CUDART_INF_F, l = 0.f;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;

  constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);

  for (int jt = 0; jt < ntiles; ++jt) {
    const int k0 = jt * BKV;
    const uint32_t ph = jt & 1;
    // key-side metadata for this tile
    if (t < BKV) {
      const int j = k0 + t;
      int info = 0;
      if (j >= a.S || (a.kpm && a.kpm[(size_t)b * a.S + j])) info |= (int)0x80000000u;
      if (bz.img_lut && j < bz.n_img_k) {
        const int pid = bz.k_pid[(size_t)b * bz.n_img_k + j] - 1;
        info |= 0x40000000 | ((pid % bz.ibs) << 8) | (pid / bz.ibs);
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
    __syncthreads();  // kinfo visible; also reconverges warp 0
    mbar_wait(&sm.bar_s, ph);
    tc_fence_after();
    if (t == 0 && jt + 1 < ntiles) {  // K' buffer is free: prefetch the next key tile under the softmax
      mbar_expect_tx(&sm.bar_k, 2 * BKV * 128);
      tma_load_4d(sm.k[0], &tmK, &sm.bar_k, 0, h, k0 + BKV, b);
      tma_load_4d(sm.k[1], &tmPK, &sm.bar_k, 0, h, k0 + BKV, b);
    }
    float s[BKV];
#pragma unroll
    for (int c = 0; c < BKV / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_s + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int jl = c * 32 + jj, j = k0 + jl;
        const int info = sm.kinfo[jl];
        float x = __uint_as_float(r[jj]);
        if (q_text && j >= bz.k_text_off) x += __ldg(tok_lut - j);
        if (q_img && (info & 0x40000000)) {
          const int kr = info & 0xff, kc = (info >> 8) & 0xff;
          x += __ldg(img_lut + (qr - kr + bz.ibs - 1) * w83 + (qc - kc + bz.ibs - 1));
        }
        if (info < 0 || (a.causal && j > iabs)) x = -CUDART_INF_F;
        s[jl] = x;
      }
    }
    float mx = m;
#pragma unroll
    for (int jl = 0; jl < BKV; ++jl) mx = fmaxf(mx, s[jl]);
    const float mu = (mx == -CUDART_INF_F) ? 0.f : mx;
    const float alpha = ex2((m - mu) * kLog2e);
    m = mx;
    float rs = 0.f;
    const float mneg = -mu * kLog2e;
#pragma unroll
    for (int c16 = 0; c16 < BKV / 8; ++c16) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        pv[e] = ex2(fmaf(s[c16 * 8 + e], kLog2e, mneg));
        rs += pv[e];
      }
      const uint4 pk = make_uint4(pack_bf16(pv[0], pv[1]), pack_bf16(pv[2], pv[3]), pack_bf16(pv[4], pv[5]),
                                  pack_bf16(pv[6], pv[7]));
      *reinterpret_cast<uint4*>(sm.p + t * 128 + ((c16 ^ (t & 7)) << 4)) = pk;
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
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_o + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) o[c * 32 + jj] = fmaf(o[c * 32 + jj], alpha, __uint_as_float(r[jj]));
    }
    tc_fence_before();
  }

  if (i < a.T) {
    const float inv = (l > 0.f ? 1.f / l : 0.f) * (a.head_scale ? a.head_scale[h] : 1.f);
    __nv_bfloat16* O = (__nv_bfloat16*)a.o + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      reinterpret_cast<uint4*>(O)[c] =
          make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                     pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    a.lse[((size_t)b * a.H + h) * a.T + i] = (m == -CUDART_INF_F ? 0.f : m) + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(sm.tmem_addr);
  }
}

int make_qkv_tmap(CUtensorMap* tm, const void* p, int L, int H, int B, long long ld, long long bs, int box_rows) {
  // dims {64 (head dim), H, L, B}; box {64, 1, box_rows, 1}
  uint64_t dims[4] = {(uint64_t)HD, (uint64_t)H, (uint64_t)L, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)ld * 2, (uint64_t)bs * 2};
  uint32_t box[4] = {HD, 1, (uint32_t)box_rows, 1};
  return ofa_make_tmap(tm, p, 4, dims, strides, box, 1, 2);
}

}  // namespace

extern "C" int ofa_attn_fwd_tc(const AttnArgs* a, void* stream) {
  OFA_CHECK(a->T > 0 && a->S > 0 && a->B > 0 && a->H > 0, "ofa_attn_fwd_tc: empty problem");
  OFA_CHECK(a->pq && a->pk, "ofa_attn_fwd_tc: the absolute-position operands pq/pk are required");
  OFA_CHECK(a->ldq % 8 == 0 && a->ldpq % 8 == 0 && a->ldk % 8 == 0 && a->ldpk % 8 == 0 && a->ldv % 8 == 0 &&
                a->ldo % 8 == 0 && a->bso % 8 == 0,
            "ofa_attn_fwd_tc: strides must be multiples of 8 elements");
  OFA_CHECK(a->bias.ibs < 256, "ofa_attn_fwd_tc: image bucket size must be < 256");
  CUtensorMap tq, tpq, tk, tpk, tv;
  if (int e = make_qkv_tmap(&tq, a->q, a->T, a->H, a->B, a->ldq, a->bsq, BQ)) return e;
  if (int e = make_qkv_tmap(&tpq, a->pq, a->T, a->H, a->B, a->ldpq, a->bspq, BQ)) return e;
  if (int e = make_qkv_tmap(&tk, a->k, a->S, a->H, a->B, a->ldk, a->bsk, BKV)) return e;
  if (int e = make_qkv_tmap(&tpk, a->pk, a->S, a->H, a->B, a->ldpk, a->bspk, BKV)) return e;
  if (int e = make_qkv_tmap(&tv, a->v, a->S, a->H, a->B, a->ldv, a->bsv, BKV)) return e;
  static bool configured = false;
  const int smem = (int)sizeof(TcSmem) + 1024;  // 1 KB slack for the 1024B align-up
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((a->T + BQ - 1) / BQ, a->H, a->B);
  attn_fwd_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tq, tpq, tk, tpk, tv, *a);
  OFA_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return 0;
}
