// bf16 attention backward for SHORT query sequences (T <= 16) without image relative-position bias: the cross-attention (long key
// sequence, no bias) and the causal self-attention (token relative-position LUT and its gradient) of the decoder passes over
// short targets (VG 5, caption 12, gigaword 12 tokens x 835 source positions in the
// bench step; unify_multihead_attention.py:345-398 under autograd).  attn_bwd_tc_kernel spends a whole 128-row query tile, a
// 640-thread CTA with 225 KB of shared memory and a dQ' round trip through an fp32 accumulator on every (batch, head, 128-key
// tile) -- 13 us per CTA, 4032 CTAs, 386 us per launch for 12 real query rows.  Here one 128-thread CTA owns a (batch, head):
// Q' = [q ; pos_q] and dO (16 padded rows) stay in shared memory, every warp streams its own 16-key tiles of K' = [k ; pos_k]
// and V through a private two-stage cp.async ring and works on warp-level tensor-core tiles (mma.sync m16n8k16):
//   S^T = K' Q'^T, dP^T = V dO^T  ->  P^T = exp(S^T - lse), dS^T = P^T o (c dP^T - delta)          (C fragments, keys x queries)
//   dV = P^T dO * c,  dK' = dS^T Q'            (P^T / dS^T re-used as A fragments; written straight to global memory)
//   dQ' += dS K'                               (dS through a 512-byte shared scratch + ldmatrix.trans; fp32 in registers over
//                                               the whole key sweep, summed over the four warps at the end: no accumulator
//                                               round trip, no memset / convert kernels)
// The queries are too few to fill a 128-row tcgen05 tile; the kernel is a stream over K' and V (bytes: S * 192 * 2 per head).
#include <math_constants.h>

#include "attention_common.cuh"

namespace {

constexpr int HD = 64, TQ = 16, TKW = 16, kT = 128;
constexpr int KP = 2 * HD * 2 + 16;       // pitch of a K' row (k | pos_k, bf16) in shared memory
constexpr int VP = HD * 2 + 16;           // pitch of a V / dO row
constexpr int SP = TQ * 2 + 16;           // pitch of the dS^T scratch rows
constexpr int STAGE = TKW * (KP + VP);
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kT) attn_bwd_smallq_kernel(AttnArgs a, AttnGrads g) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned char* ring = dsm;                                   // [4 warps][2 stages][K' 16 x KP | V 16 x VP]
  unsigned char* Qs = dsm + 4 * 2 * STAGE;                     // [16][KP]  q | pos_q
  unsigned char* Os = Qs + TQ * KP;                            // [16][VP]  dO
  unsigned char* Ss = Os + TQ * VP;                            // [4 warps][16][SP]  dS^T scratch
  float* red = reinterpret_cast<float*>(ring);                 // [4][16][128] partial dQ', over the drained rings
  __shared__ float tok_hist[2 * TQ];                           // d tok_lut of this (batch, head): bin (i - j) + TQ - 1
  static_assert(4 * TQ * 128 * 4 <= 4 * 2 * STAGE, "dQ' partials must fit the rings");
  const int b = blockIdx.y, h = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int Tn = a.T, S = a.S;
  pdl_sync();
  // ---- Q' and dO rows of this (batch, head); rows >= T are zero ------------------------------------------------------------
  for (int e = t; e < TQ * 24; e += kT) {
    const int i = e / 24, c = e % 24;                          // 8 chunks of q, 8 of pos_q, 8 of dO
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < Tn) {
      const T* src = c < 8 ? (const T*)a.q + (size_t)b * a.bsq + (size_t)i * a.ldq + h * HD + c * 8
                   : c < 16 ? (const T*)a.pq + (size_t)b * a.bspq + (size_t)i * a.ldpq + h * HD + (c - 8) * 8
                            : (const T*)g.dout + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD + (c - 16) * 8;
      v = *reinterpret_cast<const uint4*>(src);
    }
    unsigned char* dst = c < 16 ? Qs + i * KP + c * 16 : Os + i * VP + (c - 16) * 16;
    *reinterpret_cast<uint4*>(dst) = v;
  }
  const unsigned char* Kb = reinterpret_cast<const unsigned char*>((const T*)a.k + (size_t)b * a.bsk + h * HD);
  const unsigned char* PKb = reinterpret_cast<const unsigned char*>((const T*)a.pk + (size_t)b * a.bspk + h * HD);
  const unsigned char* Vb = reinterpret_cast<const unsigned char*>((const T*)a.v + (size_t)b * a.bsv + h * HD);
  const size_t ldk = (size_t)a.ldk * 2, ldpk = (size_t)a.ldpk * 2, ldv = (size_t)a.ldv * 2;
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)b * S : nullptr;
  const int ntile = (S + TKW - 1) / TKW;
  const int nw = ntile > warp ? (ntile - warp + 3) / 4 : 0;
  const uint32_t wring = smem_u32(ring) + warp * 2 * STAGE;
  auto issue = [&](int n) {
    if (n < nw) {
      const int j0 = (warp + 4 * n) * TKW;
      const uint32_t st = wring + (n & 1) * STAGE;
#pragma unroll
      for (int i = 0; i < 12; ++i) {                           // 16 keys x 24 chunks (8 k, 8 pos_k, 8 v) over 32 lanes
        const int c = lane + 32 * i, key = c / 24, part = c % 24;
        const int j = j0 + key, ok = j < S ? 16 : 0;
        const size_t jj = ok ? j : 0;
        if (part < 8) cp16(st + key * KP + part * 16, Kb + jj * ldk + part * 16, ok);
        else if (part < 16) cp16(st + key * KP + part * 16, PKb + jj * ldpk + (part - 8) * 16, ok);
        else cp16(st + TKW * KP + key * VP + (part - 16) * 16, Vb + jj * ldv + (part - 16) * 16, ok);
      }
    }
    cp_commit();
  };
  issue(0);
  // per-thread row statistics of its four query columns: qa, qa + 1, qa + 8, qa + 9
  const int fr = lane >> 2, fc = (lane & 3) * 2;
  float nl2[4], dl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int qi = fc + (e & 1) + (e >> 1) * 8;
    const size_t ridx = ((size_t)b * a.H + h) * Tn + (qi < Tn ? qi : 0);
    nl2[e] = -a.lse[ridx] * kLog2e;
    dl[e] = g.delta[ridx];
  }
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  const AttnBias& bz = a.bias;
  const float* lut = bz.tok_lut ? bz.tok_lut + (size_t)h * (2 * bz.tok_max - 1) + bz.tok_max - 1 : nullptr;   // indexed by i_t - j_t
  const bool want_dtok = lut && g.dtok_lut;
  if (t < 2 * TQ) tok_hist[t] = 0.f;
  __syncthreads();
  // ldmatrix lane offsets: A-type (rows (l & 7) + 8 * ((l >> 3) & 1), 8-column block l >> 4) and B-type for [row-block][col-half]
  // matrix order (rows (l & 7) + 8 * (l >> 4), 8-column block (l >> 3) & 1)
  const int rowA = (lane & 7) + ((lane >> 3) & 1) * 8, colA = (lane >> 4) * 16;
  const int rowB = (lane & 7) + (lane >> 4) * 8, colB = ((lane >> 3) & 1) * 16;
  const uint32_t qs_u = smem_u32(Qs), os_u = smem_u32(Os), ss_u = smem_u32(Ss) + warp * TQ * SP;
  float dq[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) { dq[i][0] = 0.f; dq[i][1] = 0.f; dq[i][2] = 0.f; dq[i][3] = 0.f; }

  for (int n = 0; n < nw; ++n) {
    __syncwarp();
    issue(n + 1);
    cp_wait<1>();
    __syncwarp();
    const uint32_t kt = wring + (n & 1) * STAGE, vt = kt + TKW * KP;
    const int j0 = (warp + 4 * n) * TKW;
    // ---- S^T = K' Q'^T  and  dP^T = V dO^T   [16 keys x 16 queries] --------------------------------------------------------
    float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t af[4], bf[4];
      ldsm4(af, kt + rowA * KP + colA + ks * 32);
      ldsm4(bf, qs_u + rowB * KP + colB + ks * 32);
      mma16816(st[0], af, bf[0], bf[1]);
      mma16816(st[1], af, bf[2], bf[3]);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4], bf[4];
      ldsm4(af, vt + rowA * VP + colA + ks * 32);
      ldsm4(bf, os_u + rowB * VP + colB + ks * 32);
      mma16816(dp[0], af, bf[0], bf[1]);
      mma16816(dp[1], af, bf[2], bf[3]);
    }
    // ---- P^T, dS^T (fragment element e of block nb: key fr + 8 * (e >> 1), query nb * 8 + fc + (e & 1)) -------------------------
    bool kok[2];
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const int j = j0 + fr + 8 * r2;
      kok[r2] = j < S && !(kpm && kpm[j]);
    }
    uint32_t pa[4], da[4];
    {
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qe = (e & 1) + nb * 2;                     // index into nl2 / dl
          const int qi = nb * 8 + fc + (e & 1), kj = j0 + fr + 8 * (e >> 1);
          bool ok = kok[e >> 1] && qi < Tn && !(a.causal && kj > qi);
          float x = st[nb][e];
          const bool tok_el = lut && ok && qi >= bz.q_text_off && kj >= bz.k_text_off;
          const int rel = (qi - bz.q_text_off) - (kj - bz.k_text_off);
          if (tok_el) x += lut[rel];
          const float pv = ok ? ex2f(fmaf(x, kLog2e, nl2[qe])) : 0.f;
          p[nb][e] = pv;
          ds[nb][e] = pv * fmaf(dp[nb][e], cs, -dl[qe]);
          if (want_dtok && tok_el && rel > -TQ && rel < TQ) atomicAdd(&tok_hist[rel + TQ - 1], ds[nb][e]);
        }
      pa[0] = pk2(p[0][0], p[0][1]); pa[1] = pk2(p[0][2], p[0][3]); pa[2] = pk2(p[1][0], p[1][1]); pa[3] = pk2(p[1][2], p[1][3]);
      da[0] = pk2(ds[0][0], ds[0][1]); da[1] = pk2(ds[0][2], ds[0][3]); da[2] = pk2(ds[1][0], ds[1][1]); da[3] = pk2(ds[1][2], ds[1][3]);
    }
    // dS^T -> scratch [key][query] for the transposed use below
    *reinterpret_cast<uint32_t*>(Ss + warp * TQ * SP + fr * SP + fc * 2) = da[0];
    *reinterpret_cast<uint32_t*>(Ss + warp * TQ * SP + (fr + 8) * SP + fc * 2) = da[1];
    *reinterpret_cast<uint32_t*>(Ss + warp * TQ * SP + fr * SP + (8 + fc) * 2) = da[2];
    *reinterpret_cast<uint32_t*>(Ss + warp * TQ * SP + (fr + 8) * SP + (8 + fc) * 2) = da[3];
    const int jr0 = j0 + fr, jr1 = j0 + fr + 8;
    // ---- dV = P^T dO * c   [16 keys x 64] ---------------------------------------------------------------------------------------
    {
      float dv[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { dv[i][0] = 0.f; dv[i][1] = 0.f; dv[i][2] = 0.f; dv[i][3] = 0.f; }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        ldsm4t(bf, os_u + rowA * VP + colA + np * 32);
        mma16816(dv[2 * np], pa, bf[0], bf[1]);
        mma16816(dv[2 * np + 1], pa, bf[2], bf[3]);
      }
      T* d0 = (T*)g.dv + (size_t)b * g.bsdv + (size_t)jr0 * g.lddv + h * HD + fc;
      T* d1 = (T*)g.dv + (size_t)b * g.bsdv + (size_t)jr1 * g.lddv + h * HD + fc;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (jr0 < S) *reinterpret_cast<uint32_t*>(d0 + nb * 8) = pk2(dv[nb][0] * cs, dv[nb][1] * cs);
        if (jr1 < S) *reinterpret_cast<uint32_t*>(d1 + nb * 8) = pk2(dv[nb][2] * cs, dv[nb][3] * cs);
      }
    }
    // ---- dK' = dS^T Q'   [16 keys x 128]: k part, then pos_k part ---------------------------------------------------------------
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float dk[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { dk[i][0] = 0.f; dk[i][1] = 0.f; dk[i][2] = 0.f; dk[i][3] = 0.f; }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        ldsm4t(bf, qs_u + rowA * KP + colA + (half * 4 + np) * 32);
        mma16816(dk[2 * np], da, bf[0], bf[1]);
        mma16816(dk[2 * np + 1], da, bf[2], bf[3]);
      }
      T* base = half ? (T*)g.dpk + (size_t)b * g.bsdpk + h * HD + fc : (T*)g.dk + (size_t)b * g.bsdk + h * HD + fc;
      const long long ldd = half ? g.lddpk : g.lddk;
#pragma unroll
      const bool addto = half && g.acc_pos;                 // pos_k gradient summed over the layers that share pos_k
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (jr0 < S) {
          uint32_t* d = reinterpret_cast<uint32_t*>(base + (size_t)jr0 * ldd + nb * 8);
          float2 o2 = make_float2(0.f, 0.f);
          if (addto) o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d));
          *d = pk2(dk[nb][0] + o2.x, dk[nb][1] + o2.y);
        }
        if (jr1 < S) {
          uint32_t* d = reinterpret_cast<uint32_t*>(base + (size_t)jr1 * ldd + nb * 8);
          float2 o2 = make_float2(0.f, 0.f);
          if (addto) o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d));
          *d = pk2(dk[nb][2] + o2.x, dk[nb][3] + o2.y);
        }
      }
    }
    // ---- dQ' += dS K'   [16 queries x 128]: A = dS from the scratch (transposed load), B = K' tile (transposed load) ---------------
    __syncwarp();
    {
      uint32_t sa[4];
      // matrices: (keys 0-7, q 0-7) -> a0, (keys 0-7, q 8-15) -> a1, (keys 8-15, q 0-7) -> a2, (keys 8-15, q 8-15) -> a3
      ldsm4t(sa, ss_u + ((lane & 7) + (lane >> 4) * 8) * SP + ((lane >> 3) & 1) * 16);
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        uint32_t bf[4];
        ldsm4t(bf, kt + rowA * KP + colA + np * 32);
        mma16816(dq[2 * np], sa, bf[0], bf[1]);
        mma16816(dq[2 * np + 1], sa, bf[2], bf[3]);
      }
    }
  }
  cp_wait<0>();
  __syncthreads();                                             // every warp is done with its ring
  // ---- sum of the four warps' dQ' partials -> dq (x dq_scale), dpq ---------------------------------------------------------------
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    float* r0 = red + ((size_t)warp * TQ + fr) * 128 + nb * 8 + fc;
    float* r1 = red + ((size_t)warp * TQ + fr + 8) * 128 + nb * 8 + fc;
    r0[0] = dq[nb][0]; r0[1] = dq[nb][1];
    r1[0] = dq[nb][2]; r1[1] = dq[nb][3];
  }
  __syncthreads();
  if (want_dtok && t < 2 * TQ - 1 && tok_hist[t] != 0.f)
    atomicAdd(g.dtok_lut + (size_t)h * (2 * bz.tok_max - 1) + bz.tok_max - 1 + (t - (TQ - 1)), tok_hist[t]);
  const float qsc = g.dq_scale == 0.f ? 1.f : g.dq_scale;
  for (int e = t; e < TQ * 64; e += kT) {                      // two adjacent dims per thread
    const int i = e / 64, d2 = (e % 64) * 2;
    if (i >= Tn) continue;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { v0 += red[((size_t)w * TQ + i) * 128 + d2]; v1 += red[((size_t)w * TQ + i) * 128 + d2 + 1]; }
    if (d2 < HD)
      *reinterpret_cast<uint32_t*>((T*)g.dq + (size_t)b * g.bsdq + (size_t)i * g.lddq + h * HD + d2) = pk2(v0 * qsc, v1 * qsc);
    else {
      uint32_t* d = reinterpret_cast<uint32_t*>((T*)g.dpq + (size_t)b * g.bsdpq + (size_t)i * g.lddpq + h * HD + d2 - HD);
      if (g.acc_pos) { const float2 o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d)); v0 += o2.x; v1 += o2.y; }
      *d = pk2(v0, v1);
    }
  }
}

// ---- forward for the same shapes: S = Q' K'^T per 16-key tile, online softmax per warp (running max / sum of the two query rows
// a thread holds), O += P V with P re-used as the A fragment; the four warps' (max, sum, O) are merged through shared memory.
__global__ void __launch_bounds__(kT) attn_fwd_smallq_kernel(AttnArgs a) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned char* ring = dsm;
  unsigned char* Qs = dsm + 4 * 2 * STAGE;
  float* ow = reinterpret_cast<float*>(ring);                  // [4][16][64] partial outputs, over the drained rings
  __shared__ float mw[4][TQ], lw[4][TQ];
  static_assert(4 * TQ * HD * 4 <= 4 * 2 * STAGE, "partial outputs must fit the rings");
  const int b = blockIdx.y, h = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int Tn = a.T, S = a.S;
  pdl_sync();
  for (int e = t; e < TQ * 16; e += kT) {
    const int i = e / 16, c = e % 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < Tn) {
      const T* src = c < 8 ? (const T*)a.q + (size_t)b * a.bsq + (size_t)i * a.ldq + h * HD + c * 8
                           : (const T*)a.pq + (size_t)b * a.bspq + (size_t)i * a.ldpq + h * HD + (c - 8) * 8;
      v = *reinterpret_cast<const uint4*>(src);
    }
    *reinterpret_cast<uint4*>(Qs + i * KP + c * 16) = v;
  }
  const unsigned char* Kb = reinterpret_cast<const unsigned char*>((const T*)a.k + (size_t)b * a.bsk + h * HD);
  const unsigned char* PKb = reinterpret_cast<const unsigned char*>((const T*)a.pk + (size_t)b * a.bspk + h * HD);
  const unsigned char* Vb = reinterpret_cast<const unsigned char*>((const T*)a.v + (size_t)b * a.bsv + h * HD);
  const size_t ldk = (size_t)a.ldk * 2, ldpk = (size_t)a.ldpk * 2, ldv = (size_t)a.ldv * 2;
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)b * S : nullptr;
  const int ntile = (S + TKW - 1) / TKW;
  const int nw = ntile > warp ? (ntile - warp + 3) / 4 : 0;
  const uint32_t wring = smem_u32(ring) + warp * 2 * STAGE;
  auto issue = [&](int n) {
    if (n < nw) {
      const int j0 = (warp + 4 * n) * TKW;
      const uint32_t st = wring + (n & 1) * STAGE;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int c = lane + 32 * i, key = c / 24, part = c % 24;
        const int j = j0 + key, ok = j < S ? 16 : 0;
        const size_t jj = ok ? j : 0;
        if (part < 8) cp16(st + key * KP + part * 16, Kb + jj * ldk + part * 16, ok);
        else if (part < 16) cp16(st + key * KP + part * 16, PKb + jj * ldpk + (part - 8) * 16, ok);
        else cp16(st + TKW * KP + key * VP + (part - 16) * 16, Vb + jj * ldv + (part - 16) * 16, ok);
      }
    }
    cp_commit();
  };
  issue(0);
  const int fr = lane >> 2, fc = (lane & 3) * 2;
  const AttnBias& bz = a.bias;
  const float* lut = bz.tok_lut ? bz.tok_lut + (size_t)h * (2 * bz.tok_max - 1) + bz.tok_max - 1 : nullptr;
  __syncthreads();
  const int rowA = (lane & 7) + ((lane >> 3) & 1) * 8, colA = (lane >> 4) * 16;
  const int rowB = (lane & 7) + (lane >> 4) * 8, colB = ((lane >> 3) & 1) * 16;
  const uint32_t qs_u = smem_u32(Qs);
  uint32_t qa[8][4];                                           // A fragments of Q' (16 queries x 128 dims), loaded once
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) ldsm4(qa[ks], qs_u + rowA * KP + colA + ks * 32);
  float m[2] = {-CUDART_INF_F, -CUDART_INF_F}, l[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
  for (int n = 0; n < nw; ++n) {
    __syncwarp();
    issue(n + 1);
    cp_wait<1>();
    __syncwarp();
    const uint32_t kt = wring + (n & 1) * STAGE, vt = kt + TKW * KP;
    const int j0 = (warp + 4 * n) * TKW;
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};          // block nb: keys nb * 8 + fc + (e & 1); rows fr + 8 * (e >> 1)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t bf[4];
      ldsm4(bf, kt + rowB * KP + colB + ks * 32);
      mma16816(s[0], qa[ks], bf[0], bf[1]);
      mma16816(s[1], qa[ks], bf[2], bf[3]);
    }
    float tmx[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = fr + 8 * (e >> 1), kj = j0 + nb * 8 + fc + (e & 1);
        const bool ok = kj < S && !(kpm && kpm[kj]) && !(a.causal && kj > qi);
        float x = s[nb][e];
        if (lut && qi >= bz.q_text_off && kj >= bz.k_text_off && ok) x += lut[(qi - bz.q_text_off) - (kj - bz.k_text_off)];
        x = ok ? x * kLog2e : -CUDART_INF_F;
        s[nb][e] = x;
        tmx[e >> 1] = fmaxf(tmx[e >> 1], x);
      }
    float alpha[2], mu[2];
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      tmx[r2] = fmaxf(tmx[r2], __shfl_xor_sync(0xffffffffu, tmx[r2], 1));
      tmx[r2] = fmaxf(tmx[r2], __shfl_xor_sync(0xffffffffu, tmx[r2], 2));
      const float mn = fmaxf(m[r2], tmx[r2]);
      mu[r2] = mn == -CUDART_INF_F ? 0.f : mn;
      alpha[r2] = ex2f(m[r2] - mu[r2]);
      m[r2] = mn;
    }
    float ps[2] = {0.f, 0.f};
    uint32_t pa[4];
    {
      float p[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          p[nb][e] = ex2f(s[nb][e] - mu[e >> 1]);
          ps[e >> 1] += p[nb][e];
        }
      pa[0] = pk2(p[0][0], p[0][1]); pa[1] = pk2(p[0][2], p[0][3]); pa[2] = pk2(p[1][0], p[1][1]); pa[3] = pk2(p[1][2], p[1][3]);
    }
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      ps[r2] += __shfl_xor_sync(0xffffffffu, ps[r2], 1);
      ps[r2] += __shfl_xor_sync(0xffffffffu, ps[r2], 2);
      l[r2] = l[r2] * alpha[r2] + ps[r2];
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) { o[nb][0] *= alpha[0]; o[nb][1] *= alpha[0]; o[nb][2] *= alpha[1]; o[nb][3] *= alpha[1]; }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bf[4];
      ldsm4t(bf, vt + rowA * VP + colA + np * 32);
      mma16816(o[2 * np], pa, bf[0], bf[1]);
      mma16816(o[2 * np + 1], pa, bf[2], bf[3]);
    }
  }
  cp_wait<0>();
  __syncthreads();
  if ((lane & 3) == 0) { mw[warp][fr] = m[0]; mw[warp][fr + 8] = m[1]; lw[warp][fr] = l[0]; lw[warp][fr + 8] = l[1]; }
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    float* r0 = ow + ((size_t)warp * TQ + fr) * HD + nb * 8 + fc;
    float* r1 = ow + ((size_t)warp * TQ + fr + 8) * HD + nb * 8 + fc;
    r0[0] = o[nb][0]; r0[1] = o[nb][1];
    r1[0] = o[nb][2]; r1[1] = o[nb][3];
  }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < TQ * (HD / 2); e += kT) {
    const int i = e / (HD / 2), d2 = (e % (HD / 2)) * 2;
    if (i >= Tn) continue;
    const float M = fmaxf(fmaxf(mw[0][i], mw[1][i]), fmaxf(mw[2][i], mw[3][i]));
    const float Mu = M == -CUDART_INF_F ? 0.f : M;
    float L = 0.f, v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = ex2f(mw[w][i] - Mu);
      L = fmaf(lw[w][i], f, L);
      v0 = fmaf(ow[((size_t)w * TQ + i) * HD + d2], f, v0);
      v1 = fmaf(ow[((size_t)w * TQ + i) * HD + d2 + 1], f, v1);
    }
    const float inv = L > 0.f ? cs / L : 0.f;
    *reinterpret_cast<uint32_t*>((T*)a.o + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD + d2) = pk2(v0 * inv, v1 * inv);
    if (d2 == 0) a.lse[((size_t)b * a.H + h) * Tn + i] = (Mu + log2f(L)) * 0.6931471805599453f;
  }
}

}  // namespace

bool ofa_attn_bwd_small_applicable(const AttnArgs* a, const AttnGrads* g) {
  // a token relative-position LUT needs |i_t - j_t| < 16 for every pair: self-attention (S = T) with equal text offsets
  const bool lut_ok = a->bias.tok_lut == nullptr || (a->S <= TQ && a->bias.q_text_off == a->bias.k_text_off);
  return a->T <= TQ && lut_ok && a->bias.img_lut == nullptr && a->q_pos_off == 0 &&
         a->ldq % 8 == 0 && a->ldpq % 8 == 0 && a->ldk % 8 == 0 && a->ldpk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 &&
         a->bsq % 8 == 0 && a->bspq % 8 == 0 && a->bsk % 8 == 0 && a->bspk % 8 == 0 && a->bsv % 8 == 0 && a->bso % 8 == 0 &&
         g->lddq % 2 == 0 && g->lddpq % 2 == 0 && g->lddk % 2 == 0 && g->lddpk % 2 == 0 && g->lddv % 2 == 0;
}

int ofa_attn_bwd_small_launch(const AttnArgs* a, const AttnGrads* g, cudaStream_t st) {
  const int smem = 4 * 2 * STAGE + TQ * KP + TQ * VP + 4 * TQ * SP;
  static bool configured = false;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(attn_bwd_smallq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  OFA_CUDA(ofa_launch_pdl(attn_bwd_smallq_kernel, dim3(a->H, a->B), kT, smem, st, *a, *g));
  OFA_LAUNCH_CHECK("attn_bwd_smallq_kernel");
  return 0;
}

bool ofa_attn_fwd_small_applicable(const AttnArgs* a) {
  const bool lut_ok = a->bias.tok_lut == nullptr || (a->S <= TQ && a->bias.q_text_off == a->bias.k_text_off);
  // short key sequences only (decoder self-attention): over S = 835 source positions the warp-specialised tcgen05 forward
  // (72.8 us at B = 48, T = 12) beats this one-tile-in-flight stream (79.7 us); over S = T = 12 it is 27.8 -> 14.7 us
  return a->T <= TQ && a->S <= 4 * TKW && lut_ok && a->bias.img_lut == nullptr && a->q_pos_off == 0 &&
         a->ldq % 8 == 0 && a->ldpq % 8 == 0 && a->ldk % 8 == 0 && a->ldpk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 2 == 0 &&
         a->bsq % 8 == 0 && a->bspq % 8 == 0 && a->bsk % 8 == 0 && a->bspk % 8 == 0 && a->bsv % 8 == 0 && a->bso % 2 == 0;
}

int ofa_attn_fwd_small_launch(const AttnArgs* a, cudaStream_t st) {
  const int smem = 4 * 2 * STAGE + TQ * KP;
  static bool configured = false;
  if (!configured) {
    OFA_CUDA(cudaFuncSetAttribute(attn_fwd_smallq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  OFA_CUDA(ofa_launch_pdl(attn_fwd_smallq_kernel, dim3(a->H, a->B), kT, smem, st, *a));
  OFA_LAUNCH_CHECK("attn_fwd_smallq_kernel");
  return 0;
}
