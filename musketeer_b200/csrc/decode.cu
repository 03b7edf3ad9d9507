// Incremental-decoding kernels for beam search (models/sequence_generator.py:209-598 driving the decoder layers of
// models/ofa/unify_transformer_layer.py:432-582 with `incremental_state`): single-token attention over a KV cache with
// row indirection, and the beam reorder of the self-attention cache.
//
//   ofa_attn_decode : one query token per row.  G consecutive rows (the beams of one sentence) form a group that reads
//     the SAME cache row, so cross-attention K / V / pos_k are stored once per sentence and read once per group instead of
//     being replicated per beam and re-gathered on every reorder (the reference `index_select`s [bsz*beam, H, N, 64]
//     K and V of every layer at every step: models/ofa/unify_multihead_attention.py:458-480).  Self-attention uses G = 1
//     with one cache row per beam; the shared absolute-position keys are one row for everybody (pk_row).
//     scores = q.k + pos_q.pos_k (+ rel-pos LUT for self-attention) with key padding; fp32 softmax; out = P V * c_attn.
//   ofa_cache_gather : dst[r, :L] = src[order[r], :L] for every (layer, k|v) plane in one launch -- the beam reorder touches
//     only the L valid positions of the cache, not its capacity.
// HBM / L2-bound SIMT kernels (one token of queries: nothing for the tensor cores to do); bf16 or fp32 storage.
#include <math_constants.h>

#include "common.cuh"

struct OfaDecodeArgs {   // mirrored by musketeer_b200/_lib.py
  const void* q;  const void* pq;  long long ldq, ldpq;          // [R, *] one token per row (q, pos_q pre-scaled); pq may be null
  const void* k;  const void* v;   const void* pk;               // caches: (row, j, h*64 + d) at row*bs + j*ld + h*64 + d
  long long ldk, bsk, ldv, bsv, ldpk, bspk;                      //   pk null: no absolute-position term in this launch
  const int* kv_row;                                             // [R / G] cache row of each group (null: group index)
  const int* pk_row;                                             // [R / G] cache row of pos_k (null: kv_row)
  const unsigned char* kpm;  long long kpm_stride;               // [cache rows, >= S] 1 = padded key (null: none)
  void* o;  long long ldo;                                       // [R, H*64]
  const float* head_scale;                                       // [H] c_attn or null
  const float* tok_lut;  int tok_max;  int q_pos;                // self-attention rel-pos: lut[h][(q_pos - j) + tok_max - 1]
  int R, G, H, S;
  // ---- round 2 ----
  const float* bias_in;  float* score_out;  long long bias_ld;   // [R][H][bias_ld] fp32: bias_in is added to the scores (a term
                                                                 // that is the same for every layer, computed once per step by a
                                                                 // launch with score_out set: that launch writes its raw scores and
                                                                 // does no softmax / P.V)
  const int* page_table;  int page_len, max_pages;               // paged K / V: key j of cache row r lives in page
  long long page_stride;                                         // page_table[r*max_pages + j/page_len] at offset (j % page_len)*ld;
};                                                               // page p starts at p*page_stride elements (null: contiguous rows)

namespace {

constexpr int HD = 64;
constexpr int kT = 128;
constexpr int GMAX = 8;
int g_decode_online = 1;    // A/B switch (ofa_attn_decode_set_online): 0 = the two-pass warp-MMA kernel for the long key ranges
int g_decode_short = 1;     // A/B switch (ofa_attn_decode_set_short): 0 = the (group, head) kernel for every shape

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T>
__device__ __forceinline__ float2 load2(const T* p);
template <>
__device__ __forceinline__ float2 load2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// One CTA = one (group, head); 4 warps stream the keys in tiles of 4 per warp: lane = (key of the tile, 8-element chunk of its
// 64-dim row), so one load instruction of a warp covers four whole 128-byte (bf16) rows -- the round-1 kernel gave every thread
// its own row (32 different lines per instruction) and sat at 2 TB/s on the L1 tag rate.  Scores are reduced over the 8 lanes of
// a key by shuffles, the softmax is online (running max / sum per query of the group, one pass over K and V, no score buffer
// in shared memory), every lane accumulates two output dims for all G queries; the four warps' partial (max, sum, out) are
// merged through shared memory at the end.
template <typename T, int G>
__global__ void __launch_bounds__(kT) attn_decode_kernel(OfaDecodeArgs a) {
  pdl_sync();
  __shared__ __align__(16) float qs[G][2][HD];          // q | pos_q of the group's rows, this head
  __shared__ float wm[kT / 32][G], wl[kT / 32][G];
  __shared__ float wo[kT / 32][G][HD];
  const int grp = blockIdx.x, h = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int S = a.S;
  const int row0 = grp * G;
  const int krow = a.kv_row ? a.kv_row[grp] : grp;
  const int prow = a.pk_row ? a.pk_row[grp] : krow;
  const bool has_pk = a.pk != nullptr && a.pq != nullptr;
  for (int e = t; e < G * 2 * HD; e += kT) {
    const int g = e / (2 * HD), d = e % (2 * HD);
    float v = 0.f;
    if (row0 + g < a.R) {
      if (d < HD) v = (float)reinterpret_cast<const T*>(a.q)[(size_t)(row0 + g) * a.ldq + h * HD + d];
      else if (has_pk) v = (float)reinterpret_cast<const T*>(a.pq)[(size_t)(row0 + g) * a.ldpq + h * HD + d - HD];
    }
    qs[g][d / HD][d % HD] = v;
  }
  __syncthreads();
  const T* Kb = reinterpret_cast<const T*>(a.k) + h * HD;
  const T* Vb = a.v ? reinterpret_cast<const T*>(a.v) + h * HD : nullptr;
  const T* PK = has_pk ? reinterpret_cast<const T*>(a.pk) + (size_t)prow * a.bspk + h * HD : nullptr;
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)krow * a.kpm_stride : nullptr;
  const float* lut = a.tok_lut ? a.tok_lut + (size_t)h * (2 * a.tok_max - 1) : nullptr;
  const int* ptab = a.page_table ? a.page_table + (size_t)krow * a.max_pages : nullptr;
  const int ks = lane >> 3, c = lane & 7;          // key of the 4-key tile, 8-element chunk of the row

  float m[G], l[G], o0[G], o1[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { m[g] = -CUDART_INF_F; l[g] = 0.f; o0[g] = 0.f; o1[g] = 0.f; }

  for (int j0 = warp * 4; j0 < S; j0 += (kT / 32) * 4) {
    const int j = j0 + ks;
    const bool valid = j < S;
    const bool masked = !valid || (kpm && kpm[j]);
    // address of key / value row j (contiguous cache rows, or through the page table)
    size_t off = 0;
    if (valid) off = ptab ? (size_t)ptab[j / a.page_len] * a.page_stride + (size_t)(j % a.page_len) * a.ldk
                          : (size_t)krow * a.bsk + (size_t)j * a.ldk;
    float kv[8], pv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { kv[e] = 0.f; pv[e] = 0.f; }
    if (!masked) {
      load8<T>(Kb + off + c * 8, kv);
      if (has_pk) load8<T>(PK + (size_t)j * a.ldpk + c * 8, pv);
    }
    // V rows of the tile's four keys (two output dims per lane), requested before the score arithmetic
    float2 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vv[u] = make_float2(0.f, 0.f);
      const int ju = j0 + u;
      if (Vb && ju < S) {
        const size_t offv = ptab ? (size_t)ptab[ju / a.page_len] * a.page_stride + (size_t)(ju % a.page_len) * a.ldv
                                 : (size_t)krow * a.bsv + (size_t)ju * a.ldv;
        vv[u] = load2<T>(Vb + offv + lane * 2);
      }
    }
    float extra = 0.f;
    if (lut && valid) {
      const int rel = a.q_pos - j + a.tok_max - 1;
      if (rel >= 0 && rel < 2 * a.tok_max - 1) extra = lut[rel];
    }
    float sc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float4 q0 = *reinterpret_cast<const float4*>(&qs[g][0][c * 8]), q1 = *reinterpret_cast<const float4*>(&qs[g][0][c * 8 + 4]);
      float s = q0.x * kv[0];
      s = fmaf(q0.y, kv[1], s); s = fmaf(q0.z, kv[2], s); s = fmaf(q0.w, kv[3], s);
      s = fmaf(q1.x, kv[4], s); s = fmaf(q1.y, kv[5], s); s = fmaf(q1.z, kv[6], s); s = fmaf(q1.w, kv[7], s);
      if (has_pk) {
        const float4 p0 = *reinterpret_cast<const float4*>(&qs[g][1][c * 8]), p1 = *reinterpret_cast<const float4*>(&qs[g][1][c * 8 + 4]);
        s = fmaf(p0.x, pv[0], s); s = fmaf(p0.y, pv[1], s); s = fmaf(p0.z, pv[2], s); s = fmaf(p0.w, pv[3], s);
        s = fmaf(p1.x, pv[4], s); s = fmaf(p1.y, pv[5], s); s = fmaf(p1.z, pv[6], s); s = fmaf(p1.w, pv[7], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += extra;
      if (a.bias_in && valid && row0 + g < a.R) s += a.bias_in[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j];
      sc[g] = masked ? -CUDART_INF_F : s;
    }
    if (a.score_out) {            // bias pre-pass: raw scores out, no softmax
      if (c == 0 && valid) {
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (row0 + g < a.R) a.score_out[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j] = masked ? 0.f : sc[g];
      }
      continue;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] = __shfl_sync(0xffffffffu, sc[g], u * 8);
      const float mx = fmaxf(fmaxf(s4[0], s4[1]), fmaxf(s4[2], s4[3]));
      const float mn = fmaxf(m[g], mx);
      const float mu = mn == -CUDART_INF_F ? 0.f : mn;
      const float alpha = __expf(m[g] - mu);
      m[g] = mn;
      float ps = 0.f, a0 = o0[g] * alpha, a1 = o1[g] * alpha;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float p = __expf(s4[u] - mu);
        ps += p;
        a0 = fmaf(p, vv[u].x, a0);
        a1 = fmaf(p, vv[u].y, a1);
      }
      l[g] = l[g] * alpha + ps;
      o0[g] = a0; o1[g] = a1;
    }
  }
  if (a.score_out) return;
  // merge the four warps
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (lane == 0) { wm[warp][g] = m[g]; wl[warp][g] = l[g]; }
    wo[warp][g][lane * 2] = o0[g];
    wo[warp][g][lane * 2 + 1] = o1[g];
  }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    if (row0 + g >= a.R) continue;
    float mm = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) mm = fmaxf(mm, wm[w][g]);
    const float mu = mm == -CUDART_INF_F ? 0.f : mm;
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) {
      const float f = __expf(wm[w][g] - mu);
      ll = fmaf(wl[w][g], f, ll);
      oo = fmaf(wo[w][g][d], f, oo);
    }
    reinterpret_cast<T*>(a.o)[(size_t)(row0 + g) * a.ldo + h * HD + d] = (T)((ll > 0.f ? oo / ll : 0.f) * cs);
  }
}

// ---- short-key decode attention (incremental self-attention: S <= 32 keys, one query per row) ------------------------------
// One WARP = one (row, head), four per CTA, no block-level synchronisation: lane j owns key j (its whole 64-dim K row and
// position-key row: 128-byte lines fetched in two batches of four 16-byte chunks), the softmax is three warp reductions, and
// for P.V every lane owns two output dims and walks the <= 32 value rows (one coalesced 128-byte line per key, all requested
// before the arithmetic).  The (group, head) kernel above spends ~20 us on 3840 such problems (CTA-wide merges through
// shared memory, two dependent tiles per warp): here the whole launch is one wave of independent warps.
template <typename T>
__global__ void __launch_bounds__(kT, sizeof(T) == 2 ? 7 : 4) attn_decode_short_kernel(OfaDecodeArgs a) {     // (960 CTAs at 320 rows x 12 heads: one wave)
  pdl_sync();
  __shared__ __align__(16) float qs[kT / 32][2][HD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (kT / 32) + warp;
  if (item >= a.R * a.H) return;
  const int r = item / a.H, h = item % a.H;
  const int S = a.S;
  const int krow = a.kv_row ? a.kv_row[r] : r;
  const int prow = a.pk_row ? a.pk_row[r] : krow;
  const bool has_pk = a.pk != nullptr && a.pq != nullptr;
  {
    const float2 q2 = load2<T>(reinterpret_cast<const T*>(a.q) + (size_t)r * a.ldq + h * HD + lane * 2);
    qs[warp][0][lane * 2] = q2.x; qs[warp][0][lane * 2 + 1] = q2.y;
    if (has_pk) {
      const float2 p2 = load2<T>(reinterpret_cast<const T*>(a.pq) + (size_t)r * a.ldpq + h * HD + lane * 2);
      qs[warp][1][lane * 2] = p2.x; qs[warp][1][lane * 2 + 1] = p2.y;
    }
  }
  __syncwarp();
  const int* ptab = a.page_table ? a.page_table + (size_t)krow * a.max_pages : nullptr;
  const T* Vb = reinterpret_cast<const T*>(a.v) + h * HD + lane * 2;
  auto v_off = [&](int ju) {
    return ptab ? (size_t)ptab[ju / a.page_len] * a.page_stride + (size_t)(ju % a.page_len) * a.ldv
                : (size_t)krow * a.bsv + (size_t)ju * a.ldv;
  };
  // value rows of the first VH keys: requested before anything else (they depend on neither q nor the scores), so the launch
  // pays one DRAM round trip for K and V together
  constexpr int VH = sizeof(T) == 2 ? 16 : 0;
  float2 vh[VH > 0 ? VH : 1];
#pragma unroll
  for (int u = 0; u < VH; ++u) {
    vh[u] = make_float2(0.f, 0.f);
    if (u < S) vh[u] = load2<T>(Vb + v_off(u));
  }
  const int j = lane;
  const bool valid = j < S;
  const bool masked = !valid || (a.kpm && a.kpm[(size_t)krow * a.kpm_stride + j]);
  float s = 0.f;
  if (!masked) {
    const size_t off = ptab ? (size_t)ptab[j / a.page_len] * a.page_stride + (size_t)(j % a.page_len) * a.ldk
                            : (size_t)krow * a.bsk + (size_t)j * a.ldk;
    const T* kp = reinterpret_cast<const T*>(a.k) + h * HD + off;
    const T* pp = has_pk ? reinterpret_cast<const T*>(a.pk) + (size_t)prow * a.bspk + h * HD + (size_t)j * a.ldpk : nullptr;
#pragma unroll
    for (int half = 0; half < 4; ++half) {      // (chunks of one 128-byte line: only the first batch waits for DRAM)
      float kv[2][8], pv[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        load8<T>(kp + (half * 2 + c) * 8, kv[c]);
        if (has_pk) load8<T>(pp + (half * 2 + c) * 8, pv[c]);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 q0 = *reinterpret_cast<const float4*>(&qs[warp][0][(half * 2 + c) * 8]);
        const float4 q1 = *reinterpret_cast<const float4*>(&qs[warp][0][(half * 2 + c) * 8 + 4]);
        s = fmaf(q0.x, kv[c][0], s); s = fmaf(q0.y, kv[c][1], s); s = fmaf(q0.z, kv[c][2], s); s = fmaf(q0.w, kv[c][3], s);
        s = fmaf(q1.x, kv[c][4], s); s = fmaf(q1.y, kv[c][5], s); s = fmaf(q1.z, kv[c][6], s); s = fmaf(q1.w, kv[c][7], s);
        if (has_pk) {
          const float4 p0 = *reinterpret_cast<const float4*>(&qs[warp][1][(half * 2 + c) * 8]);
          const float4 p1 = *reinterpret_cast<const float4*>(&qs[warp][1][(half * 2 + c) * 8 + 4]);
          s = fmaf(p0.x, pv[c][0], s); s = fmaf(p0.y, pv[c][1], s); s = fmaf(p0.z, pv[c][2], s); s = fmaf(p0.w, pv[c][3], s);
          s = fmaf(p1.x, pv[c][4], s); s = fmaf(p1.y, pv[c][5], s); s = fmaf(p1.z, pv[c][6], s); s = fmaf(p1.w, pv[c][7], s);
        }
      }
    }
    if (a.tok_lut) {
      const int rel = a.q_pos - j + a.tok_max - 1;
      if (rel >= 0 && rel < 2 * a.tok_max - 1) s += a.tok_lut[(size_t)h * (2 * a.tok_max - 1) + rel];
    }
    if (a.bias_in) s += a.bias_in[((size_t)r * a.H + h) * a.bias_ld + j];
  } else {
    s = -CUDART_INF_F;
  }
  float mx = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float mu = mx == -CUDART_INF_F ? 0.f : mx;
  const float p = __expf(s - mu);               // (masked: exp(-inf) = 0)
  float l = p;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int u = 0; u < VH; ++u) {
    const float pu = __shfl_sync(0xffffffffu, p, u);
    if (u < S) { o0 = fmaf(pu, vh[u].x, o0); o1 = fmaf(pu, vh[u].y, o1); }
  }
  for (int j0 = VH; j0 < S; j0 += 8) {
    float2 vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int ju = j0 + u;
      vv[u] = make_float2(0.f, 0.f);
      if (ju < S) vv[u] = load2<T>(Vb + v_off(ju));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float pu = __shfl_sync(0xffffffffu, p, (j0 + u) & 31);
      if (j0 + u < S) { o0 = fmaf(pu, vv[u].x, o0); o1 = fmaf(pu, vv[u].y, o1); }
    }
  }
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  const float inv = l > 0.f ? cs / l : 0.f;
  T* op = reinterpret_cast<T*>(a.o) + (size_t)r * a.ldo + h * HD + lane * 2;
  if constexpr (sizeof(T) == 2) {
    *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
  } else {
    *reinterpret_cast<float2*>(op) = make_float2(o0 * inv, o1 * inv);
  }
}

// ---- long-key decode attention (cross-attention over the encoder output: S ~ 900, G beams per sentence) ----------------
// One CTA = one (group, head), 128 threads.  K and V rows of this head (64 elements every ld elements) are streamed through a
// ring of cp.async stages in shared memory (TK keys per stage, rows padded by 16 bytes so that 16-byte shared loads of
// neighbouring keys fall into different banks): the copies of NST-1 stages are in flight while a stage is consumed, so the
// DRAM latency is paid once per launch instead of once per tile (the register-only kernel above keeps one tile per warp in
// flight and its online softmax serialises the tiles).  Two streamed passes over a score buffer in shared memory:
//   pass 1  K tiles -> scores[g][j] = q_g . k_j (+ bias_in, key padding)          thread = (key, quarter of the 64 dims)
//   softmax per query over the score buffer (one warp per query, fp32)            V tiles are already being fetched
//   pass 2  V tiles -> out_g += p[g][j] * v_j                                      thread = (two output dims, quarter of the keys)
// Algorithmic bytes per launch: groups * S * 64 * H * 2 tensors * sizeof(T).
constexpr int TK = 32;       // keys per stage
constexpr int NST = 3;       // stages of the ring (bf16, G = 5, S = 908: 33 KB per CTA -> 6 CTAs per SM, 768 CTAs in one wave)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int G>
__global__ void __launch_bounds__(kT) attn_decode_long_kernel(OfaDecodeArgs a) {
  constexpr int ROWB = HD * (int)sizeof(T);            // bytes of one key row of this head
  constexpr int PITCH = ROWB + 16;
  constexpr int CH = ROWB / 16;                        // 16-byte chunks per row
  constexpr int STAGE = TK * PITCH;
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned char* ring = dsm;                           // [NST][TK][PITCH]
  const int S = a.S, Sp = (S + TK - 1) / TK * TK;
  float* sc = reinterpret_cast<float*>(dsm + NST * STAGE);   // [G][Sp]: bias / key-padding mask, then scores, then probabilities
  __shared__ __align__(16) float qs[G][HD];
  __shared__ float inv_l[G];
  float (*wo)[G][HD] = reinterpret_cast<float(*)[G][HD]>(ring);      // [4][G][HD] partial outputs, over the drained ring
  static_assert(sizeof(float) * (kT / 32) * G * HD <= (size_t)NST * STAGE, "partial outputs must fit the ring");
  const int grp = blockIdx.x, h = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int row0 = grp * G;
  pdl_sync();
  const int krow = a.kv_row ? a.kv_row[grp] : grp;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    qs[g][d] = row0 + g < a.R ? (float)reinterpret_cast<const T*>(a.q)[(size_t)(row0 + g) * a.ldq + h * HD + d] : 0.f;
  }
  const unsigned char* Kb = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.k) + (size_t)krow * a.bsk + h * HD);
  const unsigned char* Vb = a.v ? reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.v) + (size_t)krow * a.bsv + h * HD) : nullptr;
  const size_t ldk_b = (size_t)a.ldk * sizeof(T), ldv_b = (size_t)a.ldv * sizeof(T);
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)krow * a.kpm_stride : nullptr;
  const int ntile = (S + TK - 1) / TK;
  const uint32_t ring_u = smem_u32(ring);
  // copy of tile `tile` of K (which = 0) or V (which = 1) into ring stage `st`; keys beyond S are zero-filled
  auto issue = [&](int which, int tile, int st) {
    const unsigned char* base = which ? Vb : Kb;
    const size_t ld_b = which ? ldv_b : ldk_b;
    for (int e = t; e < TK * CH; e += kT) {
      const int key = e / CH, c = e % CH;
      const int j = tile * TK + key;
      const int ok = j < S;
      cp_async16(ring_u + st * STAGE + key * PITCH + c * 16, base + (size_t)(ok ? j : 0) * ld_b + c * 16, ok ? 16 : 0);
    }
  };
  const int total = Vb && !a.score_out ? 2 * ntile : ntile;     // unified tile stream: K tiles then V tiles
  auto issue_seq = [&](int n) {
    if (n < total) issue(n >= ntile, n >= ntile ? n - ntile : n, n % NST);
    cp_async_commit();
  };
#pragma unroll
  for (int n = 0; n < NST - 1; ++n) issue_seq(n);
  // score buffer <- bias_in (0 without), -inf on padded keys and beyond S: coalesced loads while the first tiles are in flight
  for (int e = t; e < G * Sp; e += kT) {
    const int g = e / Sp, j = e - g * Sp;
    float b = -CUDART_INF_F;
    if (j < S && !(kpm && kpm[j]))
      b = a.bias_in && row0 + g < a.R ? a.bias_in[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j] : 0.f;
    sc[e] = b;
  }
  // ---- pass 1: scores ---------------------------------------------------------------------------------------------------
  const int key = t >> 2, part = t & 3;                // 16 of the 64 dims per thread
  for (int n = 0; n < ntile; ++n) {
    cp_async_wait<NST - 2>();
    __syncthreads();                                   // stage n landed for everybody; stage (n-1) % NST is free again
    issue_seq(n + NST - 1);
    const unsigned char* row = ring + (n % NST) * STAGE + key * PITCH + part * (ROWB / 4);
    float kf[16];
    if constexpr (sizeof(T) == 2) {
      load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(row), *reinterpret_cast<float(*)[8]>(&kf[0]));
      load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(row) + 8, *reinterpret_cast<float(*)[8]>(&kf[8]));
    } else {
      load8<float>(reinterpret_cast<const float*>(row), *reinterpret_cast<float(*)[8]>(&kf[0]));
      load8<float>(reinterpret_cast<const float*>(row) + 8, *reinterpret_cast<float(*)[8]>(&kf[8]));
    }
    const int j = n * TK + key;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 q4 = *reinterpret_cast<const float4*>(&qs[g][part * 16 + c4 * 4]);
        s = fmaf(q4.x, kf[c4 * 4], s); s = fmaf(q4.y, kf[c4 * 4 + 1], s);
        s = fmaf(q4.z, kf[c4 * 4 + 2], s); s = fmaf(q4.w, kf[c4 * 4 + 3], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == g % 4) {                             // (zero-filled keys beyond S: -inf + 0)
        const float b = sc[g * Sp + j];
        if (a.score_out) {
          if (j < S && row0 + g < a.R) a.score_out[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j] = b == -CUDART_INF_F ? 0.f : b + s;
        } else {
          sc[g * Sp + j] = b + s;
        }
      }
    }
  }
  if (a.score_out) { cp_async_wait<0>(); return; }
  __syncthreads();
  // ---- softmax over the score buffer: warp w takes queries w, w + 4 ---------------------------------------------------------
  for (int g = warp; g < G; g += kT / 32) {
    float mx = -CUDART_INF_F;
    for (int j = lane; j < Sp; j += 32) mx = fmaxf(mx, sc[g * Sp + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float mu = mx == -CUDART_INF_F ? 0.f : mx;
    float sum = 0.f;
    for (int j = lane; j < Sp; j += 32) { const float p = __expf(sc[g * Sp + j] - mu); sc[g * Sp + j] = p; sum += p; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) inv_l[g] = sum > 0.f ? 1.f / sum : 0.f;
  }
  // ---- pass 2: out += p v ------------------------------------------------------------------------------------------------
  float o0[G], o1[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { o0[g] = 0.f; o1[g] = 0.f; }
  for (int n = ntile; n < 2 * ntile; ++n) {
    cp_async_wait<NST - 2>();
    __syncthreads();                                   // (the first one also publishes the softmax results)
    issue_seq(n + NST - 1);
    const unsigned char* tile = ring + (n % NST) * STAGE;
    const int j0 = (n - ntile) * TK + warp * (TK / 4);       // this warp's 8 keys of the tile (p = 0 and v = 0 beyond S)
#pragma unroll
    for (int u4 = 0; u4 < TK / 4; u4 += 4) {
      float2 v2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v2[u] = load2<T>(reinterpret_cast<const T*>(tile + (warp * (TK / 4) + u4 + u) * PITCH) + lane * 2);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 p4 = *reinterpret_cast<const float4*>(&sc[g * Sp + j0 + u4]);
        o0[g] = fmaf(p4.x, v2[0].x, o0[g]); o1[g] = fmaf(p4.x, v2[0].y, o1[g]);
        o0[g] = fmaf(p4.y, v2[1].x, o0[g]); o1[g] = fmaf(p4.y, v2[1].y, o1[g]);
        o0[g] = fmaf(p4.z, v2[2].x, o0[g]); o1[g] = fmaf(p4.z, v2[2].y, o1[g]);
        o0[g] = fmaf(p4.w, v2[3].x, o0[g]); o1[g] = fmaf(p4.w, v2[3].y, o1[g]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                     // everybody is done reading the ring
#pragma unroll
  for (int g = 0; g < G; ++g) { wo[warp][g][lane * 2] = o0[g]; wo[warp][g][lane * 2 + 1] = o1[g]; }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    if (row0 + g >= a.R) continue;
    const float oo = (wo[0][g][d] + wo[1][g][d]) + (wo[2][g][d] + wo[3][g][d]);
    reinterpret_cast<T*>(a.o)[(size_t)(row0 + g) * a.ldo + h * HD + d] = (T)(oo * inv_l[g] * cs);
  }
}

// ---- bf16 long-key decode attention on warp-level tensor-core tiles ----------------------------------------------------------
// Same two streamed passes as attn_decode_long_kernel, but the arithmetic of a 16-key tile is 4 (q.k) + 8 (p.v) mma.sync
// m16n8k16 instructions instead of ~1400 scalar ones (the SIMT kernel above spends 81 k warp instructions per CTA and is
// issue-bound at 1.3 TB/s), and the four warps stream their own key tiles (tile i -> warp i % 4) through private two-stage
// cp.async rings: no CTA barrier inside the passes.  The queries (G <= 8 rows) are the N = 8 side of q.k and rows 0..G-1 of the
// M = 16 side of p.v; probabilities go to the tensor core in bf16 (un-normalised exp(s - max) in [0, 1], normalised in fp32).
// One query token per row has nothing for tcgen05 (M = 128 tiles): the kernel is a DRAM stream, the MMAs only get the issue
// slots out of its way.
constexpr int TKW = 16;      // keys per warp tile
constexpr int WPITCH = HD * 2 + 16;
constexpr int WSTAGE = TKW * WPITCH;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int G>
__global__ void __launch_bounds__(kT) attn_decode_mma_kernel(OfaDecodeArgs a) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned char* ring = dsm;                           // [4 warps][2 stages][TKW][WPITCH]
  const int S = a.S, Sp = (S + TKW - 1) / TKW * TKW;
  float* sc = reinterpret_cast<float*>(dsm + (kT / 32) * 2 * WSTAGE);   // [G][Sp]
  float (*qs)[HD] = reinterpret_cast<float(*)[HD]>(ring + WSTAGE);   // [G][HD] q in fp32: warp 0's second stage, until pass 1 starts
  static_assert(sizeof(float) * G * HD <= (size_t)WSTAGE, "q must fit one stage");
  __shared__ float inv_l[G];
  float (*wo)[G][HD] = reinterpret_cast<float(*)[G][HD]>(ring);
  static_assert(sizeof(float) * (kT / 32) * G * HD <= (size_t)(kT / 32) * 2 * WSTAGE, "partial outputs must fit the ring");
  const int grp = blockIdx.x, h = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int row0 = grp * G;
  pdl_sync();
  const int krow = a.kv_row ? a.kv_row[grp] : grp;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    qs[g][d] = row0 + g < a.R ? (float)reinterpret_cast<const T*>(a.q)[(size_t)(row0 + g) * a.ldq + h * HD + d] : 0.f;
  }
  const unsigned char* Kb = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.k) + (size_t)krow * a.bsk + h * HD);
  const unsigned char* Vb = a.v ? reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.v) + (size_t)krow * a.bsv + h * HD) : nullptr;
  const size_t ldk_b = (size_t)a.ldk * sizeof(T), ldv_b = (size_t)a.ldv * sizeof(T);
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)krow * a.kpm_stride : nullptr;
  const int ntile = Sp / TKW;
  const int nw = ntile > warp ? (ntile - warp + 3) / 4 : 0;            // tiles of this warp: warp, warp + 4, ...
  const bool two = Vb && !a.score_out;
  const int total = two ? 2 * nw : nw;
  const uint32_t wring = smem_u32(ring) + warp * 2 * WSTAGE;
  // this lane copies half a row (4 x 16 bytes) of key lane / 2 of the tile
  const int ckey = lane >> 1, cch = (lane & 1) * 4;
  auto issue = [&](int n) {
    if (n < total) {
      const bool isv = n >= nw;
      const int j = (warp + 4 * (isv ? n - nw : n)) * TKW + ckey;
      const int ok = j < S ? 16 : 0;
      const unsigned char* src = (isv ? Vb : Kb) + (size_t)(ok ? j : 0) * (isv ? ldv_b : ldk_b) + cch * 16;
      const uint32_t dst = wring + (n & 1) * WSTAGE + ckey * WPITCH + cch * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) cp_async16(dst + c * 16, src + c * 16, ok);
    }
    cp_async_commit();
  };
  issue(0);
  for (int e = t; e < G * Sp; e += kT) {             // score buffer <- bias_in (0 without), -inf on padded keys and beyond S
    const int g = e / Sp, j = e - g * Sp;
    float b = -CUDART_INF_F;
    if (j < S && !(kpm && kpm[j]))
      b = a.bias_in && row0 + g < a.R ? a.bias_in[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j] : 0.f;
    sc[e] = b;
  }
  __syncthreads();
  // B fragments of q (dims x queries): query lane / 4, dims ks * 16 + 2 * (lane % 4) + {0, 1} and + 8
  const int fr = lane >> 2, fc = (lane & 3) * 2;
  uint32_t bq[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float* qr = &qs[fr < G ? fr : 0][ks * 16 + fc];
    bq[ks][0] = fr < G ? pack_bf16(qr[0], qr[1]) : 0u;
    bq[ks][1] = fr < G ? pack_bf16(qr[8], qr[9]) : 0u;
  }
  __syncthreads();                                     // q is in registers: its stage may be filled
  // ldmatrix lane address inside a tile: row (l & 7) + 8 * ((l >> 3) & 1), column block (l >> 4) * 8 elements
  const uint32_t lm_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * WPITCH + (lane >> 4) * 16);
  // ---- pass 1: scores ---------------------------------------------------------------------------------------------------
  for (int n = 0; n < nw; ++n) {
    __syncwarp();
    issue(n + 1);
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t st = wring + (n & 1) * WSTAGE + lm_off;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t af[4];
      ldmatrix_x4(af, st + ks * 32);
      mma_bf16_16816(c, af, bq[ks][0], bq[ks][1]);
    }
    const int j = (warp + 4 * n) * TKW + fr;           // c[0], c[1]: key j, queries fc, fc + 1;  c[2], c[3]: key j + 8
    if (fc < G) { sc[fc * Sp + j] += c[0]; sc[fc * Sp + j + 8] += c[2]; }
    if (fc + 1 < G) { sc[(fc + 1) * Sp + j] += c[1]; sc[(fc + 1) * Sp + j + 8] += c[3]; }
  }
  __syncthreads();
  if (a.score_out) {
    cp_async_wait<0>();
    for (int e = t; e < G * Sp; e += kT) {
      const int g = e / Sp, j = e - g * Sp;
      if (j < S && row0 + g < a.R) a.score_out[((size_t)(row0 + g) * a.H + h) * a.bias_ld + j] = sc[e] == -CUDART_INF_F ? 0.f : sc[e];
    }
    return;
  }
  // ---- softmax over the score buffer: warp w takes queries w, w + 4 ---------------------------------------------------------
  for (int g = warp; g < G; g += kT / 32) {
    float mx = -CUDART_INF_F;
    for (int j = lane; j < Sp; j += 32) mx = fmaxf(mx, sc[g * Sp + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float mu = mx == -CUDART_INF_F ? 0.f : mx;
    float sum = 0.f;
    for (int j = lane; j < Sp; j += 32) { const float p = __expf(sc[g * Sp + j] - mu); sc[g * Sp + j] = p; sum += p; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) inv_l[g] = sum > 0.f ? 1.f / sum : 0.f;
  }
  __syncthreads();
  // ---- pass 2: out += p v  (A = probabilities: rows = queries, B = V tile through ldmatrix.trans) --------------------------------
  float o[8][4];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) { o[nb][0] = 0.f; o[nb][1] = 0.f; o[nb][2] = 0.f; o[nb][3] = 0.f; }
  for (int n = nw; n < 2 * nw; ++n) {
    __syncwarp();
    issue(n + 1);
    cp_async_wait<1>();
    __syncwarp();
    const int j0 = (warp + 4 * (n - nw)) * TKW;
    uint32_t pa[4] = {0u, 0u, 0u, 0u};
    if (fr < G) {
      const float2 p0 = *reinterpret_cast<const float2*>(&sc[fr * Sp + j0 + fc]);
      const float2 p1 = *reinterpret_cast<const float2*>(&sc[fr * Sp + j0 + fc + 8]);
      pa[0] = pack_bf16(p0.x, p0.y);
      pa[2] = pack_bf16(p1.x, p1.y);
    }
    const uint32_t st = wring + (n & 1) * WSTAGE + lm_off;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {                   // dims nb * 16 .. + 16: two n-blocks of 8
      uint32_t bf[4];
      ldmatrix_x4_trans(bf, st + nb * 32);
      mma_bf16_16816(o[2 * nb], pa, bf[0], bf[1]);
      mma_bf16_16816(o[2 * nb + 1], pa, bf[2], bf[3]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                     // everybody is done with the rings
  if (fr < G) {
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) { wo[warp][fr][nb * 8 + fc] = o[nb][0]; wo[warp][fr][nb * 8 + fc + 1] = o[nb][1]; }
  }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    if (row0 + g >= a.R) continue;
    const float oo = (wo[0][g][d] + wo[1][g][d]) + (wo[2][g][d] + wo[3][g][d]);
    reinterpret_cast<T*>(a.o)[(size_t)(row0 + g) * a.ldo + h * HD + d] = (T)(oo * inv_l[g] * cs);
  }
}

template <int G>
int launch_decode_mma(const OfaDecodeArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(kT / 32) * 2 * WSTAGE + (size_t)G * ((a.S + TKW - 1) / TKW * TKW) * sizeof(float);
  if (smem > 200 * 1024) return -1;
  static bool configured = false;
  if (smem > 48 * 1024 && !configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_mma_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid((a.R + G - 1) / G, a.H);
  return (int)ofa_launch_pdl(attn_decode_mma_kernel<G>, grid, kT, smem, st, a);
}

// ---- bf16 long-key decode attention, one streamed pass (online softmax in registers) --------------------------------------------
// The two-pass kernel above keeps G x S fp32 scores in shared memory (18 KB at G = 5, S = 908: half of the CTA's budget), so
// its rings hold one 2 KB tile in flight per warp and the launch sits at 2.6 TB/s on bytes in flight.  Here a stage is the K
// AND the V tile of 16 keys (4 KB, 128-byte rows XOR-swizzled by key so that ldmatrix is conflict-free without padding), the
// scores never leave registers: S = Q K^T with the queries as the M side (rows 0..G-1 of 16), so the C fragment of a tile is
// already the A fragment of P V (rows = queries, columns = keys), running max / sum per query row, O rescaled per tile.  Three
// warps with three stages each = 36 KB per CTA (six CTAs per SM: the whole grid of 64 x 12 in one wave) with 8 KB per warp in
// flight and no drain between passes: 69 -> 52 us per launch at the captioning shape (ncu: 193.6 MB in 50.2 us = 3.86 TB/s, 59 %
// of the copy peak; four warps x two stages measured the same).  bias_in / key padding of the NEXT tile are fetched into
// registers while the current one is computed -- that fetch is now the top stall (long scoreboard, 53 % of the samples); a
// three-tile-deep register prefetch spilled and was slower.
constexpr int OSTAGE = TKW * HD * 2 * 2;       // K tile + V tile, bf16
constexpr int OW = 3;        // warps per CTA
constexpr int ONST = 3;      // stages per warp: 3 x 3 x 4 KB = 36 KB per CTA, two of three stages (8 KB per warp) in flight

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int G>
__global__ void __launch_bounds__(OW * 32, 6) attn_decode_online_kernel(OfaDecodeArgs a) {
  using T = __nv_bfloat16;
  static_assert(G <= 8, "the queries of a group are rows 0..7 of the M = 16 tile");
  extern __shared__ __align__(128) unsigned char dsm[];
  float (*wo)[G][HD] = reinterpret_cast<float(*)[G][HD]>(dsm);           // partial outputs, over the drained rings
  static_assert(sizeof(float) * OW * G * HD <= (size_t)OW * ONST * OSTAGE, "partial outputs must fit the rings");
  __shared__ float mw[OW][8], lw[OW][8];
  const int grp = blockIdx.x, h = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int row0 = grp * G, S = a.S;
  constexpr float kL2E = 1.4426950408889634f;
  pdl_sync();
  const int krow = a.kv_row ? a.kv_row[grp] : grp;
  const unsigned char* Kb = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.k) + (size_t)krow * a.bsk + h * HD);
  const unsigned char* Vb = reinterpret_cast<const unsigned char*>(reinterpret_cast<const T*>(a.v) + (size_t)krow * a.bsv + h * HD);
  const size_t ldk_b = (size_t)a.ldk * sizeof(T), ldv_b = (size_t)a.ldv * sizeof(T);
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)krow * a.kpm_stride : nullptr;
  const int ntile = (S + TKW - 1) / TKW;
  const int nw = ntile > warp ? (ntile - warp + OW - 1) / OW : 0;      // tiles of this warp: warp, warp + OW, ...
  const uint32_t wring = smem_u32(dsm) + warp * ONST * OSTAGE;
  // this lane copies four 16-byte chunks of the K row and of the V row of key lane / 2; the two lanes of a key take ADJACENT
  // chunks in every instruction, so a warp-wide copy asks L2 for whole 32-byte sectors (cp.async.cg goes to L2 per request:
  // half-sector requests would read every sector twice)
  const int ckey = lane >> 1, cch = lane & 1;
  auto issue = [&](int n) {
    if (n < nw) {
      const int j = (warp + OW * n) * TKW + ckey;
      const int ok = j < S ? 16 : 0;
      const size_t jj = ok ? j : 0;
      const uint32_t dk = wring + (n % ONST) * OSTAGE + ckey * 128, dv = dk + TKW * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ch = 2 * c + cch;
        const uint32_t sw = (uint32_t)((ch ^ (ckey & 7)) * 16);
        cp_async16(dk + sw, Kb + jj * ldk_b + ch * 16, ok);
        cp_async16(dv + sw, Vb + jj * ldv_b + ch * 16, ok);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int n = 0; n < ONST - 1; ++n) issue(n);
  const int fr = lane >> 2, fc = (lane & 3) * 2;
  const bool qrow = fr < G && row0 + fr < a.R;
  // A fragments of q (16 x 64, rows >= G zero): a0 = (row fr, dims ks*16 + fc, +1), a2 = (row fr, dims ks*16 + 8 + fc, +1)
  uint32_t qa[4][4];
  {
    const T* qp = reinterpret_cast<const T*>(a.q) + (size_t)(qrow ? row0 + fr : 0) * a.ldq + h * HD;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = qrow ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + fc) : 0u;
      qa[ks][2] = qrow ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 8 + fc) : 0u;
      qa[ks][1] = 0u; qa[ks][3] = 0u;
    }
  }
  const float* brow = a.bias_in && qrow ? a.bias_in + ((size_t)(row0 + fr) * a.H + h) * a.bias_ld : nullptr;
  // additive term of this thread's four keys of a tile (keys j0 + nb * 8 + fc + e): bias, -inf on padded keys and beyond S
  auto fetch_bias = [&](int n, float (&b)[4]) {
    const int j0 = (warp + OW * n) * TKW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kj = j0 + (i >> 1) * 8 + fc + (i & 1);
      const bool inb = n < nw && kj < S;
      const float bv = inb && brow ? brow[kj] : 0.f;                   // (two independent loads, not mask -> bias)
      const unsigned char pm = inb && kpm ? kpm[kj] : (unsigned char)0;
      b[i] = (!inb || pm) ? -CUDART_INF_F : bv;
    }
  };
  float bn[4];
  fetch_bias(0, bn);
  // ldmatrix lane addressing inside a tile of 128-byte rows, chunk (16 bytes) c of key r stored at chunk c ^ (r & 7)
  const int rowB = (lane & 7) + (lane >> 4) * 8, chB = (lane >> 3) & 1;        // K as the B operand (keys x dims)
  const int rowA = (lane & 7) + ((lane >> 3) & 1) * 8, chA = lane >> 4;        // V through ldmatrix.trans (keys x dims)
  float m = -CUDART_INF_F, l = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
  for (int n = 0; n < nw; ++n) {
    __syncwarp();
    issue(n + ONST - 1);
    float bc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bc[i] = bn[i];
    fetch_bias(n + 1, bn);
    cp_async_wait<ONST - 1>();
    __syncwarp();
    const uint32_t kt = wring + (n % ONST) * OSTAGE, vt = kt + TKW * 128;
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};       // block nb: keys nb * 8 + fc + (e & 1); row fr (e < 2)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bf[4];
      ldmatrix_x4(bf, kt + rowB * 128 + (((chB + 2 * ks) ^ (rowB & 7)) * 16));
      mma_bf16_16816(s[0], qa[ks], bf[0], bf[1]);
      mma_bf16_16816(s[1], qa[ks], bf[2], bf[3]);
    }
    // row fr only (rows 8..15 of the tile are no queries): x = (s + bias) * log2 e
    float x[4];
    x[0] = (s[0][0] + bc[0]) * kL2E; x[1] = (s[0][1] + bc[1]) * kL2E;
    x[2] = (s[1][0] + bc[2]) * kL2E; x[3] = (s[1][1] + bc[3]) * kL2E;
    float tm = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
    tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
    tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
    const float mn = fmaxf(m, tm);
    const float mu = mn == -CUDART_INF_F ? 0.f : mn;
    const float alpha = ex2_approx(m - mu);
    m = mn;
    float p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = ex2_approx(x[i] - mu);
    float ps = (p[0] + p[1]) + (p[2] + p[3]);
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    l = l * alpha + ps;
    uint32_t pa[4];
    pa[0] = pack_bf16(p[0], p[1]); pa[2] = pack_bf16(p[2], p[3]); pa[1] = 0u; pa[3] = 0u;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) { o[nb][0] *= alpha; o[nb][1] *= alpha; }
#pragma unroll
    for (int np = 0; np < 4; ++np) {                   // dims np * 16 .. + 16: two n-blocks of 8
      uint32_t bf[4];
      ldmatrix_x4_trans(bf, vt + rowA * 128 + (((chA + 2 * np) ^ (rowA & 7)) * 16));
      mma_bf16_16816(o[2 * np], pa, bf[0], bf[1]);
      mma_bf16_16816(o[2 * np + 1], pa, bf[2], bf[3]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                     // everybody is done with the rings
  if (fr < G) {
    if ((lane & 3) == 0) { mw[warp][fr] = m; lw[warp][fr] = l; }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) { wo[warp][fr][nb * 8 + fc] = o[nb][0]; wo[warp][fr][nb * 8 + fc + 1] = o[nb][1]; }
  }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < G * HD; e += OW * 32) {
    const int g = e / HD, d = e % HD;
    if (row0 + g >= a.R) continue;
    float mm = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < OW; ++w) mm = fmaxf(mm, mw[w][g]);
    const float mu = mm == -CUDART_INF_F ? 0.f : mm;
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < OW; ++w) {
      const float f = ex2_approx(mw[w][g] - mu);
      ll = fmaf(lw[w][g], f, ll);
      oo = fmaf(wo[w][g][d], f, oo);
    }
    reinterpret_cast<T*>(a.o)[(size_t)(row0 + g) * a.ldo + h * HD + d] = (T)((ll > 0.f ? oo / ll : 0.f) * cs);
  }
}

template <int G>
int launch_decode_online(const OfaDecodeArgs& a, cudaStream_t st) {
  dim3 grid((a.R + G - 1) / G, a.H);
  return (int)ofa_launch_pdl(attn_decode_online_kernel<G>, grid, OW * 32, (size_t)OW * ONST * OSTAGE, st, a);
}

template <typename T, int G>
int launch_decode_long(const OfaDecodeArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)NST * TK * (HD * sizeof(T) + 16) + (size_t)G * ((a.S + TK - 1) / TK * TK) * sizeof(float);
  if (smem > 200 * 1024) return -1;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_long_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = 200 * 1024;
  }
  dim3 grid((a.R + G - 1) / G, a.H);
  return (int)ofa_launch_pdl(attn_decode_long_kernel<T, G>, grid, kT, smem, st, a);
}

template <typename T>
int launch_decode(const OfaDecodeArgs& a, cudaStream_t st) {
  dim3 grid((a.R + a.G - 1) / a.G, a.H);
  // incremental self-attention (one query per row, a key per lane): independent warps
  if (g_decode_short && a.G == 1 && a.S <= 32 && !a.score_out && a.v && a.o && (a.ldo % 2) == 0 && (a.ldq % 2) == 0 &&
      (!a.pq || (a.ldpq % 2) == 0))
    return (int)ofa_launch_pdl(attn_decode_short_kernel<T>, dim3((a.R * a.H + kT / 32 - 1) / (kT / 32)), kT, 0, st, a);
  // long key ranges without the self-attention extras (position keys, rel-pos LUT, pages): the staged two-pass kernel
  if (a.S >= 64 && !a.pk && !a.tok_lut && !a.page_table) {
    int rc = -1;
    switch (a.G) {
#define OFA_DEC_CASE(g) case g: \
        if constexpr (sizeof(T) == 2) \
          rc = (g_decode_online && !a.score_out && a.v && (a.ldq % 2) == 0) ? launch_decode_online<g>(a, st) : launch_decode_mma<g>(a, st); \
        else rc = launch_decode_long<T, g>(a, st); \
        break;
      OFA_DEC_CASE(1) OFA_DEC_CASE(2) OFA_DEC_CASE(3) OFA_DEC_CASE(4) OFA_DEC_CASE(5) OFA_DEC_CASE(6) OFA_DEC_CASE(7) OFA_DEC_CASE(8)
#undef OFA_DEC_CASE
    }
    if (rc >= 0) return rc;      // -1: score buffer too large for shared memory -> the register kernel below
  }
  switch (a.G) {
#define OFA_DEC_CASE(g) case g: return (int)ofa_launch_pdl(attn_decode_kernel<T, g>, grid, kT, 0, st, a);
    OFA_DEC_CASE(1) OFA_DEC_CASE(2) OFA_DEC_CASE(3) OFA_DEC_CASE(4) OFA_DEC_CASE(5) OFA_DEC_CASE(6) OFA_DEC_CASE(7) OFA_DEC_CASE(8)
#undef OFA_DEC_CASE
  }
  return (int)cudaErrorInvalidValue;
}

// dst[p][r][l][:] = src[p][order[r]][l][:]  for l < L; planes p are separated by plane_stride elements
template <typename T>
__global__ void cache_gather_kernel(const T* __restrict__ src, T* __restrict__ dst, const long long* __restrict__ order,
                                    int rows, int L, int D, long long row_stride, long long plane_stride) {
  pdl_sync();
  const int r = blockIdx.x, p = blockIdx.y;
  const long long so = (long long)p * plane_stride + order[r] * row_stride;
  const long long d_o = (long long)p * plane_stride + (long long)r * row_stride;
  const int n8 = L * D / 8;   // positions of a row are contiguous (stride D)
  for (int i = threadIdx.x; i < n8 * (int)(sizeof(T) * 8 / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(dst + d_o)[i] = reinterpret_cast<const uint4*>(src + so)[i];
}

// ---- paged self-attention KV cache (beam search) ----------------------------------------------------------------------
// The cache of the incremental decoder is a pool of pages [slot][plane = 2*layer + (k|v)][page_len][D]; row r reads key j
// through page_table[r][j / page_len].  FULL pages are immutable and shared: a beam reorder only copies table entries (the
// reference index_selects every K / V tensor of every layer at every step, unify_multihead_attention.py:458-480; round 1 of
// this package copied the valid prefix).  The PARTIAL last page is copied on write: its `off` valid positions go to the row's
// own slot for this step's parity (slot = (r*max_pages + page)*2 + parity), which no table references yet.
template <typename T>
__global__ void page_reorder_kernel(T* __restrict__ pool, const int* __restrict__ tab_src, int* __restrict__ tab_dst,
                                    const long long* __restrict__ order, int max_pages, int page, int off, int parity,
                                    int page_len, int D, int planes) {
  pdl_sync();
  const int r = blockIdx.x, pl = blockIdx.y;
  const int parent = order ? (int)order[r] : r;
  const int home = (r * max_pages + page) * 2 + parity;
  if (pl == 0)
    for (int i = threadIdx.x; i < max_pages; i += blockDim.x)
      tab_dst[(size_t)r * max_pages + i] = i < page ? tab_src[(size_t)parent * max_pages + i] : (i == page ? home : -1);
  if (off > 0) {
    const size_t pstride = (size_t)planes * page_len * D;
    const T* src = pool + (size_t)tab_src[(size_t)parent * max_pages + page] * pstride + (size_t)pl * page_len * D;
    T* dst = pool + (size_t)home * pstride + (size_t)pl * page_len * D;
    const int n16 = off * D * (int)sizeof(T) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  }
}

// pool[page_table[r][page]][plane_k | plane_k + 1][off][:] = k[r][:] | v[r][:]   (the new token of every row, one layer)
template <typename T>
__global__ void page_write_kernel(T* __restrict__ pool, const int* __restrict__ tab, const T* __restrict__ k, const T* __restrict__ v,
                                  long long ldk, long long ldv, int max_pages, int page, int off, int page_len, int D, int planes,
                                  int plane_k) {
  pdl_sync();
  const int r = blockIdx.x;
  const size_t pstride = (size_t)planes * page_len * D;
  T* base = pool + (size_t)tab[(size_t)r * max_pages + page] * pstride + (size_t)off * D;
  T* dk = base + (size_t)plane_k * page_len * D;
  T* dv = base + (size_t)(plane_k + 1) * page_len * D;
  const int n16 = D * (int)sizeof(T) / 16;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {
    reinterpret_cast<uint4*>(dk)[i] = reinterpret_cast<const uint4*>(k + (size_t)r * ldk)[i];
    reinterpret_cast<uint4*>(dv)[i] = reinterpret_cast<const uint4*>(v + (size_t)r * ldv)[i];
  }
}

}  // namespace

extern "C" int ofa_page_reorder(void* pool, const int* tab_src, int* tab_dst, const long long* order, int rows, int max_pages,
                                int page, int off, int parity, int page_len, int D, int planes, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && max_pages > 0 && page >= 0 && page < max_pages && off >= 0 && off < page_len && planes > 0,
            "ofa_page_reorder: bad arguments");
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  OFA_CHECK(((long long)D * esz) % 16 == 0 && ((uintptr_t)pool & 15) == 0, "ofa_page_reorder: rows must be 16-byte aligned");
  dim3 grid(rows, planes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(page_reorder_kernel<__nv_bfloat16>, grid, 128, 0, st, (__nv_bfloat16*)pool, tab_src, tab_dst, order, max_pages, page, off, parity, page_len, D, planes));
  else
    OFA_CUDA(ofa_launch_pdl(page_reorder_kernel<float>, grid, 128, 0, st, (float*)pool, tab_src, tab_dst, order, max_pages, page, off, parity, page_len, D, planes));
  OFA_LAUNCH_CHECK("page_reorder_kernel");
  return 0;
}

extern "C" int ofa_page_write(void* pool, const int* tab, const void* k, const void* v, long long ldk, long long ldv, int rows,
                              int max_pages, int page, int off, int page_len, int D, int planes, int plane_k, int dtype,
                              void* stream) {
  OFA_CHECK(rows > 0 && page >= 0 && page < max_pages && off >= 0 && off < page_len && plane_k >= 0 && plane_k + 1 < planes,
            "ofa_page_write: bad arguments");
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  OFA_CHECK(((long long)D * esz) % 16 == 0 && (ldk * esz) % 16 == 0 && (ldv * esz) % 16 == 0 &&
                (((uintptr_t)pool | (uintptr_t)k | (uintptr_t)v) & 15) == 0, "ofa_page_write: rows must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(page_write_kernel<__nv_bfloat16>, dim3(rows), 128, 0, st, (__nv_bfloat16*)pool, tab, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldk, ldv, max_pages, page, off, page_len, D, planes, plane_k));
  else
    OFA_CUDA(ofa_launch_pdl(page_write_kernel<float>, dim3(rows), 128, 0, st, (float*)pool, tab, (const float*)k, (const float*)v, ldk, ldv, max_pages, page, off, page_len, D, planes, plane_k));
  OFA_LAUNCH_CHECK("page_write_kernel");
  return 0;
}

extern "C" int ofa_attn_decode_set_online(int on) {
  const int old = g_decode_online;
  g_decode_online = on;
  return old;
}

extern "C" int ofa_attn_decode_set_short(int on) {
  const int old = g_decode_short;
  g_decode_short = on;
  return old;
}

// see include/ofa_b200.h
extern "C" int ofa_attn_decode(const OfaDecodeArgs* a, int dtype, void* stream) {
  OFA_CHECK(a->R > 0 && a->S > 0 && a->H > 0 && a->G >= 1 && a->G <= GMAX, "ofa_attn_decode: bad sizes R=%d S=%d H=%d G=%d",
            a->R, a->S, a->H, a->G);
  OFA_CHECK(a->q && a->k && (a->score_out || (a->v && a->o)), "ofa_attn_decode: null operand");
  OFA_CHECK((a->pq == nullptr) == (a->pk == nullptr), "ofa_attn_decode: pq and pk go together");
  OFA_CHECK(!(a->bias_in || a->score_out) || a->bias_ld >= a->S, "ofa_attn_decode: bias_ld=%lld < S=%d", a->bias_ld, a->S);
  OFA_CHECK(!a->page_table || (a->page_len > 0 && a->max_pages * a->page_len >= a->S), "ofa_attn_decode: page table too short");
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  OFA_CHECK((a->ldk * esz) % 16 == 0 && (a->bsk * esz) % 16 == 0 && ((uintptr_t)a->k & 15) == 0 && (a->page_stride * esz) % 16 == 0 &&
                (!a->pk || ((a->ldpk * esz) % 16 == 0 && (a->bspk * esz) % 16 == 0 && ((uintptr_t)a->pk & 15) == 0)) &&
                (a->ldv * esz) % 8 == 0 && (a->bsv * esz) % 8 == 0,
            "ofa_attn_decode: cache rows must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == OFA_BF16) rc = launch_decode<__nv_bfloat16>(*a, st);
  else if (dtype == OFA_F32) rc = launch_decode<float>(*a, st);
  else return ofa_set_error("ofa_attn_decode: bad dtype %d", dtype);
  if (rc != 0) return ofa_set_error("ofa_attn_decode: launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  OFA_LAUNCH_CHECK("attn_decode_kernel");
  return 0;
}

extern "C" int ofa_cache_gather(const void* src, void* dst, const long long* order, int rows, int L, int D,
                                long long row_stride, long long plane_stride, int planes, int dtype, void* stream) {
  OFA_CHECK(rows > 0 && L > 0 && planes > 0, "ofa_cache_gather: empty problem");
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  OFA_CHECK(((long long)D * esz) % 16 == 0 && (row_stride * esz) % 16 == 0 && (plane_stride * esz) % 16 == 0 &&
                (((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "ofa_cache_gather: rows must be 16-byte aligned");
  dim3 grid(rows, planes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(cache_gather_kernel<__nv_bfloat16>, grid, 128, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, order, rows, L, D,
                                                             row_stride, plane_stride));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(cache_gather_kernel<float>, grid, 128, 0, st, (const float*)src, (float*)dst, order, rows, L, D, row_stride, plane_stride));
  else
    return ofa_set_error("ofa_cache_gather: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("cache_gather_kernel");
  return 0;
}
