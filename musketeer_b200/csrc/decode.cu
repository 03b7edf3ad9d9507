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

template <typename T>
int launch_decode(const OfaDecodeArgs& a, cudaStream_t st) {
  dim3 grid((a.R + a.G - 1) / a.G, a.H);
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
