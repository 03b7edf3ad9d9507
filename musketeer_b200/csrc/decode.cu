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
  const void* q;  const void* pq;  long long ldq, ldpq;          // [R, *] one token per row (q, pos_q pre-scaled)
  const void* k;  const void* v;   const void* pk;               // caches: (row, j, h*64 + d) at row*bs + j*ld + h*64 + d
  long long ldk, bsk, ldv, bsv, ldpk, bspk;
  const int* kv_row;                                             // [R / G] cache row of each group (null: group index)
  const int* pk_row;                                             // [R / G] cache row of pos_k (null: kv_row)
  const unsigned char* kpm;  long long kpm_stride;               // [cache rows, >= S] 1 = padded key (null: none)
  void* o;  long long ldo;                                       // [R, H*64]
  const float* head_scale;                                       // [H] c_attn or null
  const float* tok_lut;  int tok_max;  int q_pos;                // self-attention rel-pos: lut[h][(q_pos - j) + tok_max - 1]
  int R, G, H, S;
};

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
__global__ void __launch_bounds__(kT) attn_decode_kernel(OfaDecodeArgs a) {
  pdl_sync();
  extern __shared__ float smem[];
  float* qs = smem;                     // [G][128]  q | pos_q of the group's rows, this head
  float* ps = smem + a.G * 128;         // [G][S]    scores, then probabilities
  __shared__ float red[GMAX][kT / 32];
  __shared__ float stat[GMAX];
  __shared__ float oacc[kT / 32][GMAX][HD];
  const int grp = blockIdx.x, h = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int G = a.G, S = a.S;
  const int row0 = grp * G;
  const int krow = a.kv_row ? a.kv_row[grp] : grp;
  const int prow = a.pk_row ? a.pk_row[grp] : krow;
  for (int e = t; e < G * 128; e += kT) {
    const int g = e >> 7, d = e & 127;
    float v = 0.f;
    if (row0 + g < a.R)
      v = d < HD ? (float)reinterpret_cast<const T*>(a.q)[(size_t)(row0 + g) * a.ldq + h * HD + d]
                 : (float)reinterpret_cast<const T*>(a.pq)[(size_t)(row0 + g) * a.ldpq + h * HD + d - HD];
    qs[e] = v;
  }
  __syncthreads();
  const T* K = reinterpret_cast<const T*>(a.k) + (size_t)krow * a.bsk + h * HD;
  const T* PK = reinterpret_cast<const T*>(a.pk) + (size_t)prow * a.bspk + h * HD;
  const T* V = reinterpret_cast<const T*>(a.v) + (size_t)krow * a.bsv + h * HD;
  const unsigned char* kpm = a.kpm ? a.kpm + (size_t)krow * a.kpm_stride : nullptr;
  const float* lut = a.tok_lut ? a.tok_lut + (size_t)h * (2 * a.tok_max - 1) : nullptr;

  // phase 1: scores; one key per thread and iteration, 128-dim dot product against every query of the group
  float mx[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) mx[g] = -CUDART_INF_F;
  for (int j = t; j < S; j += kT) {
    float acc[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
    const bool masked = kpm && kpm[j];
    if (!masked) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const T* src = (half ? PK + (size_t)j * a.ldpk : K + (size_t)j * a.ldk);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float kv[8];
          load8<T>(src + c * 8, kv);
#pragma unroll
          for (int g = 0; g < GMAX; ++g) {
            if (g < G) {
              const float* qq = qs + g * 128 + half * HD + c * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[g] = fmaf(qq[e], kv[e], acc[g]);
            }
          }
        }
      }
    }
    float bias = 0.f;
    if (lut) {
      const int rel = a.q_pos - j + a.tok_max - 1;
      if (rel >= 0 && rel < 2 * a.tok_max - 1) bias = lut[rel];
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
      if (g < G) {
        const float s = masked ? -CUDART_INF_F : acc[g] + bias;
        ps[g * S + j] = s;
        mx[g] = fmaxf(mx[g], s);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    if (g < G) {
      float m = mx[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) red[g][warp] = m;
    }
  }
  __syncthreads();
  if (t < G) {
    float m = red[t][0];
    for (int w = 1; w < kT / 32; ++w) m = fmaxf(m, red[t][w]);
    stat[t] = m == -CUDART_INF_F ? 0.f : m;
  }
  __syncthreads();
  // phase 2: exponentials and row sums
  float sum[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) sum[g] = 0.f;
  for (int j = t; j < S; j += kT) {
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
      if (g < G) {
        const float e = __expf(ps[g * S + j] - stat[g]);
        ps[g * S + j] = e;
        sum[g] += e;
      }
    }
  }
  __syncthreads();   // every thread has read stat[] and written its probabilities
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    if (g < G) {
      float s = sum[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[g][warp] = s;
    }
  }
  __syncthreads();
  if (t < G) {
    float s = 0.f;
    for (int w = 0; w < kT / 32; ++w) s += red[t][w];
    stat[t] = s;
  }
  // phase 3: out[g][d] = sum_j p[g][j] V[j][d]; lane = pair of output dims, warp = key residue class
  float o0[GMAX], o1[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) { o0[g] = 0.f; o1[g] = 0.f; }
  // eight keys per iteration: the eight V loads are issued together (the loop is L2-latency-bound otherwise)
  constexpr int U = 8, NW = kT / 32;
  for (int j0 = warp; j0 < S; j0 += U * NW) {
    float v0[U], v1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NW;
      v0[u] = 0.f; v1[u] = 0.f;
      if (j < S) {
        if (sizeof(T) == 2) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(
              reinterpret_cast<const __nv_bfloat16*>(V) + (size_t)j * a.ldv + lane * 2));
          v0[u] = f.x; v1[u] = f.y;
        } else {
          const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(V) + (size_t)j * a.ldv + lane * 2);
          v0[u] = f.x; v1[u] = f.y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NW;
      if (j < S) {
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
          if (g < G) {
            const float p = ps[g * S + j];
            o0[g] = fmaf(p, v0[u], o0[g]);
            o1[g] = fmaf(p, v1[u], o1[g]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    if (g < G) { oacc[warp][g][lane * 2] = o0[g]; oacc[warp][g][lane * 2 + 1] = o1[g]; }
  }
  __syncthreads();
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  for (int e = t; e < G * HD; e += kT) {
    const int g = e / HD, d = e % HD;
    if (row0 + g >= a.R) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) s += oacc[w][g][d];
    const float l = stat[g];
    s = (l > 0.f ? s / l : 0.f) * cs;
    reinterpret_cast<T*>(a.o)[(size_t)(row0 + g) * a.ldo + h * HD + d] = (T)s;
  }
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

}  // namespace

// see include/ofa_b200.h
extern "C" int ofa_attn_decode(const OfaDecodeArgs* a, int dtype, void* stream) {
  OFA_CHECK(a->R > 0 && a->S > 0 && a->H > 0 && a->G >= 1 && a->G <= GMAX, "ofa_attn_decode: bad sizes R=%d S=%d H=%d G=%d",
            a->R, a->S, a->H, a->G);
  OFA_CHECK(a->q && a->pq && a->k && a->pk && a->v && a->o, "ofa_attn_decode: null operand");
  const int esz = dtype == OFA_BF16 ? 2 : 4;
  OFA_CHECK((a->ldk * esz) % 16 == 0 && (a->ldpk * esz) % 16 == 0 && (a->bsk * esz) % 16 == 0 && (a->bspk * esz) % 16 == 0 &&
                (((uintptr_t)a->k | (uintptr_t)a->pk) & 15) == 0 && (a->ldv * esz) % 8 == 0 && (a->bsv * esz) % 8 == 0,
            "ofa_attn_decode: cache rows must be 16-byte aligned");
  const size_t smem = (size_t)a->G * (128 + a->S) * sizeof(float);
  OFA_CHECK(smem <= 160 * 1024, "ofa_attn_decode: G*S=%d too large for the shared-memory score buffer", a->G * a->S);
  dim3 grid((a->R + a->G - 1) / a->G, a->H);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16) {
    static bool configured = false;
    if (!configured) {
      OFA_CUDA(cudaFuncSetAttribute(attn_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured = true;
    }
    OFA_CUDA(ofa_launch_pdl(attn_decode_kernel<__nv_bfloat16>, grid, kT, smem, st, *a));
  } else if (dtype == OFA_F32) {
    static bool configured = false;
    if (!configured) {
      OFA_CUDA(cudaFuncSetAttribute(attn_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured = true;
    }
    OFA_CUDA(ofa_launch_pdl(attn_decode_kernel<float>, grid, kT, smem, st, *a));
  } else {
    return ofa_set_error("ofa_attn_decode: bad dtype %d", dtype);
  }
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
