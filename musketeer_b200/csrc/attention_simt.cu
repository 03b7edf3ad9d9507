// Exact-precision (fp32 math on the SIMT cores) OFA attention, forward and backward.  This is the "fp32 parity mode"
// path (north_star: logits within 1e-4, tokens bit-exact in fp32 mode) and the on-device cross-check for the tcgen05
// flash kernels in attention_tc.cu, which serve the bf16 mode.
//
//   S[i,j] = q_i.k_j + pq_i.pk_j + tokLUT[h][(i_t - j_t) + 1023] (text x text) + imgLUT[h][bucket(pid_i, pid_j)] (img x img)
//   S[i,j] = -inf  for padded keys and (causal) j > i ;  P = softmax_fp32(S) ;  O = c_attn[h] * P V
//   restates models/ofa/unify_multihead_attention.py:345-398 with the bias assembly of
//   models/ofa/unify_transformer.py:640-658,906-933 (encoder) and :1282-1318,1519-1529 (decoder) folded in:
//   the [B,H,N,N] bias tensor is never materialised.
// q/k/v/pos tensors are [B, L, H, 64] views (row stride ld*, batch stride bs*), q and pq already scaled.
#include <math_constants.h>

#include "attention_common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kT = 128;

template <typename T>
__device__ __forceinline__ float tof(T v) { return (float)v; }

__device__ __forceinline__ float blk_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
}
__device__ __forceinline__ float blk_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1] + sh[2] + sh[3];
}

// scores of query row i against all keys into sc[0..S) (fp32, bias and masks applied)
template <typename T>
__device__ void score_row(const AttnArgs& a, int b, int h, int i, const float* qrow /*smem[128]*/, float* sc) {
  const T* K = (const T*)a.k + (size_t)b * a.bsk + h * HD;
  const T* PK = (const T*)a.pk + (size_t)b * a.bspk + h * HD;
  const AttnBias& bz = a.bias;
  const int iabs = i + a.q_pos_off;
  for (int j = threadIdx.x; j < a.S; j += kT) {
    const T* kr = K + (size_t)j * a.ldk;
    const T* pr = PK + (size_t)j * a.ldpk;
    float s = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) s += qrow[d] * tof(kr[d]);
#pragma unroll 8
    for (int d = 0; d < HD; ++d) s += qrow[HD + d] * tof(pr[d]);
    s += attn_bias_at(bz, b, h, iabs, j);
    if (a.causal && j > iabs) s = -CUDART_INF_F;
    if (a.kpm && a.kpm[(size_t)b * a.S + j]) s = -CUDART_INF_F;
    sc[j] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(kT) attn_fwd_simt(AttnArgs a) {
  extern __shared__ float smf[];
  float* qrow = smf;        // 128
  float* red = smf + 128;   // 4
  float* part = smf + 132;  // 128 (two half-sums of 64)
  float* sc = smf + 260;    // S
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* Q = (const T*)a.q + (size_t)b * a.bsq + (size_t)i * a.ldq + h * HD;
  const T* PQ = (const T*)a.pq + (size_t)b * a.bspq + (size_t)i * a.ldpq + h * HD;
  qrow[threadIdx.x] = threadIdx.x < HD ? tof(Q[threadIdx.x]) : tof(PQ[threadIdx.x - HD]);
  __syncthreads();
  score_row<T>(a, b, h, i, qrow, sc);
  __syncthreads();
  float m = -CUDART_INF_F;
  for (int j = threadIdx.x; j < a.S; j += kT) m = fmaxf(m, sc[j]);
  m = blk_max(m, red);
  const float mu = (m == -CUDART_INF_F) ? 0.f : m;
  float l = 0.f;
  for (int j = threadIdx.x; j < a.S; j += kT) {
    const float p = expf(sc[j] - mu);
    sc[j] = p;
    l += p;
  }
  l = blk_sum(l, red);
  const float inv = l > 0.f ? 1.f / l : 0.f;
  // O[d] = sum_j p_j v[j][d]: threads 0-63 take even j, 64-127 odd j
  const int d = threadIdx.x & 63, par = threadIdx.x >> 6;
  const T* Vp = (const T*)a.v + (size_t)b * a.bsv + h * HD + d;
  float o = 0.f;
  for (int j = par; j < a.S; j += 2) {
    float p = sc[j] * inv;
    if (a.p_round_bf16) p = __bfloat162float(__float2bfloat16(p));
    o += p * tof(Vp[(size_t)j * a.ldv]);
  }
  part[threadIdx.x] = o;
  __syncthreads();
  if (threadIdx.x < HD) {
    const float cs = a.head_scale ? a.head_scale[h] : 1.f;
    T* O = (T*)a.o + (size_t)b * a.bso + (size_t)i * a.ldo + h * HD;
    O[threadIdx.x] = (T)((part[threadIdx.x] + part[threadIdx.x + 64]) * cs);
    if (threadIdx.x == 0) a.lse[((size_t)b * a.H + h) * a.T + i] = mu + logf(l);
  }
}

// backward pass 1: one CTA per query row.  Recomputes P, forms dS = P o (dP - delta), writes dQ', materialises
// P and dS rows (fp32 workspaces [B,H,T,S]) for pass 2 and accumulates the rel-pos table gradients with fp32 REDs.
template <typename T>
__global__ void __launch_bounds__(kT) attn_bwd_simt_q(AttnArgs a, AttnGrads g) {
  extern __shared__ float smf[];
  float* qrow = smf;       // 128
  float* red = smf + 128;  // 4
  float* dorow = smf + 132;  // 64
  float* part = smf + 196;   // 128
  float* sc = smf + 324;     // S
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* Q = (const T*)a.q + (size_t)b * a.bsq + (size_t)i * a.ldq + h * HD;
  const T* PQ = (const T*)a.pq + (size_t)b * a.bspq + (size_t)i * a.ldpq + h * HD;
  qrow[threadIdx.x] = threadIdx.x < HD ? tof(Q[threadIdx.x]) : tof(PQ[threadIdx.x - HD]);
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  float dl = 0.f;
  if (threadIdx.x < HD) {
    const size_t off = (size_t)b * a.bso + (size_t)i * a.ldo + h * HD + threadIdx.x;
    const float dout = tof(((const T*)g.dout)[off]);
    dorow[threadIdx.x] = dout * cs;                   // dO w.r.t. the unscaled P.V
    dl = dout * tof(((const T*)a.o)[off]);            // delta_i = sum_d dOut * Out
  }
  const float delta = blk_sum(dl, red);
  if (threadIdx.x == 0) g.delta[((size_t)b * a.H + h) * a.T + i] = delta;
  score_row<T>(a, b, h, i, qrow, sc);
  __syncthreads();
  const float lse = a.lse[((size_t)b * a.H + h) * a.T + i];
  const T* Vb = (const T*)a.v + (size_t)b * a.bsv + h * HD;
  float* Prow = g.P + (((size_t)b * a.H + h) * a.T + i) * a.S;
  float* dSrow = g.dS + (((size_t)b * a.H + h) * a.T + i) * a.S;
  const int iabs = i + a.q_pos_off;
  for (int j = threadIdx.x; j < a.S; j += kT) {
    const float s = sc[j];
    const float p = (s == -CUDART_INF_F) ? 0.f : expf(s - lse);
    const T* vr = Vb + (size_t)j * a.ldv;
    float dp = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) dp += dorow[d] * tof(vr[d]);
    const float ds = p * (dp - delta);
    Prow[j] = a.p_round_bf16 ? __bfloat162float(__float2bfloat16(p)) : p;
    dSrow[j] = ds;
    sc[j] = ds;
    attn_bias_grad_at(a.bias, g, b, h, iabs, j, ds);
  }
  __syncthreads();
  // dQ'[d] = sum_j dS_j K'[j][d]   (d < 64: k, d >= 64: pos_k)
  {
    const int d = threadIdx.x;
    const T* base = d < HD ? (const T*)a.k + (size_t)b * a.bsk + h * HD + d : (const T*)a.pk + (size_t)b * a.bspk + h * HD + (d - HD);
    const long long ldx = d < HD ? a.ldk : a.ldpk;
    float acc = 0.f;
    for (int j = 0; j < a.S; ++j) acc += sc[j] * tof(base[(size_t)j * ldx]);
    if (d < HD) ((T*)g.dq)[(size_t)b * g.bsdq + (size_t)i * g.lddq + h * HD + d] = (T)acc;
    else ((T*)g.dpq)[(size_t)b * g.bsdpq + (size_t)i * g.lddpq + h * HD + (d - HD)] = (T)acc;
  }
  (void)part;
}

// backward pass 2: one CTA per key row.  dK'[j] = sum_i dS[i,j] Q'[i],  dV[j] = sum_i P[i,j] dO[i]
template <typename T>
__global__ void __launch_bounds__(kT) attn_bwd_simt_kv(AttnArgs a, AttnGrads g) {
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int d = threadIdx.x;
  const float cs = a.head_scale ? a.head_scale[h] : 1.f;
  const float* Pcol = g.P + ((size_t)b * a.H + h) * a.T * a.S + j;
  const float* dScol = g.dS + ((size_t)b * a.H + h) * a.T * a.S + j;
  const T* qb = d < HD ? (const T*)a.q + (size_t)b * a.bsq + h * HD + d : (const T*)a.pq + (size_t)b * a.bspq + h * HD + (d - HD);
  const long long ldx = d < HD ? a.ldq : a.ldpq;
  float dk = 0.f, dv = 0.f;
  for (int i = 0; i < a.T; ++i) {
    dk += dScol[(size_t)i * a.S] * tof(qb[(size_t)i * ldx]);
    if (d < HD) dv += Pcol[(size_t)i * a.S] * tof(((const T*)g.dout)[(size_t)b * a.bso + (size_t)i * a.ldo + h * HD + d]);
  }
  if (d < HD) {
    ((T*)g.dk)[(size_t)b * g.bsdk + (size_t)j * g.lddk + h * HD + d] = (T)dk;
    ((T*)g.dv)[(size_t)b * g.bsdv + (size_t)j * g.lddv + h * HD + d] = (T)(dv * cs);
  } else {
    ((T*)g.dpk)[(size_t)b * g.bsdpk + (size_t)j * g.lddpk + h * HD + (d - HD)] = (T)dk;
  }
}

}  // namespace

extern "C" int ofa_attn_fwd_simt(const AttnArgs* a, int dtype, void* stream) {
  OFA_CHECK(a->T > 0 && a->S > 0 && a->B > 0 && a->H > 0, "ofa_attn_fwd_simt: empty problem");
  const size_t smem = (260 + (size_t)a->S) * sizeof(float);
  OFA_CHECK(smem <= 48 * 1024, "ofa_attn_fwd_simt: S=%d too long for the SIMT path", a->S);
  dim3 grid(a->T, a->H, a->B);
  if (dtype == OFA_BF16) attn_fwd_simt<__nv_bfloat16><<<grid, kT, smem, (cudaStream_t)stream>>>(*a);
  else if (dtype == OFA_F32) attn_fwd_simt<float><<<grid, kT, smem, (cudaStream_t)stream>>>(*a);
  else return ofa_set_error("ofa_attn_fwd_simt: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("attn_fwd_simt");
  return 0;
}

extern "C" int ofa_attn_bwd_simt(const AttnArgs* a, const AttnGrads* g, int dtype, void* stream) {
  OFA_CHECK(a->T > 0 && a->S > 0 && a->B > 0 && a->H > 0, "ofa_attn_bwd_simt: empty problem");
  OFA_CHECK(g->P && g->dS && g->delta, "ofa_attn_bwd_simt: workspaces P, dS [B,H,T,S] and delta [B,H,T] are required");
  const size_t smem = (324 + (size_t)a->S) * sizeof(float);
  OFA_CHECK(smem <= 48 * 1024, "ofa_attn_bwd_simt: S=%d too long for the SIMT path", a->S);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 gq(a->T, a->H, a->B), gk(a->S, a->H, a->B);
  if (dtype == OFA_BF16) {
    attn_bwd_simt_q<__nv_bfloat16><<<gq, kT, smem, st>>>(*a, *g);
    attn_bwd_simt_kv<__nv_bfloat16><<<gk, kT, 0, st>>>(*a, *g);
  } else if (dtype == OFA_F32) {
    attn_bwd_simt_q<float><<<gq, kT, smem, st>>>(*a, *g);
    attn_bwd_simt_kv<float><<<gk, kT, 0, st>>>(*a, *g);
  } else {
    return ofa_set_error("ofa_attn_bwd_simt: bad dtype %d", dtype);
  }
  OFA_LAUNCH_CHECK("attn_bwd_simt");
  return 0;
}
