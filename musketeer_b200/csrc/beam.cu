// One beam-search step after the decoder: last-token logits -> (temperature, constraint masks) -> fp32 log-softmax ->
// (min-len / max-len / pad / unk / n-gram / zero-shot masks) -> + cumulative beam scores -> top 2*beam candidates per sentence.
// Replaces, per step, the ~15 ATen kernels of models/sequence_generator.py:352-437 (`forward_decoder` tail :852-889 incl. the
// per-row Python trie walk `constraint_trie.get_next_layer(tokens.tolist())`, the masks :381-404, NGramRepeatBlock :425-426 --
// a host loop over `tokens.cpu()` in fairseq -- and `BeamSearch.step` = models/search.py:119-144: topk over beam*V) by two
// launches that read the logits once:
//   beam_row_kernel   : one CTA per beam row.  Pass A: max / sum-exp over the softmax domain (whole vocabulary, a constraint
//                       range, or the children of the row's trie node).  Pass B: every thread keeps its K best
//                       (log-prob + previous score) in registers; banned tokens (pad, blocked eos, repeated n-grams, post-softmax
//                       constraints) are a bit mask over the vocabulary in shared memory; K rounds of block arg-max emit the
//                       row's K best, sorted.
//   beam_merge_kernel : one warp per sentence merges beam x K candidates into the K best (value descending, flat index
//                       ascending among equals), as int64 flat indices beam * V + token like torch.topk over [bsz, beam * V].
//   trie_advance_kernel: per-row trie node after the beam reorder (node of the parent beam advanced by the chosen token).
// HBM/L2-bound: R * V logits read twice from L2 (they were just written by the output GEMM).
#include <math_constants.h>

#include "common.cuh"

struct OfaBeamArgs {   // mirrored by musketeer_b200/_lib.py and include/ofa_b200.h
  const void* logits; long long ld; int dtype;
  int R, beam, V, K;
  float temperature;
  const float* prev_scores;
  int step0;
  int eos, pad, unk; float unk_penalty;
  int block_eos, force_eos, eos_one;
  int range_lo, range_hi, range_post;
  const int* trie_ptr; const int* trie_tok; const int* node; int trie_post;
  const long long* tokens; long long ldtok; int step; int ngram;
  const long long* prefix_tok; const float* prefix_fill;
  float* row_val; int* row_idx;
  float* cand_scores; long long* cand_index;
};

namespace {

constexpr int kT = 256;
constexpr int KMAX = 16;
constexpr int CAP = 1024;     // candidate list of the threshold path

template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i) { return (float)p[i]; }

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// insert (v, i) into a descending sorted list of K entries held in registers
template <int K>
__device__ __forceinline__ void insert(float (&val)[K], int (&idx)[K], float v, int i) {
  if (!better(v, i, val[K - 1], idx[K - 1])) return;
  val[K - 1] = v; idx[K - 1] = i;
#pragma unroll
  for (int k = K - 1; k > 0; --k) {
    if (better(val[k], idx[k], val[k - 1], idx[k - 1])) {
      const float tv = val[k]; val[k] = val[k - 1]; val[k - 1] = tv;
      const int ti = idx[k]; idx[k] = idx[k - 1]; idx[k - 1] = ti;
    }
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(kT) beam_row_kernel(OfaBeamArgs a) {
  pdl_sync();
  extern __shared__ uint32_t ban[];                 // bit per vocabulary entry: excluded from the candidates
  __shared__ float red[kT / 32], reds[kT / 32];
  __shared__ float cv[kT / 32][K];
  __shared__ int ci[kT / 32][K];
  __shared__ float bcast;
  __shared__ float tmax[kT];
  __shared__ float lv[CAP];
  __shared__ int li[CAP];
  __shared__ int cnt, pfound;
  __shared__ int redi[kT / 32];
  const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int V = a.V;
  float* out_v = a.row_val + (size_t)r * K;
  int* out_i = a.row_idx + (size_t)r * K;
  if (a.step0 && (r % a.beam) != 0) {               // step 0: every beam of a sentence is the same hypothesis (search.py:128-131)
    if (t < K) { out_v[t] = -CUDART_INF_F; out_i[t] = 0x7fffffff; }
    return;
  }
  const T* x = reinterpret_cast<const T*>(a.logits) + (size_t)r * a.ld;
  const float inv_t = 1.f / a.temperature;
  const int nw = (V + 31) / 32;
  // allowed list of the trie node (-1 = dead prefix -> [eos]; utils/trie.py:23-30)
  const int* al = nullptr;
  int nal = -1;
  int eos_only = 0;
  if (a.node) {
    const int nd = a.node[r];
    if (nd == -2) nal = -1;                         // a row still inside its forced prefix: every token allowed (:867-868)
    else if (nd < 0) eos_only = 1, nal = 1;
    else { al = a.trie_tok + a.trie_ptr[nd]; nal = a.trie_ptr[nd + 1] - a.trie_ptr[nd]; }
  }
  // forced prefix token of this row (:604-613): every other entry takes the fill value
  const long long ptok = a.prefix_tok ? a.prefix_tok[r] : (long long)a.pad;
  const bool has_prefix = ptok != (long long)a.pad;
  const float pfill = has_prefix && a.prefix_fill ? *a.prefix_fill : -CUDART_INF_F;
  // the V - 1 filler entries of such a row tie: torch.topk takes them in no defined order; the highest indices go first here
  // (the lowest ones would put eos among the candidates of every prefix step)
  const bool flip = has_prefix && pfill > -CUDART_INF_F;
  const bool pre_list = nal >= 0 && !a.trie_post;
  const bool pre_range = a.range_lo >= 0 && !a.range_post;
  // ---- ban mask -----------------------------------------------------------------------------------------------------
  // (a forced-prefix row overrides every entry but its prefix token AFTER the post-softmax masks, :372-380 after :878-889: only
  // the prefix token itself can still be masked by them)
  const bool post_list = nal >= 0 && a.trie_post && !has_prefix;
  if (t == 0) pfound = 0;
  for (int w = t; w < nw; w += kT) ban[w] = post_list ? 0xffffffffu : 0u;
  __syncthreads();
  if (post_list) {
    for (int e = t; e < nal; e += kT) { const int tok = eos_only ? a.eos : al[e]; atomicAnd(&ban[tok >> 5], ~(1u << (tok & 31))); }
    __syncthreads();
  } else if (has_prefix && nal >= 0 && a.trie_post) {
    for (int e = t; e < nal; e += kT) if ((eos_only ? a.eos : al[e]) == (int)ptok) pfound = 1;
    __syncthreads();
    if (t == 0 && !pfound) atomicOr(&ban[(int)ptok >> 5], 1u << ((int)ptok & 31));
  }
  if (a.range_lo >= 0 && a.range_post) {
    if (has_prefix) {
      if (t == 0 && ptok >= 4 && (ptok < a.range_lo || ptok >= a.range_hi)) atomicOr(&ban[(int)ptok >> 5], 1u << ((int)ptok & 31));
    } else {
      for (int v = 4 + t; v < V; v += kT) if (v < a.range_lo || v >= a.range_hi) atomicOr(&ban[v >> 5], 1u << (v & 31));
    }
  }
  if (t == 0) {
    atomicOr(&ban[a.pad >> 5], 1u << (a.pad & 31));
    if (a.block_eos) atomicOr(&ban[a.eos >> 5], 1u << (a.eos & 31));
  }
  if (a.ngram > 0 && a.step + 2 - a.ngram >= 0) {
    // generated so far: tokens[r][0 .. step]; the last n-1 of them are the key, every earlier occurrence bans its successor
    const long long* tk = a.tokens + (size_t)r * a.ldtok;
    const int n = a.ngram, key0 = a.step + 2 - n;
    for (int i = t; i + n - 1 <= a.step; i += kT) {
      bool same = true;
      for (int e = 0; e < n - 1; ++e) same = same && tk[i + e] == tk[key0 + e];
      if (same) { const int tok = (int)tk[i + n - 1]; if (tok >= 0 && tok < V) atomicOr(&ban[tok >> 5], 1u << (tok & 31)); }
    }
  }
  __syncthreads();
  // ---- pass A: log-sum-exp over the softmax domain (one pass: running max / rescaled sum per thread) ---------------------
  auto in_domain = [&](int v) { return !pre_range || v < 4 || (v >= a.range_lo && v < a.range_hi); };
  // visits every in-domain vocabulary entry of the row as fn(v, logit / temperature); 16-byte loads over the aligned bulk
  const bool vec_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  constexpr int VE = 16 / (int)sizeof(T);
  auto for_each = [&](auto&& fn) {
    if (pre_list) {
      for (int e = t; e < nal; e += kT) { const int v = eos_only ? a.eos : al[e]; fn(v, ldf(x, v) * inv_t); }
      return;
    }
    const int nv = vec_ok ? V / VE : 0;
    for (int i = t; i < nv; i += kT) {
      const uint4 u = reinterpret_cast<const uint4*>(x)[i];
      float f[VE];
      if constexpr (sizeof(T) == 2) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 p2 = __bfloat1622float2(h2[e]); f[2 * e] = p2.x; f[2 * e + 1] = p2.y; }
      } else {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
      }
#pragma unroll
      for (int e = 0; e < VE; ++e) if (in_domain(i * VE + e)) fn(i * VE + e, f[e] * inv_t);
    }
    for (int v = nv * VE + t; v < V; v += kT) if (in_domain(v)) fn(v, ldf(x, v) * inv_t);
  };
  // whole-vocabulary rows on 16-byte groups: fn(first index, values / temperature, count); entries outside a constraint range
  // arrive as -inf (they are outside the softmax domain and never candidates), so a group costs no per-element branch
  const bool grouped = vec_ok && !pre_list;
  auto for_groups = [&](auto&& fn) {
    const int nv = V / VE;
    for (int i = t; i < nv; i += kT) {
      const uint4 u = reinterpret_cast<const uint4*>(x)[i];
      float f[VE];
      if constexpr (sizeof(T) == 2) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 p2 = __bfloat1622float2(h2[e]); f[2 * e] = p2.x; f[2 * e + 1] = p2.y; }
      } else {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
      }
#pragma unroll
      for (int e = 0; e < VE; ++e) f[e] *= inv_t;
      if (pre_range) {
#pragma unroll
        for (int e = 0; e < VE; ++e) if (!in_domain(i * VE + e)) f[e] = -CUDART_INF_F;
      }
      fn(i * VE, f, VE);
    }
    for (int v = nv * VE + t; v < V; v += kT) {
      float f[VE];
      f[0] = in_domain(v) ? ldf(x, v) * inv_t : -CUDART_INF_F;
      fn(v, f, 1);
    }
  };
  float mx = -CUDART_INF_F, sum = 0.f;
  if (grouped) {
    for_groups([&](int, const float (&f)[VE], int n) {
      float m = f[0];
#pragma unroll
      for (int e = 1; e < VE; ++e) if (e < n) m = fmaxf(m, f[e]);      // (fmaxf drops NaNs: they poison the sum below)
      if (m > mx) { sum *= __expf(mx - m); mx = m; }                    // (exp(-inf) = 0 on the first finite group)
      if (mx > -CUDART_INF_F) {
#pragma unroll
        for (int e = 0; e < VE; ++e) if (e < n) sum += __expf(f[e] - mx);
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) if (e < n && f[e] != f[e]) sum = f[e];
      }
    });
  } else {
    for_each([&](int, float s) {
      if (s > mx) { sum = sum * __expf(mx - s) + 1.f; mx = s; }          // (exp(-inf) = 0 on the first entry)
      else if (s > -CUDART_INF_F) sum += __expf(s - mx);
      else if (s != s) sum = s;                                           // a NaN logit poisons the row like F.log_softmax: all -inf below
    });
  }
  const float tmx = mx;            // this thread's largest logit / temperature (threshold of the fast candidate filter below)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
    const float nm = fmaxf(mx, om);
    const float mu = nm == -CUDART_INF_F ? 0.f : nm;
    sum = sum * __expf(mx - mu) + os * __expf(om - mu);
    mx = nm;
  }
  if (lane == 0) { red[warp] = mx; reds[warp] = sum; }
  __syncthreads();
  if (t == 0) {
    float m = red[0];
    for (int w = 1; w < kT / 32; ++w) m = fmaxf(m, red[w]);
    const float mu = m == -CUDART_INF_F ? 0.f : m;
    float ssum = 0.f;
    for (int w = 0; w < kT / 32; ++w) ssum += reds[w] * __expf(red[w] - mu);
    bcast = mu + logf(ssum);                       // all domain entries -inf / NaN: lse = -inf -> every log-prob NaN -> -inf below
  }
  __syncthreads();
  const float lse = bcast;
  const float prev = a.prev_scores ? a.prev_scores[r] : 0.f;
  auto final_val = [&](int v, float s) {
    float lp = s - lse;
    if (has_prefix && v != (int)ptok) lp = pfill;
    if (lp != lp) lp = -CUDART_INF_F;                                   // sequence_generator.py:386
    if ((ban[v >> 5] >> (v & 31)) & 1u) lp = -CUDART_INF_F;
    if (v == a.unk) lp -= a.unk_penalty;
    if (a.force_eos) lp = v == a.eos ? (a.eos_one ? 1.f : lp) : -CUDART_INF_F;   // :399-404
    return lp + prev;
  };
  // ---- pass B, whole-vocabulary rows: threshold filter ------------------------------------------------------------------
  // B1: every thread's best candidate value; the K-th largest of the 256 thread maxima, tau, is a lower bound of the row's
  // K-th best (K different entries reach it).  B2: the few entries >= tau go to a list in shared memory, warp 0 sorts out the
  // K best of the list.  (Keeping K sorted entries per thread, the general path below, costs a K-deep insertion chain per
  // element for the whole warp: 260 us per launch at 320 rows x 59457 against 25 us for this path.)
  bool general = a.force_eos || pre_list || flip;
  // Fast filter (no post-softmax list / range, no forced prefix): log-prob + beam score is monotonic in the logit except for the
  // nb banned entries (pad, eos below min-len, n-gram bans) and unk (penalty), so the (K + nb)-th largest of the thread maxima
  // of PASS A bounds the row's K-th best candidate -- no second max pass, and the filter pass compares raw logits per 16-byte
  // group (3 instructions per element instead of the full candidate value).
  bool fast = !general && grouped && !has_prefix && !(nal >= 0 && a.trie_post) && !(a.range_lo >= 0 && a.range_post);
  if (fast) {
    int nb = 0;
    for (int w = t; w < nw; w += kT) nb += __popc(ban[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
    if (lane == 0) redi[warp] = nb;
    tmax[t] = tmx;
    if (t == 0) cnt = 0;
    __syncthreads();
    nb = a.unk_penalty != 0.f ? 1 : 0;
    for (int w = 0; w < kT / 32; ++w) nb += redi[w];
    const int rank = K + nb;
    if (rank > 64) {
      fast = false;                        // (many banned entries: the exact two-pass filter below)
    } else {
      if (warp == 0) {
        float tau = -CUDART_INF_F;
        for (int k = 0; k < rank; ++k) {
          float bv = -CUDART_INF_F; int be = lane;
          for (int e = lane; e < kT; e += 32) if (tmax[e] > bv) { bv = tmax[e]; be = e; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oe = __shfl_xor_sync(0xffffffffu, be, o);
            if (ov > bv || (ov == bv && oe < be)) { bv = ov; be = oe; }
          }
          tau = bv;
          if ((be & 31) == lane) tmax[be] = -CUDART_INF_F;
          __syncwarp();
        }
        if (lane == 0) bcast = tau;
      }
      __syncthreads();
      const float tau = bcast;
      const int unk_group = a.unk_penalty != 0.f ? a.unk / VE * VE : -1;
      for_groups([&](int v0, const float (&f)[VE], int n) {
        float m = f[0];
#pragma unroll
        for (int e = 1; e < VE; ++e) if (e < n) m = fmaxf(m, f[e]);
        if (!(m >= tau) && v0 != unk_group) return;
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if (e < n && (f[e] >= tau || v0 + e == a.unk)) {
            const float fv = final_val(v0 + e, f[e]);
            if (fv > -CUDART_INF_F) {
              const int pos = atomicAdd(&cnt, 1);
              if (pos < CAP) { lv[pos] = fv; li[pos] = v0 + e; }
            }
          }
        }
      });
      __syncthreads();
    }
  }
  if (!general && !fast) {
    float tm = -CUDART_INF_F;
    for_each([&](int v, float s) { tm = fmaxf(tm, final_val(v, s)); });
    tmax[t] = tm;
    if (t == 0) cnt = 0;
    __syncthreads();
    if (warp == 0) {
      float tau = -CUDART_INF_F;
      for (int k = 0; k < K; ++k) {
        float bv = -CUDART_INF_F; int be = lane;
        for (int e = lane; e < kT; e += 32) if (tmax[e] > bv) { bv = tmax[e]; be = e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oe = __shfl_xor_sync(0xffffffffu, be, o);
          if (ov > bv || (ov == bv && oe < be)) { bv = ov; be = oe; }
        }
        tau = bv;
        if ((be & 31) == lane) tmax[be] = -CUDART_INF_F;
        __syncwarp();
      }
      if (lane == 0) bcast = tau;
    }
    __syncthreads();
    const float tau = bcast;
    for_each([&](int v, float s) {
      const float f = final_val(v, s);
      if (f >= tau && f > -CUDART_INF_F) {
        const int pos = atomicAdd(&cnt, 1);
        if (pos < CAP) { lv[pos] = f; li[pos] = v; }
      }
    });
    __syncthreads();
  }
  if (!general) {
    const int n = cnt;
    if (n > CAP) {
      general = true;                      // (a row of equal logits, say): the exact general path
    } else if (warp == 0) {
      for (int k = 0; k < K; ++k) {
        float bv = -CUDART_INF_F; int bi = 0x7fffffff, be = lane;
        for (int e = lane; e < n; e += 32) if (better(lv[e], li[e], bv, bi)) { bv = lv[e]; bi = li[e]; be = e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          const int oe = __shfl_xor_sync(0xffffffffu, be, o);
          if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; be = oe; }
        }
        if (lane == 0) { out_v[k] = bv; out_i[k] = bi; }
        if (bi != 0x7fffffff && (be & 31) == lane) { lv[be] = -CUDART_INF_F; li[be] = 0x7fffffff; }
        __syncwarp();
      }
    }
  }
  if (!general) return;
  // ---- pass B, general path: per-thread K best ------------------------------------------------------------------------------
  float val[K]; int idx[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { val[k] = -CUDART_INF_F; idx[k] = 0x7fffffff; }
  auto consider = [&](int v, float s) { insert<K>(val, idx, final_val(v, s), flip ? V - 1 - v : v); };
  if (a.force_eos && !pre_list) {
    if (t == 0 && in_domain(a.eos)) consider(a.eos, ldf(x, a.eos) * inv_t);
  } else {
    for_each(consider);
  }
  // ---- K rounds of block arg-max -------------------------------------------------------------------------------------------
  // warp level first: each warp reduces to its K best (K rounds of shuffles), then warp 0 merges the 8 lists
  int head = 0;        // next unconsumed entry of this thread's sorted list
  for (int k = 0; k < K; ++k) {
    float bv = head < K ? val[0] : -CUDART_INF_F;
    int bi = head < K ? idx[0] : 0x7fffffff;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; }
    }
    if (lane == 0) { cv[warp][k] = bv; ci[warp][k] = bi; }
    if (lane == bl && head < K) {      // pop: shift the list (registers, static indexing)
#pragma unroll
      for (int q = 0; q < K - 1; ++q) { val[q] = val[q + 1]; idx[q] = idx[q + 1]; }
      val[K - 1] = -CUDART_INF_F; idx[K - 1] = 0x7fffffff;
      ++head;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // merge the per-warp sorted lists: lane w < 8 walks list w
    int pos = 0;
    for (int k = 0; k < K; ++k) {
      float bv = (lane < kT / 32 && pos < K) ? cv[lane][pos] : -CUDART_INF_F;
      int bi = (lane < kT / 32 && pos < K) ? ci[lane][pos] : 0x7fffffff;
      int bl = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; }
      }
      if (lane == 0) { out_v[k] = bv; out_i[k] = (flip && bi != 0x7fffffff) ? V - 1 - bi : bi; }
      if (lane == bl) ++pos;
    }
  }
}

template <int KT>      // KT: width of the per-row lists in the workspace; a.K candidates are emitted
__global__ void beam_merge_kernel(OfaBeamArgs a) {
  pdl_sync();
  const int s = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s * a.beam >= a.R) return;
  // lane b < beam walks the sorted list of beam row b
  int pos = 0;
  const float* rv = a.row_val + ((size_t)s * a.beam + (lane < a.beam ? lane : 0)) * KT;
  const int* ri = a.row_idx + ((size_t)s * a.beam + (lane < a.beam ? lane : 0)) * KT;
  for (int k = 0; k < a.K; ++k) {
    const bool have = lane < a.beam && pos < KT && ri[pos] != 0x7fffffff;
    float bv = have ? rv[pos] : -CUDART_INF_F;
    long long bi = have ? (long long)lane * a.V + ri[pos] : 0x7fffffffffffffffLL;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
    }
    if (lane == 0) {
      a.cand_scores[(size_t)s * a.K + k] = bv;
      // fewer than K candidates exist (tiny constraint sets): fillers with score -inf, as torch.topk returns
      a.cand_index[(size_t)s * a.K + k] = bi == 0x7fffffffffffffffLL ? (long long)(a.V - 1 - k) : bi;
    }
    if (lane == bl) ++pos;
  }
}

// node_out[r] = child of node_in[parent[r]] along token tok[r]; -1 when the prefix has left the trie
__global__ void trie_advance_kernel(const int* __restrict__ trie_ptr, const int* __restrict__ trie_tok, const int* __restrict__ trie_child,
                                    const int* __restrict__ node_in, const long long* __restrict__ parent,
                                    const long long* __restrict__ tok, long long tok_stride, int* __restrict__ node_out, int R) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int nd = node_in[parent ? parent[r] : r];
  int out = -1;
  if (nd >= 0) {
    const long long tk = tok[(size_t)r * tok_stride];
    for (int e = trie_ptr[nd]; e < trie_ptr[nd + 1]; ++e)
      if (trie_tok[e] == tk) { out = trie_child[e]; break; }
  }
  node_out[r] = out;
}

// All-candidate scoring (utils/eval_utils.py:203-209, tasks/mm_tasks/vqa_gen.py:296-304): for every candidate row the sum over
// its counted positions of log_softmax(logits restricted to the trie's next layer)[target].  The reference builds a dense bool
// mask [rows, T, V] on the host per chunk, masked_fills the logits, takes a full-vocabulary log-softmax and gathers; here a
// position only reads the logits of its node's children.  One CTA per row, positions in order (a fixed summation order).
//   node[p] >= 0: children of that trie node;  -1: the whole vocabulary;  -2: position does not count (prompt / padding).
template <typename T>
__global__ void __launch_bounds__(128) trie_score_kernel(const T* __restrict__ logits, long long ld, int V, const int* __restrict__ seg_off,
                                                         const int* __restrict__ node, const long long* __restrict__ target,
                                                         const int* __restrict__ trie_ptr, const int* __restrict__ trie_tok, int pad,
                                                         float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[4];
  __shared__ int found_s;
  const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float total = 0.f;
  for (int p = seg_off[r]; p < seg_off[r + 1]; ++p) {
    const int nd = node[p];
    const long long tgt = target[p];
    if (nd == -2 || tgt == pad) continue;
    const int* al = nd >= 0 ? trie_tok + trie_ptr[nd] : nullptr;
    const int n = nd >= 0 ? trie_ptr[nd + 1] - trie_ptr[nd] : V;
    if (n == 0) continue;                             // empty next layer: the mask row is all False (eval_utils.py:208)
    const T* x = logits + (size_t)p * ld;
    if (t == 0) found_s = 0;
    float mx = -CUDART_INF_F;
    bool found = false;
    for (int e = t; e < n; e += 128) {
      const int v = al ? al[e] : e;
      found = found || v == tgt;
      mx = fmaxf(mx, (float)x[v]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if (lane == 0) red[warp] = mx;
    if (found) found_s = 1;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const bool hit = found_s != 0;
    float sum = 0.f;
    for (int e = t; e < n; e += 128) sum += expf((float)x[al ? al[e] : e] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = (red[0] + red[1]) + (red[2] + red[3]);
    total += hit ? ((float)x[tgt] - mx) - logf(sum) : -CUDART_INF_F;
  }
  if (t == 0) out[r] = total;
}

template <typename T>
int launch_rows(const OfaBeamArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)((a.V + 31) / 32) * 4;
  if (a.K <= 4) return (int)ofa_launch_pdl(beam_row_kernel<T, 4>, dim3(a.R), kT, smem, st, a);
  if (a.K <= 8) return (int)ofa_launch_pdl(beam_row_kernel<T, 8>, dim3(a.R), kT, smem, st, a);
  if (a.K <= 10) return (int)ofa_launch_pdl(beam_row_kernel<T, 10>, dim3(a.R), kT, smem, st, a);
  return (int)ofa_launch_pdl(beam_row_kernel<T, 16>, dim3(a.R), kT, smem, st, a);
}

}  // namespace

extern "C" int ofa_beam_topk(const OfaBeamArgs* a, void* stream) {
  OFA_CHECK(a->R > 0 && a->beam > 0 && a->R % a->beam == 0 && a->V > 4 && a->K >= 1 && a->K <= KMAX && a->beam <= 32,
            "ofa_beam_topk: bad sizes R=%d beam=%d V=%d K=%d", a->R, a->beam, a->V, a->K);
  OFA_CHECK(a->logits && a->row_val && a->row_idx && a->cand_scores && a->cand_index, "ofa_beam_topk: null operand");
  OFA_CHECK(a->temperature > 0.f, "ofa_beam_topk: temperature must be > 0");
  OFA_CHECK(!a->node || (a->trie_ptr && a->trie_tok), "ofa_beam_topk: trie arrays missing");
  OFA_CHECK(a->ngram <= 0 || a->tokens, "ofa_beam_topk: n-gram blocking needs the token buffer");
  OFA_CHECK((size_t)((a->V + 31) / 32) * 4 <= 48 * 1024, "ofa_beam_topk: vocabulary too large for the shared-memory mask");
  cudaStream_t st = (cudaStream_t)stream;
  // the row kernel writes K' >= K entries per row (template sizes 4 / 8 / 10 / 16): the workspace rows are K' wide
  OfaBeamArgs b = *a;
  const int Kt = a->K <= 4 ? 4 : a->K <= 8 ? 8 : a->K <= 10 ? 10 : 16;
  b.K = Kt;
  int rc = a->dtype == OFA_BF16 ? launch_rows<__nv_bfloat16>(b, st) : launch_rows<float>(b, st);
  if (rc != 0) return ofa_set_error("ofa_beam_topk: launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  OFA_LAUNCH_CHECK("beam_row_kernel");
  const int bsz = a->R / a->beam;
  b.K = a->K;
  // (merge reads rows of width Kt, emits K)
  if (Kt == 4) OFA_CUDA(ofa_launch_pdl(beam_merge_kernel<4>, dim3((bsz + 3) / 4), 128, 0, st, b));
  else if (Kt == 8) OFA_CUDA(ofa_launch_pdl(beam_merge_kernel<8>, dim3((bsz + 3) / 4), 128, 0, st, b));
  else if (Kt == 10) OFA_CUDA(ofa_launch_pdl(beam_merge_kernel<10>, dim3((bsz + 3) / 4), 128, 0, st, b));
  else OFA_CUDA(ofa_launch_pdl(beam_merge_kernel<16>, dim3((bsz + 3) / 4), 128, 0, st, b));
  OFA_LAUNCH_CHECK("beam_merge_kernel");
  return 0;
}

extern "C" int ofa_beam_topk_width(int K) { return K <= 4 ? 4 : K <= 8 ? 8 : K <= 10 ? 10 : 16; }

extern "C" int ofa_trie_advance(const int* trie_ptr, const int* trie_tok, const int* trie_child, const int* node_in,
                                const long long* parent, const long long* tok, long long tok_stride, int* node_out, int R,
                                void* stream) {
  OFA_CHECK(R > 0 && trie_ptr && trie_tok && trie_child && node_in && tok && node_out, "ofa_trie_advance: null operand");
  OFA_CUDA(ofa_launch_pdl(trie_advance_kernel, dim3((R + 127) / 128), 128, 0, (cudaStream_t)stream, trie_ptr, trie_tok, trie_child,
                          node_in, parent, tok, tok_stride, node_out, R));
  OFA_LAUNCH_CHECK("trie_advance_kernel");
  return 0;
}

extern "C" int ofa_trie_score(const void* logits, long long ld, int dtype, int V, const int* seg_off, const int* node,
                              const long long* target, const int* trie_ptr, const int* trie_tok, int pad, float* out, int rows,
                              void* stream) {
  OFA_CHECK(rows > 0 && V > 0 && logits && seg_off && node && target && out, "ofa_trie_score: null operand or empty problem");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == OFA_BF16)
    OFA_CUDA(ofa_launch_pdl(trie_score_kernel<__nv_bfloat16>, dim3(rows), 128, 0, st, (const __nv_bfloat16*)logits, ld, V, seg_off, node, target, trie_ptr, trie_tok, pad, out));
  else if (dtype == OFA_F32)
    OFA_CUDA(ofa_launch_pdl(trie_score_kernel<float>, dim3(rows), 128, 0, st, (const float*)logits, ld, V, seg_off, node, target, trie_ptr, trie_tok, pad, out));
  else
    return ofa_set_error("ofa_trie_score: bad dtype %d", dtype);
  OFA_LAUNCH_CHECK("trie_score_kernel");
  return 0;
}

// ---- beam bookkeeping of one step (models/sequence_generator.py:438-586 when no hypothesis finishes) -------------------------
// From the top 2*beam candidates of every sentence: eos candidates among the first `beam` (not ignored, finite) are COUNTED
// (eos_n[s]; the host takes the reference's finalisation path for the whole step when any count is non-zero), and -- the
// common case, valid when all counts are zero -- the `beam` first non-eos / non-ignored candidates become the new hypotheses
// (:528-560: active_mask = eos_mask * cand_size + arange, topk(smallest)): their parents' token / score prefixes are gathered
// into the OTHER token / score buffers and the chosen token / cumulative score appended (:562-586).  Replaces ~20 index kernels
// (eq, ne, and, masked_select, add, topk, ge, gather x3, index_select x2, copies) by one launch; one CTA per sentence.
namespace {
__global__ void __launch_bounds__(32) beam_advance_kernel(const float* __restrict__ cand_scores, const long long* __restrict__ cand_index,
                                                          int C2, const unsigned char* __restrict__ ignore_in,
                                                          const long long* __restrict__ tok_in, long long ldtok,
                                                          const float* __restrict__ sc_in, long long ldsc,
                                                          long long* __restrict__ tok_out, float* __restrict__ sc_out,
                                                          unsigned char* __restrict__ ignore_out, long long* __restrict__ active_bbsz,
                                                          int* __restrict__ eos_n, int beam, int V, int eos, int step) {
  pdl_sync();
  __shared__ long long sel_row[16], sel_tok[16];
  __shared__ float sel_score[16];
  const int s = blockIdx.x, lane = threadIdx.x;
  if (lane == 0) {
    unsigned mask = 0;      // bit c: candidate c is an eos candidate or (c < beam) ignored
    int n = 0;
    for (int c = 0; c < C2; ++c) {
      const long long idx = cand_index[(size_t)s * C2 + c];
      const float sc = cand_scores[(size_t)s * C2 + c];
      bool e = (idx % V) == eos && sc != -CUDART_INF_F;                      // :444
      const bool ign = c < beam && ignore_in[(size_t)s * beam + c] != 0;
      if (ign) e = false;                                                      // :446
      if (c < beam && e) ++n;
      if (e || ign) mask |= 1u << c;                                           // :528
    }
    eos_n[s] = n;
    int k = 0;
    for (int pass = 0; pass < 2 && k < beam; ++pass)                           // smallest active_mask first: unmasked in candidate order,
      for (int c = 0; c < C2 && k < beam; ++c)                                 // then masked ones (:533-539)
        if (((mask >> c) & 1u) == (unsigned)pass) {
          const long long idx = cand_index[(size_t)s * C2 + c];
          sel_row[k] = idx / V + (long long)s * beam;                          // cand_bbsz_idx (:440)
          sel_tok[k] = idx % V;
          sel_score[k] = cand_scores[(size_t)s * C2 + c];
          ignore_out[(size_t)s * beam + k] = pass ? 1 : 0;                     // :544
          active_bbsz[(size_t)s * beam + k] = sel_row[k];
          ++k;
        }
  }
  __syncwarp();
  for (int k = 0; k < beam; ++k) {
    const long long src = sel_row[k], dst = (long long)s * beam + k;
    for (int c = lane; c <= step; c += 32) tok_out[dst * ldtok + c] = tok_in[src * ldtok + c];          // :563-565
    for (int c = lane; c < step; c += 32) sc_out[dst * ldsc + c] = sc_in[src * ldsc + c];              // :571-574
    if (lane == 0) {
      tok_out[dst * ldtok + step + 1] = sel_tok[k];                                                     // :567-569
      sc_out[dst * ldsc + step] = sel_score[k];                                                         // :575-577
    }
  }
}
}  // namespace

extern "C" int ofa_beam_advance(const float* cand_scores, const long long* cand_index, int C2, const unsigned char* ignore_in,
                                const long long* tok_in, long long ldtok, const float* sc_in, long long ldsc, long long* tok_out,
                                float* sc_out, unsigned char* ignore_out, long long* active_bbsz, int* eos_n, int bsz, int beam,
                                int V, int eos, int step, void* stream) {
  OFA_CHECK(bsz > 0 && beam > 0 && beam <= 16 && C2 >= beam && C2 <= 32 && V > 0 && step >= 0, "ofa_beam_advance: bad sizes");
  OFA_CHECK(cand_scores && cand_index && ignore_in && tok_in && sc_in && tok_out && sc_out && ignore_out && active_bbsz && eos_n,
            "ofa_beam_advance: null operand");
  OFA_CHECK(tok_in != tok_out && sc_in != sc_out, "ofa_beam_advance: the gather needs separate output buffers");
  OFA_CUDA(ofa_launch_pdl(beam_advance_kernel, dim3(bsz), 32, 0, (cudaStream_t)stream, cand_scores, cand_index, C2, ignore_in, tok_in,
                          ldtok, sc_in, ldsc, tok_out, sc_out, ignore_out, active_bbsz, eos_n, beam, V, eos, step));
  OFA_LAUNCH_CHECK("beam_advance_kernel");
  return 0;
}
