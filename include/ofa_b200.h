/* C ABI of libofa_b200.so -- the sm_100a kernels behind Musketeer's OFA forward/backward hot path.
 *
 * Boundary: the reference is pure Python; its drop-in surface is the fairseq `ofa` model plugin
 * (models/ofa/ofa.py:25-171).  The reference has no FFI of its own, so this header is the binding one level below
 * that plugin: each entry point replaces the ATen/cuBLAS call sites named beside it.  Plain pointers and sizes only
 * (device pointers unless noted), an explicit CUDA stream, `int` status (0 = ok, message via ofa_last_error()).
 * No entry point allocates or frees device memory or synchronises the device; callers own every buffer.
 * dtype codes: 0 = float32, 1 = bfloat16.                                                                         */
#ifndef OFA_B200_H_
#define OFA_B200_H_
#ifdef __cplusplus
extern "C" {
#endif

const char* ofa_last_error(void);
int ofa_abi_version(void);
/* programmatic dependent launch of the library's kernels (prologue overlap with the predecessor's tail): 1 = on (default);
 * returns the previous setting */
int ofa_set_pdl(int enabled);

/* ---- GEMM: every nn.Linear on the path, forward / dgrad / wgrad --------------------------------------------------
 * D[b](m,n) = act((sum_k A[b](m,k) B[b](n,k) + bias[n]) * alpha) + resid[b](m,n);  A,B bf16, fp32 accumulate (tcgen05).
 * When D rows are padded to a multiple of 8 elements (ldd == ceil8(N)) the padding columns [N, ldd) may be overwritten.
 * a_mn_major=0: A stored row-major [M][K] (lda);  =1: stored [K][M].  Same for B with N.  out_dtype: D/bias/resid type;
 * out_dtype 2 = fp32 accumulate: D (fp32) += alpha * A.B^T by TMA reduce-add (weight-gradient accumulation across K
 * slices, micro-batches and tasks without a workspace; no residual / activation).  In this mode `bias`, if not NULL,
 * names an fp32 [M] buffer that receives alpha * sum_k A(m,k): the bias gradient of a weight-gradient GEMM (A = dY^T),
 * computed by one extra N = 16 tcgen05.mma per k-step against an all-ones tile.
 * replaces: unify_multihead_attention.py:213-232,399 (q/k/v/out proj), unify_transformer_layer.py:280-284,557-561
 * (fc1/fc2), unify_transformer.py:739 (image_proj), :906-911,1303-1316 (pos q/k), :1577-1583 (tied output proj)
 * and their autograd transposes.                                                                                    */
int ofa_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int batch, long long lda, long long ldb,
                  long long ldd, long long stride_a, long long stride_b, long long stride_d, int a_mn_major,
                  int b_mn_major, int out_dtype, const void* bias, float alpha, int act, const void* resid,
                  long long ldr, long long stride_r, void* workspace, long long workspace_bytes, int alpha_cols,
                  void* stream);   /* alpha_cols > 0: alpha scales output columns < alpha_cols only (fused q|k|v projection) */
/* host helper: bytes of fp32 split-K scratch the call above wants for this problem (0 = none; passing less is legal) */
long long ofa_gemm_workspace_bytes(int M, int N, int K, int batch);
/* kernel variant switch (A/B testing): 0 single-CTA tiles, 1 CTA pair with TMA-multicast B, 2 cta_group::2 MMA on
 * 256x256 pair tiles (default); returns the previous setting */
int ofa_gemm_set_pair_mode(int enabled);
/* bf16 epilogue variant switch (A/B testing): 1 staged through shared memory + TMA store (default), 0 per-thread row
 * stores; returns the previous setting */
int ofa_gemm_set_tma_store(int enabled);
/* weight-gradient tile-shape switch (A/B testing): 1 = 128 x 256 tiles for fp32-accumulate problems (default) */
int ofa_gemm_set_wgrad_bn256(int enabled);
/* small-M tile-shape switch (A/B testing): 1 = 128 x 64 tiles, unsplit up to K = 1984, for bf16 forward / dgrad problems with
 * fewer 128-wide tiles than half the SMs (default) */
int ofa_gemm_set_small64(int enabled);
/* smallest number of 256 x 256 pair tiles for which the cta_group::2 kernel is chosen (default 38; A/B testing) */
int ofa_gemm_set_pair_min_tiles(int n);

/* fp32 -> three bf16 terms laid out as six K-blocks (fp32 parity mode operands for ofa_gemm_bf16) */
int ofa_split3_bf16(const float* x, long long ldx, int rows, int C, void* out, long long ldo, long long blk_stride,
                    int pattern, void* stream);

/* ---- LayerNorm (fairseq.modules.LayerNorm call sites: unify_transformer_layer.py:259-283,466-560;
 * unify_transformer.py:731-747,898-904,951,1300,1486-1493,1567).  y = LN(f(x))*gamma+beta (+resid), f = id | gelu      */
int ofa_layernorm_fwd(const void* x, const void* gamma, const void* beta, const void* resid, void* y, float* mean,
                      float* rstd, int rows, int C, float eps, int gelu_in, int dtype, void* stream);
int ofa_layernorm_bwd_nparts(int rows); /* host helper: workspace = 2 * nparts * C floats */
/* A/B switch: wide rows (> 2048 columns) of the backward stream through a shared-memory ring filled by bulk copies
   (default 1); returns the previous setting */
int ofa_layernorm_set_staged(int enabled);
/* accumulate=1: dgamma/dbeta += (gradient accumulation across micro-batches).  dskip (optional, [rows, C]): the gradient
   that reached x through the residual branch that forks off before the LayerNorm (unify_transformer_layer.py:259-262:
   residual = x; x = self_attn_layer_norm(x)); it is added into dx in the same pass instead of by a separate kernel. */
int ofa_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, void* dx,
                      void* dgamma, void* dbeta, float* workspace, int rows, int C, int gelu_in, int accumulate,
                      const void* dskip, int dtype, void* stream);

/* ---- glue -------------------------------------------------------------------------------------------------------- */
int ofa_colsum(const void* x, long long ld, int rows, int C, void* out, float* workspace /* 64*C floats */, float alpha,
               int accumulate, int dtype, void* stream); /* bias gradients: out = (accumulate ? out : 0) + alpha * colsum(x) */
int ofa_embed_gather(const long long* idx, const void* table, const void* addvec, void* out, long long ldo, int rows,
                     int C, int dtype, void* stream); /* unify_transformer.py:725-730,885,1450,1475 */
int ofa_embed_scatter_add(const long long* idx, const void* dout, long long ldo, void* dtable, int rows, int C,
                          long long skip_idx, int dtype, void* stream);
int ofa_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream);
int ofa_mask_rows(void* x, const unsigned char* rowmask, int rows, int C, int dtype, void* stream); /* :892-893 */
int ofa_gelu(const void* x, const void* dy, void* out, long long n, int backward, int dtype, void* stream);
/* y = resid + x * bernoulli_mask(seed) / (1-p) * row_scale[sample]: dropout + drop-path + residual in one pass
 * (FairseqDropout + unify_transformer_layer.py:19-35,197-198); the backward is the same call on dy with resid = NULL */
int ofa_dropout_residual(const void* x, const void* resid, void* y, long long n, int C, int rows_per_sample, float p,
                         const float* row_scale, const unsigned long long* seed, int dtype, void* stream);

/* ---- BatchNorm2d (+ReLU, +residual) on NHWC activations viewed as [R = N*H*W, C]; training-mode batch statistics with
 * running-stat update (momentum, unbiased variance), or eval / frozen mode (models/ofa/resnet.py:113-133,211-220,
 * frozen_bn.py:36-57).  stats = 4*C floats (mean | rstd | scale | shift); mean and rstd feed the backward.
 * workspace: ofa_batchnorm_workspace_floats() floats that are ZERO ON ENTRY; the kernels leave them zero again, so one
 * zero-initialised scratch per stream serves every call (no memset per layer).
 * groups > 1: the R rows are `groups` equal, consecutive groups (the images of several tasks batched through the stem);
 * statistics / normalisation / backward are per group (stats = groups * 4*C floats), the running statistics are updated
 * group after group and dgamma / dbeta are summed over the groups -- the arithmetic of `groups` separate calls.          */
long long ofa_batchnorm_workspace_floats(int C);
/* tuning switches: resident-CTA waves per launch (0 = automatic, the default) and rows in flight per thread of the
   backward kernels (2, 4 or 6; default 4); values outside the accepted set leave the setting unchanged */
int ofa_batchnorm_set_tuning(int waves, int bwd_unroll);
int ofa_batchnorm_fwd(const void* x, const void* res, void* y, const void* gamma, const void* beta, void* running_mean,
                      void* running_var, long long R, int C, float eps, float momentum, int training, int relu,
                      float* stats, float* workspace, int groups, int dtype, void* stream);
/* y may be NULL when relu && no residual: the mask is then recomputed from x with the forward's scale / shift */
int ofa_batchnorm_bwd(const void* x, const void* dy, const void* y, const void* gamma, const float* stats, void* dx,
                      void* dres, void* dgamma, void* dbeta, int accumulate, long long R,
                      int C, int batch_stats, int relu, float* workspace, int groups, int dtype, void* stream);

/* ---- 3x3 / stride 1 / padding 1 convolution as an implicit GEMM (models/ofa/resnet.py:107-108,119-121: conv2 of the
 * stride-1 bottlenecks).  bf16 NHWC activations [N][H][W][C]; weight bytes [Cout][3][3][Cin] (a channels_last
 * [Cout,Cin,3,3] tensor); Cin, Cout multiples of 64.  dgrad = 0: in = x (Cin channels) -> out = y (Cout channels);
 * dgrad = 1: in = dy (Cout channels) -> out = dx (Cin channels).  The weight gradient is split along the pixel axis
 * into an fp32 workspace and reduced; accumulate = 1 adds into dw (gradient accumulation across micro-batches);
 * accumulate = 2: dw is an fp32 [Cout][3][3][Cin] buffer that every tile is reduce-added into (no workspace).       */
int ofa_conv3x3_bf16(const void* in, const void* weight, void* out, int NI, int H, int W, int Cin, int Cout, int dgrad,
                     void* stream);
long long ofa_conv3x3_wgrad_workspace_bytes(int NI, int H, int W, int Cin, int Cout);
int ofa_conv3x3_wgrad_bf16(const void* x, const void* dy, void* dw, int NI, int H, int W, int Cin, int Cout,
                           int accumulate, void* workspace, long long workspace_bytes, void* stream);

/* ---- incremental decoding for beam search (models/sequence_generator.py:209-598; decoder layers with
 * incremental_state, unify_transformer_layer.py:432-582; cache reorder unify_multihead_attention.py:458-480).
 * ofa_attn_decode: one query token per row; G consecutive rows (the beams of a sentence) share cache row kv_row[group], so
 * cross-attention K / V / pos_k live once per sentence.  scores = q.k + pos_q.pos_k (+ tok_lut[h][(q_pos - j) + tok_max - 1]),
 * key padding kpm, fp32 softmax, out = P V * head_scale.  ofa_cache_gather: dst[p][r][:L] = src[p][order[r]][:L] over all
 * (layer, k|v) planes p in one launch (beam reorder of the self-attention cache; only the L valid positions move).     */
typedef struct OfaDecodeArgs {
  const void* q; const void* pq; long long ldq, ldpq;   /* pq / pk NULL (together): no absolute-position term in this launch */
  const void* k; const void* v; const void* pk;
  long long ldk, bsk, ldv, bsv, ldpk, bspk;      /* element (row, j, h*64+d) at row*bs + j*ld + h*64 + d */
  const int* kv_row; const int* pk_row;          /* [ceil(R/G)] cache rows (NULL: group index / kv_row) */
  const unsigned char* kpm; long long kpm_stride;
  void* o; long long ldo; const float* head_scale; const float* tok_lut; int tok_max; int q_pos;
  int R, G, H, S;
  /* a score term shared by all layers (the cross-attention pos_q.pos_k^T: unify_transformer.py:1461-1466) is computed once per
   * step by a launch with score_out set (raw scores, no softmax, v / o unused) and added by the layers' launches as bias_in;
   * both [R][H][bias_ld] fp32 */
  const float* bias_in; float* score_out; long long bias_ld;
  /* paged K / V: key j of cache row r lives in page page_table[r*max_pages + j/page_len] at (j % page_len)*ld; page p starts
   * at p*page_stride elements from k / v (NULL: contiguous rows) */
  const int* page_table; int page_len, max_pages; long long page_stride;
} OfaDecodeArgs;
int ofa_attn_decode(const OfaDecodeArgs* args, int dtype, void* stream);
/* A/B switch: 0 = the (group, head) decode kernel also for short self-attention problems; returns the previous setting */
int ofa_attn_decode_set_short(int on);
/* A/B switch: 0 = the two-pass warp-MMA kernel (score buffer in shared memory) for bf16 long-key decode attention */
int ofa_attn_decode_set_online(int on);
int ofa_cache_gather(const void* src, void* dst, const long long* order, int rows, int L, int D, long long row_stride,
                     long long plane_stride, int planes, int dtype, void* stream);
/* paged self-attention cache of the incremental decoder: pool [slot][plane = 2*layer + (k|v)][page_len][D].  A beam reorder
 * (models/sequence_generator.py:337-350; the reference index_selects every K / V tensor, unify_multihead_attention.py:455-473)
 * copies TABLE ENTRIES of the full, immutable pages and copies only the `off` valid positions of the partial last page into the
 * row's own slot (r*max_pages + page)*2 + parity; ofa_page_write appends one layer's new K / V token of every row.           */
int ofa_page_reorder(void* pool, const int* table_src, int* table_dst, const long long* order, int rows, int max_pages, int page,
                     int off, int parity, int page_len, int D, int planes, int dtype, void* stream);
int ofa_page_write(void* pool, const int* table, const void* k, const void* v, long long ldk, long long ldv, int rows,
                   int max_pages, int page, int off, int page_len, int D, int planes, int plane_k, int dtype, void* stream);

/* ---- one beam-search step after the decoder (models/sequence_generator.py:352-437,852-889; models/search.py:119-144):
 * logits -> temperature -> constraint masks -> fp32 log-softmax -> min-len / max-len / pad / unk / n-gram / zero-shot masks ->
 * + previous beam scores -> the K = 2*beam best (score, beam*V + token) per sentence, sorted.  Constraint trie as CSR
 * (children of node n: trie_tok[trie_ptr[n] .. trie_ptr[n+1])) with the per-row node in `node` (-1: the prefix left the trie,
 * only eos is allowed, utils/trie.py:23-30); ofa_trie_advance moves the nodes along the chosen tokens after a reorder.
 * row_val / row_idx: workspace [R][ofa_beam_topk_width(K)].                                                                  */
typedef struct OfaBeamArgs {
  const void* logits; long long ld; int dtype;
  int R, beam, V, K;
  float temperature;
  const float* prev_scores;            /* [R] or NULL */
  int step0;                           /* 1: only the first beam of every sentence takes part */
  int eos, pad, unk; float unk_penalty;
  int block_eos, force_eos, eos_one;   /* step < min_len | step >= max_len | with force_eos: lprobs[eos] = 1 (ignore_eos) */
  int range_lo, range_hi, range_post;  /* constraint range: keep tokens < 4 and [lo, hi); lo < 0: none; post: after the softmax */
  const int* trie_ptr; const int* trie_tok; const int* node; int trie_post;
  const long long* tokens; long long ldtok; int step; int ngram;   /* n-gram blocking over tokens[r][0..step]; ngram 0: off */
  /* forced prefix (models/sequence_generator.py:600-613): prefix_tok[r] != pad keeps the log-prob of that token and sets every
   * other entry of the row to *prefix_fill (device scalar: min over the rows of the prefix log-probs - 1) or, prefix_fill NULL,
   * to -inf (the generator has a constraint trie).  NULL: none.  node[r] == -2: row not constrained by the trie (:867-868).   */
  const long long* prefix_tok; const float* prefix_fill;
  float* row_val; int* row_idx;
  float* cand_scores; long long* cand_index;   /* [R/beam][K] */
} OfaBeamArgs;
int ofa_beam_topk(const OfaBeamArgs* args, void* stream);
int ofa_beam_topk_width(int K);
int ofa_trie_advance(const int* trie_ptr, const int* trie_tok, const int* trie_child, const int* node_in, const long long* parent,
                     const long long* tok, long long tok_stride, int* node_out, int R, void* stream);

/* ---- input hand-off: decoded uint8 HWC pixels [N][H][W][3] -> normalised [N][3][H][W] activations (bf16 or fp32) with the
 * host transforms of data/mm_data/*_dataset.py (ToTensor: x / 255; Normalize: (x - mean) / std; IEEE fp32, same order), so the
 * loader ships 1 byte per value instead of 4 (trainer.py:1246-1284 moves the fp32 tensor, then casts).  mean3 / std3: HOST floats. */
int ofa_normalize_u8(const void* x, void* y, int N, int H, int W, const float* mean3, const float* std3, int dtype, void* stream);

/* ---- beam bookkeeping of one step without finalisation (models/sequence_generator.py:438-586): eos_n[s] = eos candidates among
 * the first `beam` of sentence s (the caller takes the reference's finalisation path when any is non-zero); otherwise the first
 * `beam` non-eos / non-ignored candidates become the new hypotheses: parents' token / score prefixes gathered into the OUTPUT
 * buffers, chosen token (column step + 1) and cumulative score (column step) appended, ignore flags and the reorder index
 * (active_bbsz, int64 [bsz*beam]) written.  cand_index = beam * V + token as produced by ofa_beam_topk; C2 = 2 * beam.      */
int ofa_beam_advance(const float* cand_scores, const long long* cand_index, int C2, const unsigned char* ignore_in,
                     const long long* tok_in, long long ldtok, const float* sc_in, long long ldsc, long long* tok_out,
                     float* sc_out, unsigned char* ignore_out, long long* active_bbsz, int* eos_n, int bsz, int beam, int V,
                     int eos, int step, void* stream);

/* ---- all-candidate scoring (utils/eval_utils.py:203-209; tasks/mm_tasks/vqa_gen.py:296-304, snli_ve.py:203-210): out[r] = sum
 * over the positions p in [seg_off[r], seg_off[r+1]) of log_softmax(logits[p] restricted to the next layer of trie node
 * node[p])[target[p]].  node[p] = -1: whole vocabulary, -2: position not counted; target == pad and empty layers count 0; a
 * target outside its layer gives -inf.  logits [positions][ld], fp32 arithmetic, fixed summation order.                  */
int ofa_trie_score(const void* logits, long long ld, int dtype, int V, const int* seg_off, const int* node,
                   const long long* target, const int* trie_ptr, const int* trie_tok, int pad, float* out, int rows,
                   void* stream);

/* ---- 3x3 / stride 2 / padding 1 max-pool of the stem on bf16 NHWC activations (models/ofa/resnet.py:179,216).  idx: one
 * byte per output element (window position of the first maximum); the backward gathers, no atomics.  C % 8 == 0.      */
int ofa_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* idx, int N, int H, int W, int C, void* stream);
int ofa_maxpool3x3s2_bwd(const void* dy, const unsigned char* idx, void* dx, int N, int H, int W, int C, void* stream);

/* ---- patch matrix of the 7x7 / stride 2 / padding 3 stem convolution (models/ofa/resnet.py:176,214): x bf16 NHWC with
 * C = 3 -> col bf16 [N*OH*OW][152], columns (c, kh, kw) as in weight.view(64, 147), 5 zero columns of row padding; the
 * convolution and its weight gradient are ofa_gemm_bf16 calls on col.                                                   */
int ofa_stem_patches(const void* x, void* col, int N, int H, int W, void* stream);

/* ---- patch matrix (im2col) of an NHWC activation and its adjoint: every convolution that the implicit-GEMM kernel above does
 * not cover (the two stride-2 3x3 convolutions of the stem, models/ofa/resnet.py:34-37,107-121, and all k > 1 convolutions of
 * the fp32 parity mode incl. the 7x7 stem convolution :176,214) is `patch matrix x weight^T` on ofa_gemm_bf16.
 * col[(n,oh,ow)][(kh*KW + kw)*C + c] = x[n][oh*stride - pad + kh][ow*stride - pad + kw][c] (0 outside), row stride ldcol;
 * ofa_col2im is the adjoint (dx = sum of the patch entries that read each pixel, fixed order, no atomics).           */
int ofa_im2col(const void* x, void* col, int N, int H, int W, int C, int KH, int KW, int stride, int pad, long long ldcol,
               int dtype, void* stream);
int ofa_col2im(const void* dcol, void* dx, int N, int H, int W, int C, int KH, int KW, int stride, int pad, long long ldcol,
               int dtype, void* stream);
/* nn.MaxPool2d(3, 2, 1) on NHWC activations of either dtype (the fp32 parity mode; bf16 with C % 8 == 0 uses the vectorised
 * kernels above).  backward = 0: in = x, out = y, idx written;  backward = 1: in = dy, out = dx, idx read.             */
int ofa_maxpool3x3s2_any(const void* in, unsigned char* idx, void* out, int N, int H, int W, int C, int backward, int dtype,
                         void* stream);

/* ---- stride-2 pixel subsampling of an NHWC activation = the input side of the stride-2 1x1 downsample convolutions
 * (models/ofa/resnet.py:196-203), and its adjoint (zero fill + scatter in one pass).  [N, H, W, C] are the dimensions of
 * the FULL-resolution tensor; the subsampled one is [N, (H+1)/2, (W+1)/2, C].  esize = bytes per element (2 | 4),
 * C * esize % 16 == 0.  backward = 0: dst = src[:, ::2, ::2, :];  backward = 1: dst (full) = adjoint of src (subsampled) */
int ofa_subsample2(const void* src, void* dst, int N, int H, int W, int C, int esize, int backward, void* stream);

/* ---- fused optimizer step (trainer.py:863-898 multiply_grads -> clip_grad_norm -> optimizer.step,
 * with the un-vendored fairseq Adam / FP16Optimizer arithmetic: fp32 master weights, decoupled weight decay
 * p -= wd*lr*p, bias-corrected step size, global-norm clipping with coefficient clip/(norm+1e-6)).
 * chunk_table: device array of n_chunks records {void* param; const void* grad; float* master; float* m; float* v;
 * long long n;} (48 bytes each, n <= 65536 elements per chunk); partial_sqnorm: n_chunks floats of scratch;
 * grad_norm_out: optional device float receiving the (scaled, unclipped) global gradient norm.  Two launches, no host
 * synchronisation; clip_norm <= 0 disables clipping; step >= 1 is the 1-based update count.                         */
int ofa_adam_step(const void* chunk_table, int n_chunks, float* partial_sqnorm, float* grad_norm_out, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, float clip_norm,
                  int param_dtype, void* stream);

/* ---- label-smoothed CE (+R-Drop KL): loss rows and d(logits) in one kernel, gradient written in place --------------
 * replaces criterions/label_smoothed_cross_entropy.py:81-126,228-260.                                                */
/* grad_row_scale (optional, [R]): the gradient rows are written multiplied by it -- the caller's promise of each row's
 * upstream gradient (e.g. 1 / sample_size of the row's task), so that the later ofa_scale_rows pass over the logits-sized
 * gradient degenerates to skipped rows (scale exactly 1); loss_rows / nll_rows stay unscaled.                          */
int ofa_ls_ce_fwd_bwd(void* logits, long long ld, const long long* target, const unsigned char* cmask,
                      const float* conf, int rows_per_sample, int R, int V, long long pad_idx, float eps, int cs,
                      int ce, int rdrop, float reg_alpha, float* loss_rows, float* nll_rows, float* kl_rows,
                      const float* grad_row_scale, int dtype, void* stream);
int ofa_scale_rows(void* x, long long ld, int R, int V, const float* scale, const unsigned char* row_keep,
                   int scale_per_row, int dtype, void* stream); /* scale: device scalar, or one factor per row; rows with factor 1 are skipped */

/* ---- attention with OFA position biases (unify_multihead_attention.py:345-398; bias assembly of
 * unify_transformer.py:640-658,906-933,1282-1318,1519-1529 computed in-kernel) ------------------------------------- */
typedef struct {
  const float* tok_lut;       /* [H][2*tok_max-1] text rel-pos bias by (i_t - j_t) + tok_max - 1, or NULL */
  int tok_max;                /* 1024 */
  int q_text_off, k_text_off; /* first text index on the query / key axis */
  const float* img_lut;       /* [H][n_img_rel] image rel-pos bias, or NULL */
  int n_img_rel;              /* (2*ibs-1)^2 + 3 */
  int ibs;                    /* image bucket size (42) */
  const int* q_pid;           /* [B][n_img_q] 1-based image position ids */
  const int* k_pid;           /* [B][n_img_k] */
  int n_img_q, n_img_k;       /* image tokens occupy [0, n_img) */
} OfaAttnBias;

typedef struct {
  const void *q, *pq, *k, *pk, *v; /* [B, L, H, 64] views; q and pq pre-scaled */
  void* o;                         /* [B, T, H, 64] */
  float* lse;                      /* [B, H, T] */
  long long ldq, ldpq, ldk, ldpk, ldv, ldo;       /* token strides (elements) */
  long long bsq, bspq, bsk, bspk, bsv, bso;       /* batch strides (elements) */
  int B, H, T, S;
  int causal;                  /* mask j > i + q_pos_off */
  int q_pos_off;               /* absolute position of query row 0 (incremental decoding) */
  const unsigned char* kpm;    /* [B][S] key padding mask (1 = pad) or NULL */
  const float* head_scale;     /* [H] c_attn or NULL */
  int p_round_bf16;            /* SIMT path only: round P to bf16 before P.V (mimics the bf16 reference path) */
  OfaAttnBias bias;
} OfaAttnArgs;

typedef struct {
  const void* dout;                 /* [B, T, H, 64], strides as o */
  void *dq, *dpq, *dk, *dpk, *dv;
  long long lddq, lddpq, lddk, lddpk, lddv;
  long long bsdq, bsdpq, bsdk, bsdpk, bsdv;
  float* dtok_lut;                  /* [H][2*tok_max-1] fp32, pre-zeroed, accumulated */
  float* dimg_lut;                  /* [H][n_img_rel]   fp32, pre-zeroed, accumulated */
  float* delta;                     /* [B,H,T] workspace; sum over rows / c_attn[h] = d c_attn[h] */
  float* P;                         /* SIMT path: [B,H,T,S] fp32 workspace */
  float* dS;                        /* SIMT path: [B,H,T,S] fp32 workspace */
  float dq_scale;                   /* tcgen05 path: dq (not dpq) is multiplied by this on its way out (the softmax scaling
                                       of a q that came unscaled-in-weights from a fused q|k|v projection); 0 means 1 */
  int acc_pos;                      /* tcgen05 / short-target paths: 1 = dpq and dpk are ADDED to the buffers' contents (the
                                       position projections pos_q / pos_k are shared by every layer of a stack,
                                       unify_transformer.py:906-912: the layers' backward passes sum into one buffer instead of
                                       autograd adding six tensors); 0 = overwritten */
} OfaAttnGrads;

int ofa_attn_fwd_simt(const OfaAttnArgs* args, int dtype, void* stream);
int ofa_attn_bwd_simt(const OfaAttnArgs* args, const OfaAttnGrads* grads, int dtype, void* stream);
/* bf16, TMA + tcgen05/TMEM flash attention */
int ofa_attn_fwd_tc(const OfaAttnArgs* args, void* stream);
/* A/B switch: 1 = warp-specialised forward (TMA producer warp, tcgen05 issuer warp, two softmax warpgroups; default),
 * 0 = the single-role kernel of round 1; returns the previous setting */
int ofa_attn_set_fwd_ws(int enabled);
/* A/B switch: 1 = attention backward with T <= 16 query rows and no relative-position bias (short-target cross-attention) runs on
 * the warp-level kernel of csrc/attention_small.cu (default), 0 = always the 128-row tcgen05 kernel; returns the previous setting */
int ofa_attn_set_bwd_small(int enabled);
/* backward: fills grads->delta, dq/dpq/dk/dpk/dv and accumulates dtok_lut / dimg_lut; dq_acc = B*T*H*128 floats scratch */
int ofa_attn_bwd_tc(const OfaAttnArgs* args, const OfaAttnGrads* grads, float* dq_acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif
