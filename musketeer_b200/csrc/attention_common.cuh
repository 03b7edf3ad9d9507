// Bias / mask helpers shared by the SIMT and tcgen05 attention kernels.
#pragma once
#include "../../include/ofa_b200.h"
#include "common.cuh"

typedef OfaAttnArgs AttnArgs;
typedef OfaAttnBias AttnBias;
typedef OfaAttnGrads AttnGrads;

// csrc/attention_small.cu: bf16 backward for T <= 16 query rows without relative-position bias (short-target cross-attention)
bool ofa_attn_bwd_small_applicable(const AttnArgs* a, const AttnGrads* g);
int ofa_attn_bwd_small_launch(const AttnArgs* a, const AttnGrads* g, cudaStream_t st);
bool ofa_attn_fwd_small_applicable(const AttnArgs* a);
int ofa_attn_fwd_small_launch(const AttnArgs* a, cudaStream_t st);

#ifdef __CUDACC__
// index into the per-head image LUT for position ids (1-based, row-major over an ibs x ibs grid):
// closed form of make_image_bucket_position (models/ofa/unify_transformer.py:66-81) for ids >= 1
__device__ __forceinline__ int img_bucket(int pid_q, int pid_k, int ibs) {
  const int qr = (pid_q - 1) / ibs, qc = (pid_q - 1) % ibs;
  const int kr = (pid_k - 1) / ibs, kc = (pid_k - 1) % ibs;
  return (qr - kr + ibs - 1) * (2 * ibs - 1) + (qc - kc + ibs - 1);
}
__device__ __forceinline__ float attn_bias_at(const AttnBias& bz, int b, int h, int i, int j) {
  float r = 0.f;
  if (bz.tok_lut && i >= bz.q_text_off && j >= bz.k_text_off)
    r += bz.tok_lut[(size_t)h * (2 * bz.tok_max - 1) + (i - bz.q_text_off) - (j - bz.k_text_off) + bz.tok_max - 1];
  if (bz.img_lut && i < bz.n_img_q && j < bz.n_img_k)
    r += bz.img_lut[(size_t)h * bz.n_img_rel +
                    img_bucket(bz.q_pid[(size_t)b * bz.n_img_q + i], bz.k_pid[(size_t)b * bz.n_img_k + j], bz.ibs)];
  return r;
}
__device__ __forceinline__ void attn_bias_grad_at(const AttnBias& bz, const AttnGrads& g, int b, int h, int i, int j,
                                                  float ds) {
  if (ds == 0.f) return;
  if (bz.tok_lut && g.dtok_lut && i >= bz.q_text_off && j >= bz.k_text_off)
    atomicAdd(g.dtok_lut + (size_t)h * (2 * bz.tok_max - 1) + (i - bz.q_text_off) - (j - bz.k_text_off) + bz.tok_max - 1, ds);
  if (bz.img_lut && g.dimg_lut && i < bz.n_img_q && j < bz.n_img_k)
    atomicAdd(g.dimg_lut + (size_t)h * bz.n_img_rel +
                  img_bucket(bz.q_pid[(size_t)b * bz.n_img_q + i], bz.k_pid[(size_t)b * bz.n_img_k + j], bz.ibs), ds);
}
#endif
