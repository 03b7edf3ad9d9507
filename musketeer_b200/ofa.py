"""B200-native OFA model: a drop-in for the reference's fairseq `ofa` model (models/ofa/ofa.py:25-171,
models/ofa/unify_transformer.py:126-1745) whose forward/backward runs on the kernels of libofa_b200.so.

Contract kept from the reference:
  * `OFAModel.build_model(args, task)`, `forward(src_tokens, src_lengths, prev_output_tokens, patch_images, ...)`
    -> `(logits [B,T,V], {"attn": [...], "inner_states": [...]})`, `.encoder(...)` -> dict of lists
    (`encoder_out [N,B,d]`, `encoder_padding_mask [B,N]`, `position_embeddings [B,N,d]`, ...), `.decoder(...)`,
    `reorder_encoder_out`, `get_normalized_probs`, `get_targets`, `max_positions`, `max_decoder_positions`,
    `enc_timer/dec_timer/cls_timer` attributes (never synchronised here).
  * state_dict keys / shapes / order identical to the reference (tests/golden/state_dict_spec.json), tied embedding.
nn.Linear / nn.LayerNorm / nn.Embedding objects below are parameter holders only -- their torch forward is never
called; arithmetic goes through `musketeer_b200.ops`.  Internally activations are batch-first [B, L, d].
"""
import logging
import math
import random
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._fairseq_compat import (FairseqEncoder, FairseqEncoderDecoderModel, FairseqIncrementalDecoder, HAVE_FAIRSEQ)
from .resnet import ResNetStem

logger = logging.getLogger(__name__)

DEFAULT_MAX_SOURCE_POSITIONS = 1024
DEFAULT_MAX_TARGET_POSITIONS = 1024


# ---------------------------------------------------------------------------------------------------------------------
# integer tables (bit-exact restatements of unify_transformer.py:53-81; verified in tests against the reference dump)
# ---------------------------------------------------------------------------------------------------------------------
def make_token_bucket_position(bucket_size, max_position=DEFAULT_MAX_SOURCE_POSITIONS):
    ctx = torch.arange(max_position, dtype=torch.long)[:, None]
    mem = torch.arange(max_position, dtype=torch.long)[None, :]
    rel = ctx - mem
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), mid - 1, torch.abs(rel))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    log_pos = log_pos.int()
    bucket = torch.where(abs_pos.le(mid), rel, log_pos * sign).long()
    return bucket + bucket_size - 1


def make_image_bucket_position(bucket_size, num_relative_distance):
    n = bucket_size * bucket_size
    r = torch.arange(n) // bucket_size
    c = torch.arange(n) % bucket_size
    t = torch.zeros(n + 1, n + 1, dtype=torch.long)
    t[1:, 1:] = (r[:, None] - r[None, :] + bucket_size - 1) * (2 * bucket_size - 1) + \
                (c[:, None] - c[None, :] + bucket_size - 1)
    t[0, :] = num_relative_distance - 3
    t[:, 0] = num_relative_distance - 2
    t[0, 0] = num_relative_distance - 1
    return t


def Embedding(num_embeddings, embedding_dim, padding_idx=None, zero_init=False):
    m = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    nn.init.normal_(m.weight, mean=0, std=embedding_dim ** -0.5)
    if padding_idx is not None:
        nn.init.constant_(m.weight[padding_idx], 0)
    if zero_init:
        nn.init.constant_(m.weight, 0)
    return m


def Linear(in_features, out_features, bias=True):
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.0)
    return m


def init_bert_params(module):
    """fairseq init_bert_params (un-vendored): N(0, 0.02) Linear/Embedding weights, zero biases / pad row (ofa.py:33)."""
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()


def _lin(mod, x, alpha=1.0, resid=None):
    return ops.linear(x, mod.weight, mod.bias, alpha, resid)


def _ln(mod, x, resid=None, gelu_in=False, fork=False):
    """fork=True: returns (LN(x), x); taking the residual connection from the second value adds its gradient inside the
    LayerNorm backward kernel (ops._LayerNorm) instead of in a separate pass."""
    return ops.layer_norm(x, mod.weight, mod.bias, resid, gelu_in, mod.eps, fork)


def _unsupported(args, names):
    for n in names:
        if getattr(args, n, False):
            raise NotImplementedError("musketeer_b200 OFA: option --%s is outside the hot-path scope (SURVEY.md 8)" %
                                      n.replace("_", "-"))


def _upgrade_layer_norm_names(module, state_dict, name, layer_norm_map):
    """Old checkpoints: `...layer_norms.{i}.*` -> the named LayerNorms; keys the checkpoint lacks are filled from the module
    (unify_transformer_layer.py:200-221,587-615)."""
    own = module.state_dict()
    for old, new in layer_norm_map.items():
        for m in ("weight", "bias"):
            k = "{}.layer_norms.{}.{}".format(name, old, m)
            if k in state_dict:
                state_dict["{}.{}.{}".format(name, new, m)] = state_dict.pop(k)
    prefix = name + "." if name != "" else ""
    for k, v in own.items():
        if prefix + k not in state_dict:
            state_dict[prefix + k] = v


def _grow_image_positions(module, state_dict, key):
    """A checkpoint trained with a smaller image bucket grid: append freshly initialised position rows
    (unify_transformer.py:1060-1071,1647-1658)."""
    have, want = state_dict[key], module.embed_image_positions.weight
    if len(have) < len(want):
        extra = torch.zeros(len(want) - len(have), have.size(1))
        nn.init.normal_(extra, mean=0, std=have.size(1) ** -0.5)
        state_dict[key] = torch.cat([have, extra.to(dtype=have.dtype, device=have.device)])


class MultiheadAttention(nn.Module):
    """Parameter holder + call into ops.attention (unify_multihead_attention.py:20-409)."""

    def upgrade_state_dict_named(self, state_dict, name):
        """fused `in_proj_weight / in_proj_bias` of very old checkpoints -> q / k / v projections (:495-524)."""
        prefix = name + "." if name != "" else ""
        for kind in ("weight", "bias"):
            k = prefix + "in_proj_" + kind
            if k in state_dict:
                w = state_dict.pop(k)
                dim = w.shape[0] // 3
                for i, proj in enumerate(("q_proj", "k_proj", "v_proj")):
                    state_dict[prefix + proj + "." + kind] = w[i * dim:(i + 1) * dim]

    def __init__(self, embed_dim, num_heads, scale_factor=2.0, scale_heads=False, self_attention=False,
                 encoder_decoder_attention=False):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        if self.head_dim != 64:
            raise NotImplementedError("attention kernels are built for head_dim 64 (all OFA archs but ofa_huge)")
        self.scaling = float(self.head_dim * scale_factor) ** -0.5
        self.self_attention, self.encoder_decoder_attention = self_attention, encoder_decoder_attention
        self.c_attn = nn.Parameter(torch.ones((num_heads,)), requires_grad=True) if scale_heads else None
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)

    def project_kv(self, key, fused=False):
        if fused:       # one GEMM for k|v (training / teacher-forced cross-attention)
            return ops.kv_linear(key, self.k_proj, self.v_proj)
        return _lin(self.k_proj, key), _lin(self.v_proj, key)

    def forward_self(self, h, pq, pk, tok_lut, img_lut, cfg):
        """Self-attention with the fused q|k|v projection (training / teacher-forced path)."""
        q, k, v = ops.qkv_linear(h, self.q_proj, self.k_proj, self.v_proj, self.scaling)
        cfg = dict(cfg)
        # q leaves the GEMM as s * (x Wq^T + bq); the gradient handed back to the fused projection is d/d(x Wq^T + bq) = s * dq
        cfg["H"], cfg["fused_qkv"], cfg["dq_scale"] = self.num_heads, True, self.scaling
        return ops.attention(q, pq, k, pk, v, tok_lut, img_lut, self.c_attn, cfg)

    def qkv_eval(self, h):
        """Inference: q * scaling | k | v of one GEMM over the concatenated projection weights, which are built once and cached
        (keyed on the parameters' version counters) instead of being concatenated at every decoder step."""
        ps = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight, self.q_proj.bias, self.k_proj.bias, self.v_proj.bias)
        key = tuple(p._version for p in ps) + (ps[0].data_ptr(), ps[0].dtype)
        c = getattr(self, "_qkv_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                c = self._qkv_cache = (key, torch.cat(ps[:3], 0), torch.cat(ps[3:], 0))
        shp = h.shape
        D = self.embed_dim
        y = ops.gemm(h.reshape(-1, shp[-1]), c[1], h.numel() // shp[-1], 3 * D, shp[-1], bias=c[2], alpha=self.scaling, alpha_cols=D)
        y = y.view(*shp[:-1], 3 * D)
        return y[..., :D], y[..., D:2 * D], y[..., 2 * D:]

    def forward(self, query, k, v, pq, pk, tok_lut, img_lut, cfg, q_pre=None):
        q = q_pre if q_pre is not None else _lin(self.q_proj, query, alpha=self.scaling)
        dec = cfg.get("decode")
        if dec is not None:        # incremental decoding: one token per row against the KV cache (csrc/decode.cu)
            bias = dec.get("bias")         # the position term of this attention, computed once per step for all layers
            return ops.attention_decode(q, None if bias is not None else pq, k, None if bias is not None else pk, v, dec["S"],
                                        self.num_heads, dec.get("G", 1), dec.get("kv_row"), dec.get("pk_row"), dec.get("kpm"),
                                        self.c_attn, tok_lut, dec.get("q_pos", 0), bias_in=bias, page=dec.get("page"))
        cfg = dict(cfg)
        cfg["H"] = self.num_heads
        return ops.attention(q, pq, k, pk, v, tok_lut, img_lut, self.c_attn, cfg)


class _FFNMixin:
    def _stochastic(self):
        return self.training and (self.dropout_p > 0.0 or self.drop_path_rate > 0.0)

    def _post_attn(self, attn, o, ln, x):
        """residual + drop_path(dropout(ln(out_proj(o))))   (unify_transformer_layer.py:270-273,513-516,546-549)"""
        if self._stochastic():
            y = _lin(attn.out_proj, o)
            if ln is not None:
                y = _ln(ln, y)
            return ops.dropout_residual(y, x, self.dropout_p, self.drop_path_rate, True)
        if ln is not None:
            return _ln(ln, _lin(attn.out_proj, o), resid=x)      # residual fused into the LN epilogue
        return _lin(attn.out_proj, o, resid=x)                   # residual fused into the GEMM epilogue

    def _ffn(self, x):
        h, r = _ln(self.final_layer_norm, x, fork=True)
        u = _lin(self.fc1, h)
        if self.ffn_layernorm is not None:
            g = _ln(self.ffn_layernorm, u, gelu_in=True)   # GELU fused into the LN prologue
        else:
            g = ops.gelu(u)
        if self._stochastic():
            return ops.dropout_residual(_lin(self.fc2, g), r, self.dropout_p, self.drop_path_rate, True)
        return _lin(self.fc2, g, resid=r)                  # residual fused into the GEMM epilogue


class TransformerEncoderLayer(nn.Module, _FFNMixin):
    """unify_transformer_layer.py:110-293 (pre-LN; dropout / drop-path probabilities must be 0 in this round)."""

    def __init__(self, args, drop_path_rate=0.0):
        super().__init__()
        d = args.encoder_embed_dim
        self.self_attn = MultiheadAttention(d, args.encoder_attention_heads, args.attn_scale_factor,
                                            getattr(args, "scale_heads", False), self_attention=True)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, args.encoder_ffn_embed_dim)
        self.fc2 = nn.Linear(args.encoder_ffn_embed_dim, d)
        self.attn_ln = nn.LayerNorm(d) if getattr(args, "scale_attn", False) else None
        self.ffn_layernorm = nn.LayerNorm(args.encoder_ffn_embed_dim) if getattr(args, "scale_fc", False) else None
        self.final_layer_norm = nn.LayerNorm(d)
        self.drop_path_rate = drop_path_rate
        self.dropout_p = float(args.dropout)

    def forward(self, x, pq, pk, tok_lut, img_lut, cfg):
        h, x = _ln(self.self_attn_layer_norm, x, fork=True)
        o = self.self_attn.forward_self(h, pq, pk, tok_lut, img_lut, cfg)
        x = self._post_attn(self.self_attn, o, self.attn_ln, x)
        return self._ffn(x)

    def upgrade_state_dict_named(self, state_dict, name):
        self.self_attn.upgrade_state_dict_named(state_dict, name + ".self_attn")
        _upgrade_layer_norm_names(self, state_dict, name, {"0": "self_attn_layer_norm", "1": "final_layer_norm"})


class TransformerDecoderLayer(nn.Module, _FFNMixin):
    """unify_transformer_layer.py:296-582."""

    def __init__(self, args, drop_path_rate=0.0):
        super().__init__()
        d = args.decoder_embed_dim
        H = args.decoder_attention_heads
        sh = getattr(args, "scale_heads", False)
        self.self_attn = MultiheadAttention(d, H, args.attn_scale_factor, sh, self_attention=True)
        sa = getattr(args, "scale_attn", False)
        self.self_attn_ln = nn.LayerNorm(d) if sa else None
        self.cross_attn_ln = nn.LayerNorm(d) if sa else None
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = MultiheadAttention(d, H, args.attn_scale_factor, sh, encoder_decoder_attention=True)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.ffn_layernorm = nn.LayerNorm(args.decoder_ffn_embed_dim) if getattr(args, "scale_fc", False) else None
        self.fc1 = nn.Linear(d, args.decoder_ffn_embed_dim)
        self.fc2 = nn.Linear(args.decoder_ffn_embed_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)
        self.drop_path_rate = drop_path_rate
        self.dropout_p = float(args.dropout)

    def forward(self, x, self_kv, cross_kv, spq, spk, cpq, cpk, tok_lut, self_cfg, cross_cfg):
        h, x = _ln(self.self_attn_layer_norm, x, fork=True)
        if "decode" in self_cfg:                 # incremental decoding: K / V go through the (paged) cache
            q, k, v = self_kv(self.self_attn, h)
            o = self.self_attn(h, k, v, spq, spk, tok_lut, None, self_cfg, q_pre=q)
        else:
            o = self.self_attn.forward_self(h, spq, spk, tok_lut, None, self_cfg)
        x = self._post_attn(self.self_attn, o, self.self_attn_ln, x)
        h, x = _ln(self.encoder_attn_layer_norm, x, fork=True)
        k, v = cross_kv(self.encoder_attn)
        G = cross_cfg.get("rows_per_sentence", 1)
        if G > 1:
            # G consecutive decoder rows (the candidates / hypotheses of one sentence) attend to ONE encoder row: cross-attention
            # is independent per query position, so the rows of a sentence are folded into the query-length axis (a view)
            # instead of replicating the encoder output G times (utils/eval_utils.py:192-201 repeat_interleaves it)
            B, T, d = h.shape
            o = self.encoder_attn(h.reshape(B // G, G * T, d), k, v, cpq.reshape(B // G, G * T, d), cpk, None, None, cross_cfg)
            o = o.reshape(B, T, d)
        else:
            o = self.encoder_attn(h, k, v, cpq, cpk, None, None, cross_cfg)
        x = self._post_attn(self.encoder_attn, o, self.cross_attn_ln, x)
        return self._ffn(x)

    def upgrade_state_dict_named(self, state_dict, name):
        self.self_attn.upgrade_state_dict_named(state_dict, name + ".self_attn")
        self.encoder_attn.upgrade_state_dict_named(state_dict, name + ".encoder_attn")
        _upgrade_layer_norm_names(self, state_dict, name, {"0": "self_attn_layer_norm", "1": "encoder_attn_layer_norm",
                                                           "2": "final_layer_norm"})


def _rel_bucket_1d(token_rp_bucket):
    """bucket as a function of (i - j) only: index rel + 1023 for rel in [-1023, 1023]."""
    n = token_rp_bucket.shape[0]
    neg = token_rp_bucket[0, 1:].flip(0)      # rel = -(n-1) .. -1
    pos = token_rp_bucket[:, 0]               # rel = 0 .. n-1
    return torch.cat([neg, pos]).contiguous()


def _tok_lut(table_weight, rel_bucket_1d):
    return table_weight.index_select(0, rel_bucket_1d).t().float().contiguous()      # [H, 2n-1]


def _img_lut(table_weight):
    return table_weight.t().float().contiguous()                                     # [H, n_rel]


class TransformerEncoder(FairseqEncoder):
    """unify_transformer.py:493-1072."""

    def __init__(self, args, dictionary, embed_tokens):
        super().__init__(dictionary)
        self.args = args
        _unsupported(args, ["encoder_prompt", "adapter", "bitfit", "sync_bn", "interpolate_position",
                            "entangle_position_embedding", "scale_resids"])
        if args.attention_dropout or getattr(args, "activation_dropout", 0) or getattr(args, "relu_dropout", 0):
            raise NotImplementedError("attention / activation dropout are 0 in every Musketeer script and are not "
                                      "implemented in the attention / FFN kernels")
        self.dropout_p = float(args.dropout)
        self.register_buffer("version", torch.Tensor([3]))
        d = embed_tokens.embedding_dim
        self.padding_idx = embed_tokens.padding_idx
        self.max_source_positions = args.max_source_positions
        self.num_attention_heads = args.encoder_attention_heads
        self.embed_tokens = embed_tokens
        self.embed_scale = 1.0 if args.no_scale_embedding else math.sqrt(d)
        if self.embed_scale != 1.0:
            raise NotImplementedError("OFA archs set no_scale_embedding (ofa.py:405)")
        self.layernorm_embedding = nn.LayerNorm(d) if getattr(args, "layernorm_embedding", False) else None
        self.type_embedding = Embedding(2, d, padding_idx=None) if getattr(args, "add_type_embedding", False) else None
        self.embed_images = ResNetStem(args.resnet_type, frozen_bn=getattr(args, "freeze_resnet", False),
                                       drop_path_rate=args.resnet_drop_path_rate)
        self.image_proj = Linear(1024, d)
        self.patch_layernorm_embedding = nn.LayerNorm(d) if getattr(args, "patch_layernorm_embedding", False) else None
        self.embed_positions = Embedding(args.max_source_positions + 2, d)
        self.embed_image_positions = Embedding(args.image_bucket_size ** 2 + 1, d)
        self.pos_ln = nn.LayerNorm(d)
        self.image_pos_ln = nn.LayerNorm(d)
        self.pos_scaling = float(d / args.encoder_attention_heads * args.attn_scale_factor) ** -0.5
        self.pos_q_linear = nn.Linear(d, d)
        self.pos_k_linear = nn.Linear(d, d)
        dpr = [x.item() for x in torch.linspace(0, args.encoder_drop_path_rate, args.encoder_layers)]
        self.layers = nn.ModuleList([TransformerEncoderLayer(args, dpr[i]) for i in range(args.encoder_layers)])
        self.num_layers = len(self.layers)
        self.layer_norm = nn.LayerNorm(d) if args.encoder_normalize_before else None
        if self.layer_norm is None:
            raise NotImplementedError("post-LN encoders are outside the Musketeer flag set")
        n_tok = 2 * args.token_bucket_size - 1
        self.token_rel_pos_table_list = nn.ModuleList(
            [Embedding(n_tok, self.num_attention_heads, zero_init=True) for _ in range(args.encoder_layers)])
        n_img = (2 * args.image_bucket_size - 1) ** 2 + 3
        self.image_rel_pos_table_list = nn.ModuleList(
            [Embedding(n_img, self.num_attention_heads, zero_init=True) for _ in range(args.encoder_layers)])
        self.patch_image_size = getattr(args, "patch_image_size", 384)
        self.orig_patch_image_size = getattr(args, "orig_patch_image_size", 256)
        self.register_buffer("token_rp_bucket", make_token_bucket_position(args.token_bucket_size))
        self.register_buffer("image_rp_bucket", make_image_bucket_position(args.image_bucket_size, n_img))
        self._rel1d = None
        self.patch_orders_override = None   # tests: fix the random patch subset (reference uses python `random`)

    def rel_bucket_1d(self):
        if self._rel1d is None or self._rel1d.device != self.token_rp_bucket.device:
            self._rel1d = _rel_bucket_1d(self.token_rp_bucket)
        return self._rel1d

    def get_patch_images_info(self, patch_images, sample_patch_num, device, patch_features=None):
        if patch_features is not None:            # stem output computed upstream (several tasks in one grouped pass)
            feat, (h, w) = patch_features
        else:
            feat = self.embed_images(patch_images)                    # [B, h*w, 1024] (NHWC-flattened)
            h, w = self.embed_images.last_hw
        B = feat.size(0)
        P = h * w
        pid = (torch.arange(w, device=device).unsqueeze(0).expand(h, w) +
               torch.arange(h, device=device).unsqueeze(1) * self.args.image_bucket_size + 1).reshape(-1)
        pid = pid[None, :].expand(B, P)
        pad = torch.zeros(B, P, dtype=torch.bool, device=device)
        if sample_patch_num is not None:                              # unify_transformer.py:671-682
            if sample_patch_num > P:
                raise ValueError("Sample larger than population or is negative")     # what random.sample raises in the reference
            if self.patch_orders_override is not None:
                orders = self.patch_orders_override.to(device)
                if orders.shape[0] == 1:                                  # one subset for every row (tests)
                    orders = orders.expand(B, -1)
            elif feat.is_cuda:
                # a uniformly random ordered k-subset per row, like random.sample, drawn from the device generator: no host
                # loop, no host -> device copy, and capturable in the step's CUDA graph (fresh numbers on every replay)
                orders = torch.rand(B, P, device=device).argsort(dim=1)[:, :sample_patch_num]
            else:
                orders = torch.LongTensor([random.sample(range(P), k=sample_patch_num) for _ in range(B)]).to(device)
            feat = feat.gather(1, orders.unsqueeze(2).expand(-1, -1, feat.size(2)))
            P = sample_patch_num
            pad = pad.gather(1, orders)
            pid = pid.gather(1, orders)
        return feat.contiguous(), P, pad, pid.contiguous()

    def forward(self, src_tokens, src_lengths=None, patch_images=None, patch_images_2=None, patch_masks=None,
                code_masks=None, return_all_hiddens=False, token_embeddings=None, sample_patch_num=None,
                patch_features=None):
        if patch_images_2 is not None or token_embeddings is not None:
            raise NotImplementedError("patch_images_2 / token_embeddings are outside the hot-path scope")
        dev = src_tokens.device
        B, S = src_tokens.shape
        d, H = self.embed_tokens.embedding_dim, self.num_attention_heads
        w = self.embed_tokens.weight
        P, pid = 0, None
        pad_mask = src_tokens.eq(self.padding_idx)
        type_w = self.type_embedding.weight if self.type_embedding is not None else None
        # text embedding (+ type 0) -> LN                                            unify_transformer.py:725-733
        x = ops.embedding(src_tokens, w, type_w[0] if type_w is not None else None, self.padding_idx)
        if self.layernorm_embedding is not None:
            x = _ln(self.layernorm_embedding, x)
        x = ops.dropout_residual(x, None, self.dropout_p, 0.0, self.training)                # :733
        pos = ops.embedding(torch.arange(S, device=dev), self.embed_positions.weight)        # [S, d]   :885
        pos = _ln(self.pos_ln, pos).unsqueeze(0).expand(B, S, d)                             # :898
        if patch_images is not None or patch_features is not None:
            feat, P, img_pad, pid = self.get_patch_images_info(patch_images, sample_patch_num, dev, patch_features)
            img_pad = img_pad | (~patch_masks)[:, None]                                      # :872
            bias = self.image_proj.bias + type_w[1] if type_w is not None else self.image_proj.bias
            xi = ops.linear(feat.to(w.dtype), self.image_proj.weight, bias)                  # :739-744
            if self.patch_layernorm_embedding is not None:
                xi = _ln(self.patch_layernorm_embedding, xi)
            xi = ops.dropout_residual(xi, None, self.dropout_p, 0.0, self.training)          # :747
            x = torch.cat([xi, x], dim=1)
            pad_mask = torch.cat([img_pad, pad_mask], dim=1)
            ipos = _ln(self.image_pos_ln, ops.embedding(pid, self.embed_image_positions.weight))   # :695,900
            pos = torch.cat([ipos, pos], dim=1)
        pos = pos.contiguous()
        x = ops.mask_rows(x, pad_mask)                                                       # :892-893
        pq = _lin(self.pos_q_linear, pos, alpha=self.pos_scaling)                            # :906-911
        pk = _lin(self.pos_k_linear, pos)
        kpm = pad_mask.contiguous().view(torch.uint8)
        cfg = {"causal": False, "kpm": kpm,
               "bias": {"q_text_off": P, "k_text_off": P, "ibs": self.args.image_bucket_size,
                        "q_pid": pid.int().contiguous() if pid is not None else None,
                        "k_pid": None, "n_img_q": P, "n_img_k": P}}
        cfg["bias"]["k_pid"] = cfg["bias"]["q_pid"]
        if torch.is_grad_enabled() and pq.requires_grad and pk.requires_grad:
            cfg["pos_sink"] = {"n": len(self.layers)}        # the layers' d pos_q / d pos_k are summed in the kernels (ops._Attention)
        rel1d = self.rel_bucket_1d()
        states = []
        for i, layer in enumerate(self.layers):
            tok_lut = _tok_lut(self.token_rel_pos_table_list[i].weight, rel1d)
            img_lut = _img_lut(self.image_rel_pos_table_list[i].weight) if P else None
            x = layer(x, pq, pk, tok_lut, img_lut, cfg)
            if return_all_hiddens:
                states.append(x.transpose(0, 1))
        x = _ln(self.layer_norm, x)
        return {
            "encoder_out": [x.transpose(0, 1)],            # T x B x C (view; the reference layout)
            "encoder_padding_mask": [pad_mask],            # B x T
            "encoder_embedding": [],
            "encoder_states": states,
            "src_tokens": [],
            "src_lengths": [],
            "position_embeddings": [pos],                  # B x T x C
        }

    def forward_torchscript(self, net_input):
        return self.forward(**{k: v for k, v in net_input.items() if k != "prev_output_tokens"})

    def upgrade_state_dict_named(self, state_dict, name):
        """unify_transformer.py:1033-1072."""
        for i, layer in enumerate(self.layers):
            layer.upgrade_state_dict_named(state_dict, "{}.layers.{}".format(name, i))
        prefix = name + "." if name != "" else ""
        for k, v in self.state_dict().items():
            if prefix + k not in state_dict:
                state_dict[prefix + k] = v
        _grow_image_positions(self, state_dict, prefix + "embed_image_positions.weight")
        return state_dict

    def reorder_encoder_out(self, encoder_out, new_order):
        out = {}
        for k, dim in (("encoder_out", 1), ("encoder_padding_mask", 0), ("encoder_embedding", 0), ("src_tokens", 0),
                       ("src_lengths", 0), ("position_embeddings", 0)):
            out[k] = [encoder_out[k][0].index_select(dim, new_order)] if len(encoder_out[k]) else []
        out["encoder_states"] = [s.index_select(1, new_order) for s in encoder_out["encoder_states"]]
        return out

    def max_positions(self):
        return self.max_source_positions


class TransformerDecoder(FairseqIncrementalDecoder):
    """unify_transformer.py:1075-1659.  A FairseqIncrementalDecoder: the reference generator enables incremental decoding only
    for decoders of that type (models/sequence_generator.py:776-781)."""

    def __init__(self, args, dictionary, embed_tokens, no_encoder_attn=False):
        super().__init__(dictionary)
        self.args = args
        _unsupported(args, ["decoder_prompt", "adapter", "cross_self_attention", "no_cross_attention"])
        self.dropout_p = float(args.dropout)
        self.register_buffer("version", torch.Tensor([3]))
        d = args.decoder_embed_dim
        self.embed_dim = d
        self.num_attention_heads = args.decoder_attention_heads
        self.padding_idx = embed_tokens.padding_idx
        self.max_target_positions = args.max_target_positions
        self.embed_tokens = embed_tokens
        self.layernorm_embedding = nn.LayerNorm(d) if getattr(args, "layernorm_embedding", False) else None
        self.window_size = args.code_image_size // 8
        self.embed_positions = Embedding(args.max_target_positions + 2, d)
        self.embed_image_positions = Embedding(args.image_bucket_size ** 2 + 1, d)
        self.pos_ln = nn.LayerNorm(d)
        self.image_pos_ln = nn.LayerNorm(d)
        self.pos_scaling = float(d / self.num_attention_heads * args.attn_scale_factor) ** -0.5
        self.self_pos_q_linear = nn.Linear(d, d)
        self.self_pos_k_linear = nn.Linear(d, d)
        self.cross_pos_q_linear = nn.Linear(d, d)
        self.cross_pos_k_linear = nn.Linear(d, d)
        self.code_layernorm_embedding = nn.LayerNorm(d) if getattr(args, "code_layernorm_embedding", False) else None
        dpr = [x.item() for x in torch.linspace(0, args.decoder_drop_path_rate, args.decoder_layers)]
        self.layers = nn.ModuleList([TransformerDecoderLayer(args, dpr[i]) for i in range(args.decoder_layers)])
        self.num_layers = len(self.layers)
        self.layer_norm = nn.LayerNorm(d) if args.decoder_normalize_before else None
        if self.layer_norm is None:
            raise NotImplementedError("post-LN decoders are outside the Musketeer flag set")
        self.output_projection = nn.Linear(d, embed_tokens.weight.shape[0], bias=False)
        self.output_projection.weight = self.embed_tokens.weight            # :1248-1254 (tied)
        n_tok = 2 * args.token_bucket_size - 1
        self.token_rel_pos_table_list = nn.ModuleList(
            [Embedding(n_tok, self.num_attention_heads, zero_init=True) for _ in range(args.decoder_layers)])
        n_img = (2 * args.image_bucket_size - 1) ** 2 + 3
        ws = self.window_size
        ipi = torch.arange(ws).unsqueeze(0).expand(ws, ws) + torch.arange(ws).unsqueeze(1) * args.image_bucket_size + 1
        ipi = torch.cat([torch.tensor([0]), ipi.reshape(-1)])
        ipi = torch.cat([ipi, torch.tensor([1024] * 769)])
        self.image_rel_pos_table_list = nn.ModuleList(
            [Embedding(n_img, self.num_attention_heads, zero_init=True) for _ in range(args.decoder_layers)])
        self.register_buffer("token_rp_bucket", make_token_bucket_position(args.token_bucket_size))
        self.register_buffer("image_rp_bucket", make_image_bucket_position(args.image_bucket_size, n_img))
        self.register_buffer("image_position_idx", ipi)
        self._rel1d = None

    def rel_bucket_1d(self):
        if self._rel1d is None or self._rel1d.device != self.token_rp_bucket.device:
            self._rel1d = _rel_bucket_1d(self.token_rp_bucket)
        return self._rel1d

    def forward(self, prev_output_tokens, code_masks=None, encoder_out=None, incremental_state=None,
                features_only=False, full_context_alignment=False, alignment_layer=None, alignment_heads=None,
                src_lengths=None, return_all_hiddens=False, padded_logits=False):
        x, extra = self.extract_features(prev_output_tokens, code_masks, encoder_out, incremental_state,
                                         full_context_alignment)
        if not features_only:
            x = self.output_layer(x, padded=padded_logits)
        return x, extra

    def extract_features(self, prev_output_tokens, code_masks=None, encoder_out=None, incremental_state=None,
                         full_context_alignment=False, alignment_layer=None, alignment_heads=None):
        if code_masks is not None and bool(torch.any(code_masks)):
            raise NotImplementedError("image-code targets (code_masks) are outside the hot-path scope")
        if full_context_alignment:
            raise NotImplementedError("full_context_alignment is outside the hot-path scope")
        dev = prev_output_tokens.device
        B, T = prev_output_tokens.shape
        d, H = self.embed_dim, self.num_attention_heads
        enc = encoder_out["encoder_out"][0].transpose(0, 1)                     # [B, N, d]
        if not enc.is_contiguous():
            enc = enc.contiguous()
        enc_pad = encoder_out["encoder_padding_mask"][0]
        src_pos = encoder_out["position_embeddings"][0]
        incremental = incremental_state is not None
        t0 = T - 1 if incremental else 0                                        # first query position
        Tq = T - t0
        # position embeddings; in incremental mode only the last row is needed for the queries, while the self-attention
        # pos_k of earlier steps lives in the cache (the reference recomputes the whole prefix: :1449-1472)
        tpe = ops.embedding(torch.arange(t0, T, device=dev), self.embed_positions.weight)      # [Tq, d]
        tp = _ln(self.pos_ln, tpe).unsqueeze(0).expand(B, Tq, d).contiguous()
        spq = _lin(self.self_pos_q_linear, tp, alpha=self.pos_scaling)
        spk_new = _lin(self.self_pos_k_linear, tp)
        cpq = _lin(self.cross_pos_q_linear, tp, alpha=self.pos_scaling)
        toks = prev_output_tokens[:, t0:]
        x = ops.embedding(toks.contiguous(), self.embed_tokens.weight, None, self.padding_idx)
        if not self.args.disable_entangle:                                                     # :1483-1484
            x = ops.add(x, tpe.unsqueeze(0).expand(B, Tq, d).contiguous())
        if self.layernorm_embedding is not None:
            x = _ln(self.layernorm_embedding, x)
        x = ops.dropout_residual(x, None, self.dropout_p, 0.0, self.training)                  # :1495
        if incremental:
            # KV cache of the incremental decoder.  Cross-attention K / V / pos_k are projected ONCE per SENTENCE: when the
            # caller hands over the un-replicated encoder output (our SequenceGenerator), the G = rows / sentences beams of a
            # sentence read the same cache row through `sent_row`; a caller that replicated encoder_out per beam (the
            # reference generator) simply gets G = 1.  Self-attention K / V live in a PAGED pool: [slot][2*layers][page_len][d]
            # with a block table per row; a beam reorder copies table entries of the full pages and only the partial last page
            # (reorder_incremental_state_scripting).  The cross-attention position term cross_pos_q . cross_pos_k^T is the same
            # for every layer (unify_transformer.py:1461-1466): one launch per step writes it, the six layers add it.
            st = incremental_state.setdefault("_ofa_b200", {})
            PL = 16
            if "cpk" not in st or (t0 == 0 and st.get("reuse")):
                EB = enc.shape[0]
                if B % EB != 0:
                    raise ValueError("decoder rows (%d) must be a multiple of the encoder batch (%d)" % (B, EB))
                G = B // EB
                if G > 8:
                    raise NotImplementedError("more than 8 beams per sentence share a cache row group")
                cpk = _lin(self.cross_pos_k_linear, src_pos.contiguous())
                cross = [layer.encoder_attn.project_kv(enc) for layer in self.layers]
                pad8 = enc_pad.contiguous().view(torch.uint8)
                same = ("cpk" in st and st["G"] == G and st["cpk"].shape == cpk.shape and st["cpk"].dtype == cpk.dtype
                        and st["table"][0].shape[0] == B)
                if same:
                    # persistent state (the generator replays captured decoder steps that hold these pointers): refresh the
                    # contents in place
                    st["cpk"].copy_(cpk)
                    for (k0, v0), (k1, v1) in zip(st["cross"], cross):
                        k0.copy_(k1)
                        v0.copy_(v1)
                    st["enc_pad"].copy_(pad8)
                    st["sent_row"] = st["sent_row0"]
                else:
                    st["G"] = G
                    st["sent_row0"] = st["sent_row"] = torch.arange(EB, device=dev, dtype=torch.int32)
                    st["cpk"], st["cross"], st["enc_pad"] = cpk, cross, pad8.clone()
                    st["max_pages"] = 2
                    st["pool"] = torch.zeros(B * st["max_pages"] * 2, 2 * self.num_layers, PL, d, dtype=x.dtype, device=dev)
                    st["table"] = [torch.full((B, st["max_pages"]), -1, dtype=torch.int32, device=dev) for _ in range(2)]
                    st["spk"] = torch.zeros(1, st["max_pages"] * PL, d, dtype=x.dtype, device=dev)
                    st["zero_row"] = torch.zeros(B, dtype=torch.int32, device=dev)
                    st["bias"] = torch.empty(B, H, (enc.shape[1] + 3) // 4 * 4, dtype=torch.float32, device=dev)
                    st["home"] = torch.arange(B, device=dev, dtype=torch.int32)
                    st["layout"] = st.get("layout", 0) + 1          # captured steps of an older layout are stale
                st["tcur"], st["ppar"], st["len"], st["rows"], st["reordered_at"] = 0, 0, 0, B, -1
            if t0 != st["len"]:
                raise ValueError("incremental decoding expects one new token per call (cache holds %d, got position %d)"
                                 % (st["len"], t0))
            if t0 + 1 > st["max_pages"] * PL:          # grow (doubling): slots are (row * max_pages + page) * 2 + parity
                mp, mp2 = st["max_pages"], st["max_pages"] * 2
                B0 = st["table"][0].shape[0]
                pool = torch.zeros(B0 * mp2 * 2, 2 * self.num_layers, PL, d, dtype=x.dtype, device=dev)
                pool.view(B0, mp2, 2, -1)[:, :mp] = st["pool"].view(B0, mp, 2, -1)
                tabs = []
                for tb in st["table"]:
                    slot = tb.long()
                    new = ((slot // 2) // mp) * (mp2 * 2) + ((slot // 2) % mp) * 2 + slot % 2
                    t2 = torch.full((B0, mp2), -1, dtype=torch.int32, device=dev)
                    t2[:, :mp] = torch.where(tb >= 0, new.int(), tb)
                    tabs.append(t2)
                spk2 = torch.zeros(1, mp2 * PL, d, dtype=x.dtype, device=dev)
                spk2[:, :mp * PL] = st["spk"]
                st["pool"], st["table"], st["spk"], st["max_pages"] = pool, tabs, spk2, mp2
                st["layout"] = st.get("layout", 0) + 1
            page, off = t0 // PL, t0 % PL
            table = st["table"][st["tcur"]]
            if off == 0 and st["reordered_at"] != t0:       # a new page starts: every row's own slot
                st["ppar"] = 1 - st["ppar"]
                table[:B, page] = (st["home"][:B] * st["max_pages"] + page) * 2 + st["ppar"]
            st["spk"][0, t0] = spk_new[0, 0]            # pos_k depends on the position only: one shared row
            pstride = 2 * self.num_layers * PL * d
            self_dec = {"S": t0 + 1, "G": 1, "pk_row": st["zero_row"][:B], "q_pos": t0, "page": (table, PL, pstride)}
            bias = st["bias"]
            ops.attention_decode(cpq, None, st["cpk"], None, None, enc.shape[1], H, st["G"], st["sent_row"], None, st["enc_pad"],
                                 score_out=bias)
            cross_dec = {"S": enc.shape[1], "G": st["G"], "kv_row": st["sent_row"], "kpm": st["enc_pad"], "bias": bias}
            self_cfg = {"decode": self_dec}
            cross_cfg = {"decode": cross_dec}
            cpk, spk = st["cpk"], st["spk"]
        else:
            cpk = _lin(self.cross_pos_k_linear, src_pos.contiguous())
            spk = spk_new
            self_kpm = prev_output_tokens.eq(self.padding_idx).contiguous().view(torch.uint8)
            self_cfg = {"causal": True, "kpm": self_kpm, "q_pos_off": t0,
                        "bias": {"q_text_off": 0, "k_text_off": 0}}
            cross_cfg = {"causal": False, "kpm": enc_pad.contiguous().view(torch.uint8), "bias": {}, "fused_kv": True}
            if torch.is_grad_enabled() and spq.requires_grad and spk.requires_grad and cpq.requires_grad and cpk.requires_grad:
                self_cfg["pos_sink"] = {"n": len(self.layers)}
                cross_cfg["pos_sink"] = {"n": len(self.layers)}
            if enc.shape[0] != B:           # un-replicated encoder output: rows b*G .. b*G + G - 1 belong to sentence b
                if B % enc.shape[0] != 0:
                    raise ValueError("decoder rows (%d) must be a multiple of the encoder batch (%d)" % (B, enc.shape[0]))
                cross_cfg["rows_per_sentence"] = B // enc.shape[0]
        rel1d = self.rel_bucket_1d()
        inner_states = [x.transpose(0, 1)]
        for i, layer in enumerate(self.layers):
            w_rel = self.token_rel_pos_table_list[i].weight
            if incremental and not self.training and not torch.is_grad_enabled():
                # inference: the gathered table is a function of the weights only -- kept per layer until they change (their
                # version counter: optimizers and load_state_dict bump it), instead of 3 small launches per layer and step
                memo = self.__dict__.setdefault("_tok_lut_memo", {})
                key = (w_rel.data_ptr(), w_rel._version, w_rel.dtype)
                if memo.get(i, (None, None))[0] != key:
                    memo[i] = (key, _tok_lut(w_rel, rel1d))
                tok_lut = memo[i][1]
            else:
                tok_lut = _tok_lut(w_rel, rel1d)

            def self_kv(attn, h, i=i):
                q, k, v = attn.qkv_eval(h)
                st = incremental_state["_ofa_b200"]
                pool, tb = st["pool"], st["table"][st["tcur"]]
                ops.page_write(pool, tb[:B], k, v, i, t0 // 16, t0 % 16, 16)
                return q, pool[0, 2 * i], pool[0, 2 * i + 1]       # base of this layer's k / v planes inside page 0

            def cross_kv(attn, i=i):
                if incremental:
                    return incremental_state["_ofa_b200"]["cross"][i]
                memo = encoder_out.get("_cross_kv_memo") if not self.training else None
                if memo is not None:        # repeated teacher-forced passes over one encoder output (all-candidate scoring)
                    if i not in memo:
                        memo[i] = attn.project_kv(enc, fused=True)
                    return memo[i]
                return attn.project_kv(enc, fused=True)

            x = layer(x, self_kv, cross_kv, spq, spk, cpq, cpk, tok_lut, self_cfg, cross_cfg)
            inner_states.append(x.transpose(0, 1))
        if incremental:
            incremental_state["_ofa_b200"]["len"] = t0 + 1
        x = _ln(self.layer_norm, x)
        return x, {"attn": [None], "inner_states": inner_states}

    def output_layer(self, features, padded=False):
        """Tied projection to the vocabulary (:1577-1583).  padded=True returns a [B,T,V] view whose row stride is
        rounded up to 8 elements (16B-aligned rows for the fused loss kernel and the TMA-fed backward GEMMs)."""
        return ops.linear(features, self.output_projection.weight, None, 1.0, None, padded)

    def reorder_incremental_state_scripting(self, incremental_state, new_order):
        """Beam reorder (models/sequence_generator.py:339-349).  The self-attention cache is paged: the new rows' block tables
        take the parents' entries for every full page (no data moves), and only the valid prefix of the partial last page is
        copied into the row's own slot (one launch for all layers); the per-sentence cross-attention cache is never moved --
        only the group -> sentence map follows the surviving sentences."""
        st = incremental_state.get("_ofa_b200")
        if not st:
            return
        rows = int(new_order.numel())
        G = st["G"]
        t0 = st["len"]                       # the position the next decoder call writes
        if t0 > 0:
            page, off = t0 // 16, t0 % 16
            if off == 0 and page >= st["max_pages"]:
                pass                         # the pool grows in the next decoder call; full pages only: plain table copy below
            src, dst = st["table"][st["tcur"]], st["table"][1 - st["tcur"]]
            st["ppar"] = 1 - st["ppar"]
            if page < st["max_pages"]:
                ops.page_reorder(st["pool"], src, dst, new_order.contiguous(), rows, page, off, st["ppar"], 16)
                st["reordered_at"] = t0
            else:
                dst[:rows] = src.index_select(0, new_order)
            st["tcur"] = 1 - st["tcur"]
        st["sent_row"] = st["sent_row"].index_select(0, torch.div(new_order[::G], G, rounding_mode="floor"))
        st["rows"] = rows

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        logits = net_output[0]
        return F.log_softmax(logits, dim=-1, dtype=torch.float32) if log_probs else \
            F.softmax(logits, dim=-1, dtype=torch.float32)

    def max_positions(self):
        return self.max_target_positions

    def upgrade_state_dict_named(self, state_dict, name):
        """unify_transformer.py:1605-1659."""
        prefix = name + "." if name != "" else ""
        if prefix + "output_projection.weight" not in state_dict and prefix + "embed_tokens.weight" in state_dict:
            state_dict[prefix + "output_projection.weight"] = state_dict[prefix + "embed_tokens.weight"]    # tied (:1615-1626)
        for i, layer in enumerate(self.layers):
            layer.upgrade_state_dict_named(state_dict, "{}.layers.{}".format(name, i))
        own = self.state_dict()
        state_dict[prefix + "image_position_idx"] = own["image_position_idx"]       # always the model's own (:1640-1642)
        for k, v in own.items():
            if prefix + k not in state_dict:
                state_dict[prefix + k] = v
        _grow_image_positions(self, state_dict, prefix + "embed_image_positions.weight")
        return state_dict


class OFAModel(FairseqEncoderDecoderModel):
    """models/ofa/ofa.py:25-171.  Registered as fairseq model "ofa" (+ architectures ofa_tiny ... ofa_huge) by
    musketeer_b200.plugin; with fairseq on the path this IS a FairseqEncoderDecoderModel."""

    def __init__(self, args, encoder, decoder):
        super().__init__(encoder, decoder)
        self.args = args
        self.supports_align_args = True
        self.apply(init_bert_params)
        self.classification_heads = nn.ModuleDict()
        if hasattr(self.encoder, "dictionary"):
            self.eos = self.encoder.dictionary.eos()
        self.enc_timer, self.dec_timer, self.cls_timer = [0, 0], [0, 0], [0, 0]

    @staticmethod
    def add_args(parser):
        """Every flag of TransformerModel.add_args + OFAModel.add_args (unify_transformer.py:150-334, ofa.py:45-73)."""
        from .options import add_model_args
        add_model_args(parser)

    def half(self):
        """trainer.py:99-106 calls .half() under --fp16 (what train_musketeer.sh:174 passes).  The kernels of this package
        compute in bfloat16 (fp32 accumulation): same exponent range as fp32, so fairseq's fp16 loss scaler never overflows
        and the run proceeds exactly as with --bf16."""
        logger.warning("musketeer_b200: --fp16 requested; the sm_100a kernels run bfloat16 (fp32 accumulate) instead")
        return self.bfloat16()

    @classmethod
    def build_model(cls, args, task):
        from .archs import base_architecture
        base_architecture(args)
        if getattr(args, "max_source_positions", None) is None:
            args.max_source_positions = DEFAULT_MAX_SOURCE_POSITIONS
        if getattr(args, "max_target_positions", None) is None:
            args.max_target_positions = DEFAULT_MAX_TARGET_POSITIONS
        src_dict, tgt_dict = task.source_dictionary, task.target_dictionary
        if not args.share_all_embeddings:
            raise NotImplementedError("OFA is trained with --share-all-embeddings (train_musketeer.sh:127)")
        if src_dict != tgt_dict:
            raise ValueError("--share-all-embeddings requires a joined dictionary")
        if args.encoder_embed_dim != args.decoder_embed_dim:
            raise ValueError("--share-all-embeddings requires --encoder-embed-dim to match --decoder-embed-dim")
        args.vocab_size = len(src_dict)
        emb = Embedding(len(src_dict), args.encoder_embed_dim, src_dict.pad())
        args.share_decoder_input_output_embed = True
        if getattr(args, "freeze_encoder_embedding", False) or getattr(args, "freeze_decoder_embedding", False):
            emb.weight.requires_grad = False
        encoder = TransformerEncoder(args, src_dict, emb)
        decoder = TransformerDecoder(args, tgt_dict, emb)
        return cls(args, encoder, decoder)

    def forward(self, src_tokens, src_lengths, prev_output_tokens, patch_images=None, patch_images_2=None,
                patch_masks=None, code_masks=None, sample_patch_num=None, features_only=False,
                classification_head_name=None, token_embeddings=None, return_all_hiddens=False, alignment_layer=None,
                alignment_heads=None, task_name=None, padded_logits=False, patch_features=None, encoder_out=None):
        if patch_images is not None and patch_images.dtype != self.encoder.embed_tokens.weight.dtype:
            # the trainer casts float inputs with .half() under --fp16 (trainer.py:1227-1244)
            patch_images = patch_images.to(self.encoder.embed_tokens.weight.dtype)
        if classification_head_name is not None:
            raise NotImplementedError("classification heads are outside the hot-path scope (SURVEY.md 8)")
        if encoder_out is None:       # (a multi-task criterion may hand over its slice of a merged encoder pass)
            encoder_out = self.encoder(src_tokens, src_lengths=src_lengths, patch_images=patch_images,
                                       patch_masks=patch_masks, patch_images_2=patch_images_2,
                                       token_embeddings=token_embeddings, return_all_hiddens=return_all_hiddens,
                                       sample_patch_num=sample_patch_num, patch_features=patch_features)
        self.enc_timer[1] += 1
        x, extra = self.decoder(prev_output_tokens, code_masks=code_masks, encoder_out=encoder_out,
                                features_only=features_only, alignment_layer=alignment_layer,
                                alignment_heads=alignment_heads, src_lengths=src_lengths,
                                return_all_hiddens=return_all_hiddens, padded_logits=padded_logits)
        self.dec_timer[1] += 1
        self.cls_timer[1] += 1
        return x, extra

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        return self.decoder.get_normalized_probs(net_output, log_probs, sample)

    def get_targets(self, sample, net_output):
        return sample["target"]

    def max_positions(self):
        return (self.encoder.max_positions(), self.decoder.max_positions())

    def max_decoder_positions(self):
        return self.decoder.max_positions()

    def set_num_updates(self, num_updates):
        pass

    def register_embedding_tokens(self, ans2label_dict, src_dict, bpe):
        """ofa.py:173-186: BPE ids of the answer strings whose embeddings seed vocabulary rows added at load time."""
        self.ans_tensor_list = []
        for i in range(len(ans2label_dict)):
            ans = src_dict[-len(ans2label_dict) + i]
            ans = ans[5:-1].replace("_", " ")
            self.ans_tensor_list.append(src_dict.encode_line(line=bpe.encode(" {}".format(ans.lower())),
                                                             add_if_not_exist=False, append_eos=False).long())

    def register_classification_head(self, name, num_classes=None, inner_dim=None, use_two_images=False, **kwargs):
        raise NotImplementedError("classification heads are outside the hot-path scope (SURVEY.md 8; no Musketeer script uses them)")

    def upgrade_state_dict_named(self, state_dict, name):
        """Checkpoint compatibility (ofa.py:216-318): classification heads of the checkpoint that this model does not have are
        dropped; a trailing <mask> row is removed; a vocabulary that grew since the checkpoint gets new embedding rows (mean
        answer-token embedding when `register_embedding_tokens` ran, else N(0, d^-0.5)); encoder / decoder / layer upgrades."""
        prefix = name + "." if name != "" else ""
        self.encoder.upgrade_state_dict_named(state_dict, prefix + "encoder")
        self.decoder.upgrade_state_dict_named(state_dict, prefix + "decoder")
        head_prefix = prefix + "classification_heads."
        for k in [k for k in state_dict if k.startswith(head_prefix)]:
            logger.warning("deleting classification head (%s) from checkpoint not present in current model: %s",
                           k[len(head_prefix):].split(".")[0], k)
            del state_dict[k]
        emb_keys = ["encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "decoder.output_projection.weight"]
        loaded = state_dict["encoder.embed_tokens.weight"].size(0)
        n_dict = len(self.encoder.dictionary)
        has_mask = "<mask>" in self.encoder.dictionary if hasattr(self.encoder.dictionary, "__contains__") else True
        if loaded == n_dict + 1 and not has_mask:
            for k in emb_keys + ["encoder.output_projection.weight"]:
                if k in state_dict:
                    state_dict[k] = state_dict[k][:-1, :]
        if loaded < n_dict:
            w = state_dict["encoder.embed_tokens.weight"]
            new = torch.zeros(n_dict - loaded, w.size(1))
            if getattr(self, "ans_tensor_list", None):
                assert len(new) == len(self.ans_tensor_list)
                for i, ans in enumerate(self.ans_tensor_list):
                    e = F.embedding(ans.to(w.device), w)
                    new[i] = e.sum(0) / e.size(0)
            else:
                nn.init.normal_(new, mean=0, std=w.size(1) ** -0.5)
            new = new.to(dtype=w.dtype, device=w.device)
            for k in emb_keys:
                state_dict[k] = torch.cat([state_dict[k], new])
        return state_dict
