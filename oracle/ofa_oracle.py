"""CPU oracle: a restatement of the reference's OFA hot path in plain functional PyTorch (fp32).

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the product (`musketeer_b200/`) never does.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference modules from
/root/reference (through `oracle/ref_shim`) and this restatement on identical weights and batches and
commits the reference's outputs under `tests/golden/`; `tests/test_oracle_golden.py` re-checks this file
against those fixtures wherever the repo travels.  The reference itself ships no tests / golden vectors
(SURVEY.md section 4), so those fixtures are the pin.

Every function cites the reference lines it restates (paths relative to the reference root).
Weights come in as a flat `state_dict` with the reference's key names.
"""
import math
import random
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .synth import RESNET_BLOCKS, PAD, EOS, BOS, UNK

LN_EPS = 1e-5


def _ln(sd, p, x):
    # fairseq.modules.LayerNorm == torch.nn.LayerNorm(eps=1e-5) (un-vendored)
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], LN_EPS)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def gelu(x):
    # fairseq.utils.get_activation_fn("gelu"): erf-GELU computed in fp32 (unify_transformer_layer.py:140,280)
    return F.gelu(x.float()).type_as(x)


# ------------------------------------------------------------------------------------------------
# ResNet patch embedder      models/ofa/resnet.py:113-133,211-222 ; frozen_bn.py:36-57
# ------------------------------------------------------------------------------------------------
def _bn(sd, p, x, training, frozen, stats_out=None):
    w, b, rm, rv = sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]
    if frozen:  # FrozenBatchNorm2d (eps 1e-5): y = x * w*rsqrt(rv+eps) + (b - rm*scale)
        scale = w * (rv + 1e-5).rsqrt()
        bias = b - rm * scale
        return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)
    if training and stats_out is not None:
        # running-stat update of nn.BatchNorm2d: momentum 0.1, unbiased variance (SURVEY.md 7 hard part 6)
        with torch.no_grad():
            n = x.numel() / x.shape[1]
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            stats_out[p + ".running_mean"] = 0.9 * rm + 0.1 * mean
            stats_out[p + ".running_var"] = 0.9 * rv + 0.1 * var * n / max(n - 1, 1)
    if training:
        return F.batch_norm(x, None, None, w, b, True, 0.0, 1e-5)
    return F.batch_norm(x, rm, rv, w, b, False, 0.0, 1e-5)


def resnet_forward(sd, p, x, resnet_type, training=True, frozen=False, stats_out=None):
    bn = lambda q, t: _bn(sd, q, t, training, frozen, stats_out)
    x = F.conv2d(x, sd[p + ".conv1.weight"], None, stride=2, padding=3)
    x = F.relu(bn(p + ".bn1", x))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, n in enumerate(RESNET_BLOCKS[resnet_type], start=1):
        for bi in range(n):
            q = "%s.layer%d.%d" % (p, li, bi)
            stride = 2 if (bi == 0 and li > 1) else 1
            idn = x
            o = F.relu(bn(q + ".bn1", F.conv2d(x, sd[q + ".conv1.weight"])))
            o = F.relu(bn(q + ".bn2", F.conv2d(o, sd[q + ".conv2.weight"], None, stride, 1)))
            o = bn(q + ".bn3", F.conv2d(o, sd[q + ".conv3.weight"]))
            if bi == 0:
                idn = bn(q + ".downsample.1", F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride))
            x = F.relu(idn + o)  # drop_path is identity at rate 0 (resnet.py:111,130)
    return x


# ------------------------------------------------------------------------------------------------
# attention       models/ofa/unify_multihead_attention.py:117-409
# ------------------------------------------------------------------------------------------------
def attention(sd, p, cfg, query, key, H, bias, key_padding_mask=None, causal=False, kv_cache=None,
              static_kv=False):
    """query [B,T,d], key [B,S,d] (batch-first here; the reference is T x B x C).
    bias [B,H,T,S] additive (attn_bias, :350-351).  kv_cache: dict for incremental decoding (:269-307)."""
    B, T, d = query.shape
    hd = d // H
    scaling = float(hd * cfg.attn_scale_factor) ** -0.5            # :58
    q = _lin(sd, p + ".q_proj", query) * scaling                     # :214,232
    if kv_cache is not None and static_kv and "k" in kv_cache:
        k, v = kv_cache["k"], kv_cache["v"]                         # :207-209,275-276
    else:
        k = _lin(sd, p + ".k_proj", key).view(B, -1, H, hd).transpose(1, 2)
        v = _lin(sd, p + ".v_proj", key).view(B, -1, H, hd).transpose(1, 2)
        if kv_cache is not None:
            if not static_kv and "k" in kv_cache:
                k = torch.cat([kv_cache["k"], k], dim=2)            # :279,289
                v = torch.cat([kv_cache["v"], v], dim=2)
            kv_cache["k"], kv_cache["v"] = k, v
    q = q.view(B, T, H, hd).transpose(1, 2)
    w = torch.matmul(q, k.transpose(2, 3))                          # :345
    S = k.shape[2]
    if bias is not None:
        w = w + bias[..., -S:]                                      # :350-351
    if causal:
        w = w + torch.triu(torch.full((T, S), float("-inf"), dtype=w.dtype), 1 + S - T)   # :353-357, unify_transformer.py:1591-1603
    if key_padding_mask is not None:
        w = w.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))   # :363-375
    pr = F.softmax(w, dim=-1, dtype=torch.float32).type_as(w)       # :380-383
    o = torch.matmul(pr, v)                                         # :387
    if cfg.scale_heads:
        o = o * sd[p + ".c_attn"].view(1, H, 1, 1)                  # :395-398
    o = o.transpose(1, 2).reshape(B, T, d)
    return _lin(sd, p + ".out_proj", o)                             # :399


def ffn(sd, p, cfg, x):
    # unify_transformer_layer.py:277-285 / :553-562
    h = gelu(_lin(sd, p + ".fc1", x))
    if cfg.scale_fc:
        h = _ln(sd, p + ".ffn_layernorm", h)
    return _lin(sd, p + ".fc2", h)


# ------------------------------------------------------------------------------------------------
# encoder         models/ofa/unify_transformer.py:660-697,713-769,819-966
# ------------------------------------------------------------------------------------------------
def encoder_forward(sd, cfg, src_tokens, patch_images=None, patch_masks=None, sample_patch_num=None,
                    training=True, patch_orders=None, stats_out=None):
    e = "encoder"
    B, S = src_tokens.shape
    d, H = cfg.encoder_embed_dim, cfg.encoder_attention_heads
    hd = d // H
    P = 0
    pid = None
    if patch_images is not None:
        feat = resnet_forward(sd, e + ".embed_images", patch_images, cfg.resnet_type, training,
                              cfg.freeze_resnet, stats_out)                        # :661
        h, w = feat.shape[-2:]
        P = h * w
        pid = (torch.arange(w)[None, :] + torch.arange(h)[:, None] * cfg.image_bucket_size + 1).reshape(-1)
        pid = pid[None, :].expand(B, P)                                            # :665-668
        img_pad = torch.zeros(B, P, dtype=torch.bool)
        feat = feat.flatten(2).transpose(1, 2)                                     # :670
        if sample_patch_num is not None:                                           # :671-682
            if patch_orders is None:
                patch_orders = torch.LongTensor(
                    [random.sample(range(P), k=sample_patch_num) for _ in range(B)])
            feat = feat.gather(1, patch_orders.unsqueeze(2).expand(-1, -1, feat.size(2)))
            P = sample_patch_num
            img_pad = img_pad.gather(1, patch_orders)
            pid = pid.gather(1, patch_orders)
        img_pos = F.embedding(pid, sd[e + ".embed_image_positions.weight"])        # :695
        img_pad = img_pad.clone()
        img_pad[~patch_masks] = True                                               # :872
    pad_mask = src_tokens.eq(PAD)                                                  # :878
    if P:
        pad_mask = torch.cat([img_pad, pad_mask], dim=1)                           # :880
    has_pads = bool(pad_mask.any())                                                # :883

    pos = F.embedding(torch.arange(S).expand(B, S), sd[e + ".embed_positions.weight"])   # :885
    # forward_embedding :713-769 (embed_scale = 1; entangle_position_embedding False)
    x = F.embedding(src_tokens, sd[e + ".embed_tokens.weight"])
    if cfg.entangle_position_embedding:
        x = x + pos
    if cfg.add_type_embedding:
        x = x + sd[e + ".type_embedding.weight"][0]
    if cfg.layernorm_embedding:
        x = _ln(sd, e + ".layernorm_embedding", x)
    if P:
        xi = _lin(sd, e + ".image_proj", feat)
        if cfg.entangle_position_embedding:
            xi = xi + img_pos
        if cfg.add_type_embedding:
            xi = xi + sd[e + ".type_embedding.weight"][1]
        if cfg.patch_layernorm_embedding:
            xi = _ln(sd, e + ".patch_layernorm_embedding", xi)
        x = torch.cat([xi, x], dim=1)
    if has_pads:
        x = x * (1 - pad_mask.unsqueeze(-1).type_as(x))                            # :892-893

    pos = _ln(sd, e + ".pos_ln", pos)                                              # :898
    if P:
        pos = torch.cat([_ln(sd, e + ".image_pos_ln", img_pos), pos], dim=1)       # :900-901
    N = P + S
    pos_scaling = float(hd * cfg.attn_scale_factor) ** -0.5                        # :574
    pq = (_lin(sd, e + ".pos_q_linear", pos).view(B, N, H, hd).transpose(1, 2)) * pos_scaling
    pk = _lin(sd, e + ".pos_k_linear", pos).view(B, N, H, hd).transpose(1, 2)
    abs_bias = torch.matmul(pq, pk.transpose(2, 3))                                # :906-912

    tok_bucket = sd[e + ".token_rp_bucket"][:S, :S]
    for i in range(cfg.encoder_layers):
        bias = abs_bias.clone()                                                    # :923
        tb = F.embedding(tok_bucket, sd["%s.token_rel_pos_table_list.%d.weight" % (e, i)])   # :640-646
        bias[:, :, N - S:, N - S:] += tb.permute(2, 0, 1).unsqueeze(0)             # :924
        if P:
            rb = sd[e + ".image_rp_bucket"][pid[:, :, None], pid[:, None, :]]      # :648-655
            ib = F.embedding(rb, sd["%s.image_rel_pos_table_list.%d.weight" % (e, i)])
            bias[:, :, :P, :P] += ib.permute(0, 3, 1, 2)                           # :930-932
        p = "%s.layers.%d" % (e, i)
        # TransformerEncoderLayer.forward  unify_transformer_layer.py:257-293
        r = x
        h = _ln(sd, p + ".self_attn_layer_norm", x)
        y = attention(sd, p + ".self_attn", cfg, h, h, H, bias, pad_mask if has_pads else None)
        if cfg.scale_attn:
            y = _ln(sd, p + ".attn_ln", y)
        x = r + y
        r = x
        x = r + ffn(sd, p, cfg, _ln(sd, p + ".final_layer_norm", x))
    x = _ln(sd, e + ".layer_norm", x)                                              # :950-951
    return {"encoder_out": x, "encoder_padding_mask": pad_mask, "position_embeddings": pos,
            "patch_orders": patch_orders}


# ------------------------------------------------------------------------------------------------
# decoder         models/ofa/unify_transformer.py:1297-1318,1393-1583
# ------------------------------------------------------------------------------------------------
def _dec_pos_bias(sd, cfg, qn, kn, tgt_pos, src_pos=None):
    dd = "decoder"
    H = cfg.decoder_attention_heads
    B, T, d = tgt_pos.shape
    hd = d // H
    s = float(hd * cfg.attn_scale_factor) ** -0.5                                   # :1157
    tp = _ln(sd, dd + ".pos_ln", tgt_pos)                                           # :1300
    kp = tp if src_pos is None else src_pos   # encoder position_embeddings are already LN'd (:898-904,1461)
    pq = _lin(sd, dd + "." + qn, tp).view(B, T, H, hd).transpose(1, 2) * s
    pk = _lin(sd, dd + "." + kn, kp).view(B, -1, H, hd).transpose(1, 2)
    return torch.matmul(pq, pk.transpose(2, 3))


def decoder_forward(sd, cfg, prev_output_tokens, enc, incremental_state=None, features_only=False):
    """incremental_state: None (teacher forcing) or a dict holding per-layer KV caches
    (unify_multihead_attention.py:269-307); as in the reference the whole prefix is passed and only the
    last position is computed (unify_transformer.py:1468-1472,1528-1529)."""
    dd = "decoder"
    H = cfg.decoder_attention_heads
    B, T = prev_output_tokens.shape
    tgt_pos = F.embedding(torch.arange(T).expand(B, T), sd[dd + ".embed_positions.weight"])   # :1449-1450
    self_bias = _dec_pos_bias(sd, cfg, "self_pos_q_linear", "self_pos_k_linear", tgt_pos)    # :1456
    cross_bias = _dec_pos_bias(sd, cfg, "cross_pos_q_linear", "cross_pos_k_linear", tgt_pos,
                               enc["position_embeddings"])                                    # :1461-1462
    toks = prev_output_tokens
    if incremental_state is not None:
        toks = toks[:, -1:]
        cross_bias = cross_bias[:, :, -1:, :]
        tgt_pos = tgt_pos[:, -1:, :]                                                          # :1469-1472
    x = F.embedding(toks, sd[dd + ".embed_tokens.weight"])                                    # :1475
    if not cfg.disable_entangle:
        x = x + tgt_pos                                                                       # :1483-1484
    if cfg.layernorm_embedding:
        x = _ln(sd, dd + ".layernorm_embedding", x)                                           # :1486-1488
    self_pad = toks.eq(PAD) if bool(toks.eq(PAD).any()) else None                             # :1500-1502
    tok_bucket = sd[dd + ".token_rp_bucket"][:T, :T]
    enc_pad = enc["encoder_padding_mask"]
    enc_pad = enc_pad if bool(enc_pad.any()) else None
    for i in range(cfg.decoder_layers):
        p = "%s.layers.%d" % (dd, i)
        tb = F.embedding(tok_bucket, sd["%s.token_rel_pos_table_list.%d.weight" % (dd, i)])   # :1282-1287
        bias = self_bias + tb.permute(2, 0, 1).unsqueeze(0)                                   # :1519-1521
        cache = None
        if incremental_state is not None:
            bias = bias[:, :, -1:, :]                                                         # :1528-1529
            cache = incremental_state.setdefault(i, {"self": {}, "cross": {}})
        # TransformerDecoderLayer.forward  unify_transformer_layer.py:464-582
        r = x
        h = _ln(sd, p + ".self_attn_layer_norm", x)
        pad = self_pad
        if cache is not None and pad is not None:
            raise NotImplementedError("pad tokens never enter incremental decoding in the configs")
        y = attention(sd, p + ".self_attn", cfg, h, h, H, bias, pad, causal=incremental_state is None,
                      kv_cache=None if cache is None else cache["self"])
        if cfg.scale_attn:
            y = _ln(sd, p + ".self_attn_ln", y)
        x = r + y
        r = x
        h = _ln(sd, p + ".encoder_attn_layer_norm", x)
        y = attention(sd, p + ".encoder_attn", cfg, h, enc["encoder_out"], H, cross_bias, enc_pad,
                      kv_cache=None if cache is None else cache["cross"], static_kv=True)
        if cfg.scale_attn:
            y = _ln(sd, p + ".cross_attn_ln", y)
        x = r + y
        r = x
        x = r + ffn(sd, p, cfg, _ln(sd, p + ".final_layer_norm", x))
    x = _ln(sd, dd + ".layer_norm", x)                                                        # :1566-1567
    if features_only:
        return x
    return F.linear(x, sd[dd + ".output_projection.weight"])                                  # :1577-1583


def model_forward(sd, cfg, net_input, training=True, patch_orders=None, stats_out=None):
    """OFAModel.forward  models/ofa/ofa.py:80-171 (timers/syncs dropped)."""
    enc = encoder_forward(sd, cfg, net_input["src_tokens"], net_input.get("patch_images"),
                          net_input.get("patch_masks"), net_input.get("sample_patch_num"),
                          training, patch_orders, stats_out)
    return decoder_forward(sd, cfg, net_input["prev_output_tokens"], enc), enc


# ------------------------------------------------------------------------------------------------
# criterion       criterions/label_smoothed_cross_entropy.py:56-126,167-275
# ------------------------------------------------------------------------------------------------
def label_smoothed_loss(logits, target, epsilon, constraint_masks=None, conf=None, use_rdrop=False,
                        reg_alpha=1.0, ignore_prefix_size=0, drop_worst_ratio=0.0, drop_worst_after=0,
                        update_num=0, constraint_range=None):
    """logits [B,T,V], target [B,T] -> (loss, nll_loss, ntokens).  Masked vocabulary entries contribute 0 to
    the R-Drop KL (torch-1.8 kl_div semantics; torch 2.11 would give NaN)."""
    x = logits
    if constraint_masks is not None:
        x = x.masked_fill(~constraint_masks, -math.inf)                                       # :233
    cs = ce = None
    if constraint_range is not None:
        cs, ce = constraint_range
        x = x.clone()
        x[:, :, 4:cs] = -math.inf                                                             # :235-236
        x[:, :, ce:] = -math.inf
    lp = F.log_softmax(x, dim=-1, dtype=torch.float32)                                        # :237
    if conf is not None:
        lp = lp * conf[:, None, None]
    if ignore_prefix_size > 0:                                                                # :239-243
        lp = lp[:, ignore_prefix_size:]
        target = target[:, ignore_prefix_size:]
        if constraint_masks is not None:
            constraint_masks = constraint_masks[:, ignore_prefix_size:]
    V = lp.shape[-1]
    lp = lp.reshape(-1, V)
    tg = target.reshape(-1)
    keep = tg != PAD                                                                          # :257-260
    lp, tg = lp[keep], tg[keep]
    cm = constraint_masks.reshape(-1, V)[keep] if constraint_masks is not None else None
    nll = -lp.gather(1, tg[:, None]).squeeze(1)                                               # :88
    if cm is not None:
        smooth = -lp.masked_fill(~cm, 0).sum(-1)                                              # :90-91
        eps_i = epsilon / (cm.sum(1) - 1 + 1e-6)
    elif cs is not None:
        rng = [0, 1, 2, 3] + list(range(cs, ce))                                              # :93-95
        smooth = -lp[:, rng].sum(-1)
        eps_i = epsilon / (len(rng) - 1 + 1e-6)
    else:
        smooth = -lp.sum(-1)                                                                  # :97-98
        eps_i = epsilon / (V - 1)
    loss = (1.0 - epsilon - eps_i) * nll + eps_i * smooth                                     # :99
    if drop_worst_ratio > 0 and update_num > drop_worst_after:                                # :100-111
        if use_rdrop:
            tb = loss.size(0) // 2
            _, idx = torch.topk(loss[:tb], k=int(tb * (1 - drop_worst_ratio)), largest=False)
            loss = torch.cat([loss[idx], loss[idx + tb]])
            nll = torch.cat([nll[idx], nll[idx + tb]])
            lp = torch.cat([lp[idx], lp[idx + tb]])
        else:
            loss, idx = torch.topk(loss, k=int(loss.shape[0] * (1 - drop_worst_ratio)), largest=False)
            nll = nll[idx]
            lp = lp[idx]
    ntokens = loss.numel()
    nll_sum, loss_sum = nll.sum(), loss.sum()
    if use_rdrop:                                                                             # :116-124
        tb = lp.size(0) // 2
        p_, q_ = lp[:tb], lp[tb:]
        if cs is not None:
            p_, q_ = p_[:, rng], q_[:, rng]
        # kl_div(p, exp(q)) = sum exp(q)*(q-p); entries with exp(.)==0 (constraint-masked, -inf) contribute 0
        # to the value and to both gradients (torch-1.8.1 kl_div semantics the reference pins; :74-78)
        m = (p_ > -math.inf) & (q_ > -math.inf)
        pm, qm = p_.masked_fill(~m, 0.0), q_.masked_fill(~m, 0.0)
        t1 = (qm.exp() * (qm - pm)).masked_fill(~m, 0.0).sum()
        t2 = (pm.exp() * (pm - qm)).masked_fill(~m, 0.0).sum()
        loss_sum = loss_sum + reg_alpha * (t1 + t2) / 2
    return loss_sum, nll_sum, ntokens


def _rdrop(x):
    # construct_rdrop_sample  :56-71
    if isinstance(x, dict):
        return {k: _rdrop(v) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        return x.repeat(2, *([1] * (x.dim() - 1)))
    if isinstance(x, bool) or x is None:
        return x
    if isinstance(x, int):
        return x * 2
    return x


def criterion_forward(sd, cfg, sample, epsilon=0.1, use_rdrop=False, reg_alpha=1.0, sample_patch_num=0,
                      training=True, patch_orders=None, update_num=0, stats_out=None, **loss_kw):
    """AdjustLabelSmoothedCrossEntropyCriterion.forward :167-226 incl. the multi-task list recursion
    (:175-202): total = sum_t loss_t / sample_size_t, sample_size 1.
    patch_orders: optional list (one per task) fixing the random patch subset for parity."""
    if isinstance(sample, list) and len(sample) > 1:
        total, logs = 0.0, []
        for ti, s in enumerate(sample):
            s = dict(s)
            s["net_input"] = dict(s["net_input"])
            if sample_patch_num > 0 and ti < len(sample) - 1:                                  # :177-178
                s["net_input"]["sample_patch_num"] = sample_patch_num
            po = patch_orders[ti] if patch_orders is not None else None
            l, ss, lg = criterion_forward(sd, cfg, s, epsilon, use_rdrop, reg_alpha, 0, training, po,
                                          update_num, stats_out, **loss_kw)
            total = total + l / ss                                                            # :185
            logs.append(lg)
        return total, 1, {"tasks": logs}
    sample = sample[0] if isinstance(sample, list) else sample
    if use_rdrop:
        sample = _rdrop(sample)                                                               # :206-207
    logits, enc = model_forward(sd, cfg, sample["net_input"], training, patch_orders, stats_out)
    loss, nll, ntokens = label_smoothed_loss(
        logits, sample["target"], epsilon, sample.get("constraint_masks"), sample.get("conf"),
        use_rdrop, reg_alpha, update_num=update_num, **loss_kw)
    return loss, ntokens, {"loss": loss.detach(), "nll_loss": nll.detach(), "ntokens": ntokens,
                           "logits": logits, "patch_orders": enc["patch_orders"]}


# ------------------------------------------------------------------------------------------------
# beam search     models/sequence_generator.py:209-598,637-746 ; models/search.py:109-144
# ------------------------------------------------------------------------------------------------
class Trie:
    # utils/trie.py:9-30 restated with plain dicts: insert a token list; next layer of a prefix ([eos] once the prefix left the trie)
    def __init__(self, eos):
        self.root, self.eos = {}, eos

    def insert(self, word):
        cur = self.root
        for c in word:
            cur = cur.setdefault(c, {})

    def get_next_layer(self, word):
        cur = self.root
        for c in word:
            cur = cur.get(c)
            if cur is None:
                return [self.eos]
        return list(cur.keys())


def _first_beam(t, mask, beam):
    # replicate_first_beam :636-639
    t = t.view(-1, beam, t.size(-1))
    t[mask] = t[mask][:, :1, :]
    return t.view(-1, t.size(-1))


@torch.no_grad()
def generate(sd, cfg, net_input, beam=5, max_len_a=0, max_len_b=16, min_len=1, len_penalty=1.0,
             temperature=1.0, no_repeat_ngram_size=0, unk_penalty=0.0, constraint_trie=None, constraint_range=None,
             zero_shot=False, prefix_tokens=None):
    cstart = cend = None
    if constraint_range is not None:                                                            # :82-86
        cstart, cend = (int(v) for v in constraint_range.split(","))
    src = net_input["src_tokens"]
    bsz, src_len = src.shape
    V = cfg.vocab_size
    max_len = int(max_len_a * src_len + max_len_b)                                             # :267
    enc = encoder_forward(sd, cfg, src, net_input.get("patch_images"), net_input.get("patch_masks"),
                          training=False)
    order = torch.arange(bsz).view(-1, 1).repeat(1, beam).view(-1)                              # :276
    enc = {k: (v.index_select(0, order) if isinstance(v, torch.Tensor) else v) for k, v in enc.items()}
    scores = torch.zeros(bsz * beam, max_len + 1)
    tokens = torch.full((bsz * beam, max_len + 2), PAD, dtype=torch.long)
    tokens[:, 0] = BOS                                                                          # :293
    cands_to_ignore = torch.zeros(bsz, beam, dtype=torch.bool)
    finalized = [[] for _ in range(bsz)]
    finished = [False] * bsz
    num_remaining = bsz
    cand_size = 2 * beam
    bbsz_offsets = (torch.arange(bsz) * beam).unsqueeze(1)
    cand_offsets = torch.arange(cand_size)
    inc = {}
    reorder_state = None
    batch_idxs = None
    for step in range(max_len + 1):
        if reorder_state is not None:                                                           # :337-350
            if batch_idxs is not None:
                corr = batch_idxs - torch.arange(batch_idxs.numel())
                reorder_state.view(-1, beam).add_(corr.unsqueeze(-1) * beam)
            for lc in inc.values():
                lc["self"] = {k: v.index_select(0, reorder_state) for k, v in lc["self"].items()}
                lc["cross"] = {k: v.index_select(0, reorder_state) for k, v in lc["cross"].items()}
            enc = {k: (v.index_select(0, reorder_state) if isinstance(v, torch.Tensor) else v)
                   for k, v in enc.items()}
        logits = decoder_forward(sd, cfg, tokens[:, :step + 1], enc, incremental_state=inc)
        lg = logits[:, -1, :] / temperature                                                    # :852
        allowed = None
        if constraint_trie is not None:                                                         # :857-868, :878-885
            allowed = torch.zeros(lg.shape, dtype=torch.bool)
            for r, pref in enumerate(tokens[:, :step + 1].tolist()):
                # (the pre-softmax mask skips the forced prefix of the sentence, :862-868; the zero-shot one does not, :881-884)
                plen = int(prefix_tokens[r // beam].ne(PAD).sum()) if prefix_tokens is not None and not zero_shot else 0
                if len(pref) > plen:
                    allowed[r, constraint_trie.get_next_layer([0] + pref[plen + 1:])] = True
                else:
                    allowed[r] = True
        if not zero_shot:
            if allowed is not None:
                lg = lg.masked_fill(~allowed, -math.inf)
            if cstart is not None:                                                              # :869-872
                lg[:, 4:cstart] = -math.inf
                lg[:, cend:] = -math.inf
        lprobs = F.log_softmax(lg, dim=-1, dtype=torch.float32)                                # :875
        if zero_shot:
            if allowed is not None:
                lprobs = lprobs.masked_fill(~allowed, -math.inf)
            if cstart is not None:                                                              # :886-889
                lprobs[:, 4:cstart] = -math.inf
                lprobs[:, cend:] = -math.inf
        if prefix_tokens is not None and step < prefix_tokens.size(1) and step < max_len:       # :373-380, :600-634
            ptoks = prefix_tokens[:, step].unsqueeze(-1).repeat(1, beam).view(-1)
            plp = lprobs.gather(-1, ptoks.unsqueeze(-1))
            pmask = ptoks.ne(PAD)
            lprobs[pmask] = (torch.min(plp) - 1) if constraint_trie is None else -math.inf
            lprobs[pmask] = lprobs[pmask].scatter(-1, ptoks[pmask].unsqueeze(-1), plp[pmask])
            emask = ptoks.eq(EOS)
            if emask.any():           # a prefix that ends here: every beam of the sentence continues from the first one
                eb = emask.view(-1, beam)[:, 0]
                first = tokens[emask].view(-1, beam, tokens.size(-1))[:, 0, 1:step + 1]
                assert (first == prefix_tokens[eb][:, :step]).all()
                tokens, scores, lprobs = (_first_beam(t, eb, beam) for t in (tokens, scores, lprobs))
        elif step < min_len:
            lprobs[:, EOS] = -math.inf                                                          # :381-383
        lprobs[lprobs != lprobs] = -math.inf
        lprobs[:, PAD] = -math.inf                                                              # :387
        lprobs[:, UNK] -= unk_penalty
        if step >= max_len:                                                                     # :400-402
            lprobs[:, :EOS] = -math.inf
            lprobs[:, EOS + 1:] = -math.inf
        if no_repeat_ngram_size > 0:
            lprobs = _ngram_block(tokens, lprobs, step, no_repeat_ngram_size)                  # :425-426
        # BeamSearch.step  models/search.py:117-144
        lp3 = lprobs.view(bsz, -1, V)
        if step == 0:
            lp3 = lp3[:, ::beam, :].contiguous()
        else:
            lp3 = lp3 + scores.view(bsz, beam, -1)[:, :, step - 1].unsqueeze(-1)
        flat = lp3.view(bsz, -1)
        cand_scores, idx = torch.topk(flat, k=min(cand_size, flat.size(1) - 1))
        cand_beams = idx // V
        cand_indices = idx.fmod(V)
        cand_bbsz_idx = cand_beams + bbsz_offsets                                               # :440
        eos_mask = cand_indices.eq(EOS) & cand_scores.ne(-math.inf)                             # :444
        eos_mask[:, :beam][cands_to_ignore] = False
        eos_bbsz_idx = torch.masked_select(cand_bbsz_idx[:, :beam], eos_mask[:, :beam])
        finalized_sents = []
        if eos_bbsz_idx.numel() > 0:
            eos_scores = torch.masked_select(cand_scores[:, :beam], eos_mask[:, :beam])
            finalized_sents = _finalize(step, eos_bbsz_idx, eos_scores, tokens, scores, finalized, finished,
                                        beam, max_len, len_penalty)
            num_remaining -= len(finalized_sents)
        if num_remaining == 0:
            break
        assert step < max_len
        if len(finalized_sents) > 0:                                                            # :484-518
            new_bsz = bsz - len(finalized_sents)
            batch_mask = torch.ones(bsz, dtype=torch.bool)
            batch_mask[finalized_sents] = False
            batch_idxs = torch.arange(bsz).masked_select(batch_mask)
            eos_mask = eos_mask[batch_idxs]
            cand_beams = cand_beams[batch_idxs]
            bbsz_offsets = bbsz_offsets[:new_bsz]
            cand_bbsz_idx = cand_beams + bbsz_offsets
            cand_scores = cand_scores[batch_idxs]
            cand_indices = cand_indices[batch_idxs]
            cands_to_ignore = cands_to_ignore[batch_idxs]
            if prefix_tokens is not None:
                prefix_tokens = prefix_tokens[batch_idxs]                                       # :507-508
            scores = scores.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
            tokens = tokens.view(bsz, -1)[batch_idxs].view(new_bsz * beam, -1)
            bsz = new_bsz
        else:
            batch_idxs = None
        eos_mask[:, :beam] = ~((~cands_to_ignore) & (~eos_mask[:, :beam]))                      # :528
        active_mask = eos_mask.long() * cand_size + cand_offsets[: eos_mask.size(1)]
        new_ignore, active_hypos = torch.topk(active_mask, k=beam, dim=1, largest=False)        # :539
        cands_to_ignore = new_ignore.ge(cand_size)[:, :beam]
        active_bbsz_idx = torch.gather(cand_bbsz_idx, 1, active_hypos).view(-1)
        tokens[:, :step + 1] = torch.index_select(tokens[:, :step + 1], 0, active_bbsz_idx)
        tokens.view(bsz, beam, -1)[:, :, step + 1] = torch.gather(cand_indices, 1, active_hypos)
        if step > 0:
            scores[:, :step] = torch.index_select(scores[:, :step], 0, active_bbsz_idx)
        scores.view(bsz, beam, -1)[:, :, step] = torch.gather(cand_scores, 1, active_hypos)
        reorder_state = active_bbsz_idx
    for s in range(len(finalized)):                                                             # :589-597
        sc = torch.tensor([float(h["score"]) for h in finalized[s]])
        _, o = torch.sort(sc, descending=True)
        finalized[s] = [finalized[s][i] for i in o]
    return finalized


def _collate(items, pad):
    # data/data_utils.py collate_tokens, left_pad=False: right-padded stack (1-D token lists or 2-D bool mask rows)
    n = max(v.size(0) for v in items)
    out = items[0].new_full((len(items), n) + tuple(items[0].shape[1:]), pad)
    for i, v in enumerate(items):
        out[i, :v.size(0)] = v
    return out


def candidate_masks(answers, trie, vocab):
    # tasks/mm_tasks/vqa_gen.py:170-176: mask row i of an answer = next layer of the trie after [bos] + answer[:i]
    masks = []
    for ans in answers:
        m = torch.zeros(len(ans) + 1, vocab, dtype=torch.bool)
        for i in range(len(ans) + 1):
            m[i, trie.get_next_layer([BOS] + ans[:i].tolist())] = True
        masks.append(m)
    return masks


@torch.no_grad()
def score_all_candidates(sd, cfg, net_input, decoder_prompts, answers, trie, valid_batch_size):
    """All-candidate inference: utils/eval_utils.py:161-214 (= tasks/mm_tasks/vqa_gen.py:257-306, snli_ve.py:171-215).
    Encoder once; per chunk of `valid_batch_size` answers the decoder scores prompt + answer + eos teacher-forced for every
    (sentence, answer) pair with the encoder output replicated per answer; logits outside the trie's next layer are -inf before
    the log-softmax; the score is the sum of the target log-probabilities over the answer positions.  Returns [B, n_answers]."""
    V = cfg.vocab_size
    enc = encoder_forward(sd, cfg, net_input["src_tokens"], net_input.get("patch_images"), net_input.get("patch_masks"),
                          training=False)
    masks = candidate_masks(answers, trie, V)
    eos = torch.tensor([EOS])
    result = []
    for c0 in range(0, len(answers), valid_batch_size):                                         # vqa_gen.py:181-183
        ans, msk = answers[c0:c0 + valid_batch_size], masks[c0:c0 + valid_batch_size]
        C = len(ans)
        tgt = _collate([torch.cat([torch.tensor(dp[1:], dtype=torch.long), a, eos]) for dp in decoder_prompts for a in ans], PAD)
        prev = _collate([torch.cat([torch.tensor(dp, dtype=torch.long), a]) for dp in decoder_prompts for a in ans], PAD)
        cm = _collate([torch.cat([torch.zeros(len(dp) - 1, V, dtype=torch.bool), m]) for dp in decoder_prompts for m in msk], False)
        rep = {"encoder_out": enc["encoder_out"].repeat_interleave(C, 0),                      # :192-201
               "encoder_padding_mask": enc["encoder_padding_mask"].repeat_interleave(C, 0),
               "position_embeddings": enc["position_embeddings"].repeat_interleave(C, 0)}
        logits = decoder_forward(sd, cfg, prev, rep)
        logits = logits.masked_fill(~cm, -math.inf)                                             # :204
        lprobs = F.log_softmax(logits.float(), dim=-1)
        sc = lprobs.gather(-1, tgt.unsqueeze(-1)).squeeze(-1)
        sc = sc.masked_fill(tgt.eq(PAD), 0)                                                     # :207
        sc = sc.masked_fill((~cm).all(2), 0)                                                    # :208
        result.append(sc.sum(1).view(-1, C))
    return torch.cat(result, dim=-1)


def _finalize(step, bbsz_idx, eos_scores, tokens, scores, finalized, finished, beam, max_len, lenpen):
    # finalize_hypos  models/sequence_generator.py:637-746
    tokens_clone = tokens.index_select(0, bbsz_idx)[:, 1:step + 2].clone()
    tokens_clone[:, step] = EOS
    pos_scores = scores.index_select(0, bbsz_idx)[:, :step + 1].clone()
    pos_scores[:, step] = eos_scores
    pos_scores[:, 1:] = pos_scores[:, 1:] - pos_scores[:, :-1]
    eos_scores = eos_scores / (step + 1) ** lenpen                                              # :682-683
    cum_unfin, prev = [], 0
    for f in finished:
        if f:
            prev += 1
        else:
            cum_unfin.append(prev)
    cum = torch.tensor(cum_unfin, dtype=torch.long)
    unfin_idx = bbsz_idx // beam
    sent = unfin_idx + cum.index_select(0, unfin_idx)
    seen = sorted(set(zip(sent.tolist(), unfin_idx.tolist())))
    for i in range(bbsz_idx.numel()):
        s = int(sent[i])
        if len(finalized[s]) < beam:
            finalized[s].append({"tokens": tokens_clone[i], "score": eos_scores[i],
                                 "positional_scores": pos_scores[i]})
    newly = []
    for s, u in seen:
        if not finished[s] and (len(finalized[s]) == beam or step == max_len):                  # :748-764
            finished[s] = True
            newly.append(u)
    return newly


def _ngram_block(tokens, lprobs, step, n):
    # fairseq NGramRepeatBlock (un-vendored) python path: ban tokens completing an already-seen n-gram
    for r in range(tokens.size(0)):
        gen = tokens[r, :step + 1].tolist()
        if step + 2 - n < 0:
            continue
        key = tuple(gen[step + 2 - n: step + 1])
        for i in range(len(gen) - n + 1):
            if tuple(gen[i:i + n - 1]) == key:
                lprobs[r, gen[i + n - 1]] = -math.inf
    return lprobs
