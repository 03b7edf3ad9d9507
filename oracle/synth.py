"""Synthetic configs, weights and batches shared by the oracle, the golden generator and the tests.

TEST INFRASTRUCTURE ONLY -- nothing under `musketeer_b200/` imports this file.

* `make_cfg(arch, **over)`  : the reference's arch defaults (models/ofa/ofa.py:370-486,
  models/ofa/unify_transformer.py:1680-1745) plus the Musketeer flag set
  (run_scripts/musketeer/train_musketeer.sh:124-176).
* `state_spec(cfg)`         : name -> (shape, kind) of every state_dict entry of the reference model
  (checked entry-by-entry against the reference in oracle/make_golden.py).
* `synth_state_dict(cfg)`   : deterministic per-name values (independent of iteration order); LN / BN /
  c_attn / rel-pos tables are perturbed away from their init so parity exercises them.
* `make_batch(...)`         : seeded synthetic batches in the reference's `sample` layout.
"""
import hashlib
import math
from collections import OrderedDict
from types import SimpleNamespace

import torch

VOCAB = 59457  # tasks/ofa_task.py:93-116: dict.txt + <mask> + 8192 <code_i> + 1000 <bin_i>
PAD, BOS, EOS, UNK = 1, 0, 2, 3

_ARCH = {
    # name: (embed, layers_enc, layers_dec, heads, resnet)   models/ofa/ofa.py:370-486
    "ofa_tiny": (256, 4, 4, 4, "resnet50"),
    "ofa_medium": (512, 4, 4, 8, "resnet101"),
    "ofa_base": (768, 6, 6, 12, "resnet101"),
    "ofa_large": (1024, 12, 12, 16, "resnet152"),
    "ofa_huge": (1280, 24, 12, 16, "resnet152"),
    # test-only micro model (not a reference arch): fast on CPU, same code paths
    "ofa_micro": (128, 2, 2, 2, "resnet50"),
}
RESNET_BLOCKS = {"resnet50": [3, 4, 6], "resnet101": [3, 4, 23], "resnet152": [3, 8, 36]}


def make_cfg(arch="ofa_tiny", **over):
    d, le, ld, h, rn = _ARCH[arch]
    cfg = dict(
        arch=arch,
        encoder_embed_dim=d, decoder_embed_dim=d,
        encoder_ffn_embed_dim=4 * d, decoder_ffn_embed_dim=4 * d,
        encoder_layers=le, decoder_layers=ld,
        encoder_attention_heads=h, decoder_attention_heads=h,
        resnet_type=rn,
        vocab_size=VOCAB,
        max_source_positions=1024, max_target_positions=1024,
        token_bucket_size=256, image_bucket_size=42,
        attn_scale_factor=2.0,
        code_image_size=128,
        patch_image_size=384, orig_patch_image_size=256,
        # Musketeer flag set
        scale_attn=True, scale_fc=True, scale_heads=True,
        add_type_embedding=True, disable_entangle=True,
        layernorm_embedding=True, patch_layernorm_embedding=True, code_layernorm_embedding=True,
        share_all_embeddings=True,
        encoder_normalize_before=True, decoder_normalize_before=True,
        entangle_position_embedding=False,
        freeze_resnet=False,
        dropout=0.0, attention_dropout=0.0,
        encoder_drop_path_rate=0.0, decoder_drop_path_rate=0.0, resnet_drop_path_rate=0.0,
    )
    cfg.update(over)
    return SimpleNamespace(**cfg)


def to_ref_args(cfg):
    """argparse.Namespace the reference's build_model expects."""
    a = SimpleNamespace(**vars(cfg))
    a.activation_fn = "gelu"
    a.activation_dropout = 0.0
    a.relu_dropout = 0.0
    a.adaptive_input = False
    a.tie_adaptive_weights = False
    a.no_scale_embedding = True
    a.pooler_activation_fn = "tanh"
    a.pooler_dropout = 0.0
    a.pooler_classifier = "mlp"
    return a


# ----------------------------------------------------------------------------------------------
# state-dict specification
# ----------------------------------------------------------------------------------------------
def _ln(spec, p, n):
    spec[p + ".weight"] = ((n,), "ln_w")
    spec[p + ".bias"] = ((n,), "ln_b")


def _lin(spec, p, n_out, n_in, bias=True):
    spec[p + ".weight"] = ((n_out, n_in), "w")
    if bias:
        spec[p + ".bias"] = ((n_out,), "b")


def _bn(spec, p, c, frozen):
    spec[p + ".weight"] = ((c,), "bn_w")
    spec[p + ".bias"] = ((c,), "bn_b")
    spec[p + ".running_mean"] = ((c,), "bn_rm")
    spec[p + ".running_var"] = ((c,), "bn_rv")
    if not frozen:
        spec[p + ".num_batches_tracked"] = ((), "bn_nbt")


def _conv(spec, p, co, ci, k):
    spec[p + ".weight"] = ((co, ci, k, k), "conv")


def _resnet(spec, p, blocks, frozen):
    # models/ofa/resnet.py:136-209 (3-stage bottleneck net, stride in conv2 i.e. v1.5)
    _conv(spec, p + ".conv1", 64, 3, 7)
    _bn(spec, p + ".bn1", 64, frozen)
    inpl = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256), blocks), start=1):
        for bi in range(n):
            q = "%s.layer%d.%d" % (p, li, bi)
            _conv(spec, q + ".conv1", planes, inpl, 1)
            _bn(spec, q + ".bn1", planes, frozen)
            _conv(spec, q + ".conv2", planes, planes, 3)
            _bn(spec, q + ".bn2", planes, frozen)
            _conv(spec, q + ".conv3", planes * 4, planes, 1)
            _bn(spec, q + ".bn3", planes * 4, frozen)
            if bi == 0:
                _conv(spec, q + ".downsample.0", planes * 4, inpl, 1)
                _bn(spec, q + ".downsample.1", planes * 4, frozen)
            inpl = planes * 4


def _attn(spec, p, d, h, scale_heads):
    if scale_heads:
        spec[p + ".c_attn"] = ((h,), "c_attn")
    for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
        _lin(spec, p + "." + n, d, d)


def state_spec(cfg):
    """Ordered like the reference's state_dict() (registration order)."""
    s = OrderedDict()
    d, V = cfg.encoder_embed_dim, cfg.vocab_size
    H = cfg.encoder_attention_heads
    F_ = cfg.encoder_ffn_embed_dim
    n_img_rel = (2 * cfg.image_bucket_size - 1) ** 2 + 3
    n_tok_rel = 2 * cfg.token_bucket_size - 1
    n_imgpos = cfg.image_bucket_size ** 2 + 1

    e = "encoder"
    s[e + ".version"] = ((1,), "version")
    s[e + ".token_rp_bucket"] = ((1024, 1024), "token_rp_bucket")
    s[e + ".image_rp_bucket"] = ((n_imgpos, n_imgpos), "image_rp_bucket")
    s[e + ".embed_tokens.weight"] = ((V, d), "emb_pad")
    if cfg.layernorm_embedding:
        _ln(s, e + ".layernorm_embedding", d)
    if cfg.add_type_embedding:
        s[e + ".type_embedding.weight"] = ((2, d), "emb")
    _resnet(s, e + ".embed_images", RESNET_BLOCKS[cfg.resnet_type], cfg.freeze_resnet)
    _lin(s, e + ".image_proj", d, 1024)
    if cfg.patch_layernorm_embedding:
        _ln(s, e + ".patch_layernorm_embedding", d)
    s[e + ".embed_positions.weight"] = ((cfg.max_source_positions + 2, d), "emb")
    s[e + ".embed_image_positions.weight"] = ((n_imgpos, d), "emb")
    _ln(s, e + ".pos_ln", d)
    _ln(s, e + ".image_pos_ln", d)
    _lin(s, e + ".pos_q_linear", d, d)
    _lin(s, e + ".pos_k_linear", d, d)
    for i in range(cfg.encoder_layers):
        p = "%s.layers.%d" % (e, i)
        _attn(s, p + ".self_attn", d, H, cfg.scale_heads)
        _ln(s, p + ".self_attn_layer_norm", d)
        _lin(s, p + ".fc1", F_, d)
        _lin(s, p + ".fc2", d, F_)
        if cfg.scale_attn:
            _ln(s, p + ".attn_ln", d)
        if cfg.scale_fc:
            _ln(s, p + ".ffn_layernorm", F_)
        _ln(s, p + ".final_layer_norm", d)
    _ln(s, e + ".layer_norm", d)
    for i in range(cfg.encoder_layers):
        s["%s.token_rel_pos_table_list.%d.weight" % (e, i)] = ((n_tok_rel, H), "rel")
    for i in range(cfg.encoder_layers):
        s["%s.image_rel_pos_table_list.%d.weight" % (e, i)] = ((n_img_rel, H), "rel")

    dd = "decoder"
    s[dd + ".version"] = ((1,), "version")
    s[dd + ".token_rp_bucket"] = ((1024, 1024), "token_rp_bucket")
    s[dd + ".image_rp_bucket"] = ((n_imgpos, n_imgpos), "image_rp_bucket")
    s[dd + ".image_position_idx"] = (((cfg.code_image_size // 8) ** 2 + 1 + 769,), "image_position_idx")
    s[dd + ".embed_tokens.weight"] = ((V, d), "tied:encoder.embed_tokens.weight")
    if cfg.layernorm_embedding:
        _ln(s, dd + ".layernorm_embedding", d)
    s[dd + ".embed_positions.weight"] = ((cfg.max_target_positions + 2, d), "emb")
    s[dd + ".embed_image_positions.weight"] = ((n_imgpos, d), "emb")
    _ln(s, dd + ".pos_ln", d)
    _ln(s, dd + ".image_pos_ln", d)
    for n in ("self_pos_q_linear", "self_pos_k_linear", "cross_pos_q_linear", "cross_pos_k_linear"):
        _lin(s, dd + "." + n, d, d)
    if cfg.code_layernorm_embedding:
        _ln(s, dd + ".code_layernorm_embedding", d)
    Hd = cfg.decoder_attention_heads
    for i in range(cfg.decoder_layers):
        p = "%s.layers.%d" % (dd, i)
        _attn(s, p + ".self_attn", d, Hd, cfg.scale_heads)
        if cfg.scale_attn:
            _ln(s, p + ".self_attn_ln", d)
            _ln(s, p + ".cross_attn_ln", d)
        _ln(s, p + ".self_attn_layer_norm", d)
        _attn(s, p + ".encoder_attn", d, Hd, cfg.scale_heads)
        _ln(s, p + ".encoder_attn_layer_norm", d)
        if cfg.scale_fc:
            _ln(s, p + ".ffn_layernorm", cfg.decoder_ffn_embed_dim)
        _lin(s, p + ".fc1", cfg.decoder_ffn_embed_dim, d)
        _lin(s, p + ".fc2", d, cfg.decoder_ffn_embed_dim)
        _ln(s, p + ".final_layer_norm", d)
    _ln(s, dd + ".layer_norm", d)
    s[dd + ".output_projection.weight"] = ((V, d), "tied:encoder.embed_tokens.weight")
    for i in range(cfg.decoder_layers):
        s["%s.token_rel_pos_table_list.%d.weight" % (dd, i)] = ((n_tok_rel, Hd), "rel")
    for i in range(cfg.decoder_layers):
        s["%s.image_rel_pos_table_list.%d.weight" % (dd, i)] = ((n_img_rel, Hd), "rel")
    return s


# ----------------------------------------------------------------------------------------------
# bucket tables (restated; models/ofa/unify_transformer.py:53-81, :1210-1213)
# ----------------------------------------------------------------------------------------------
def token_bucket_table(bucket_size, max_position=1024):
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


def image_bucket_table(bucket_size, num_rel):
    # closed form of unify_transformer.py:66-81: index (1+r*bs+c) ; value (dr+bs-1)*(2bs-1) + (dc+bs-1)
    n = bucket_size * bucket_size
    r = torch.arange(n) // bucket_size
    c = torch.arange(n) % bucket_size
    dr = r[:, None] - r[None, :] + bucket_size - 1
    dc = c[:, None] - c[None, :] + bucket_size - 1
    t = torch.zeros(n + 1, n + 1, dtype=torch.long)
    t[1:, 1:] = dr * (2 * bucket_size - 1) + dc
    t[0, :] = num_rel - 3
    t[:, 0] = num_rel - 2
    t[0, 0] = num_rel - 1
    return t


def decoder_image_position_idx(code_image_size, image_bucket_size):
    w = code_image_size // 8
    idx = torch.arange(w).unsqueeze(0).expand(w, w) + torch.arange(w).unsqueeze(1) * image_bucket_size + 1
    idx = torch.cat([torch.tensor([0]), idx.reshape(-1)])
    return torch.cat([idx, torch.tensor([1024] * 769)])


# ----------------------------------------------------------------------------------------------
# deterministic weights
# ----------------------------------------------------------------------------------------------
def _gen(name, seed):
    h = hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synth_state_dict(cfg, seed=0, emb_std=0.02, w_std=0.02, rel_std=0.2):
    spec = state_spec(cfg)
    sd = OrderedDict()
    for name, (shape, kind) in spec.items():
        g = _gen(name, seed)
        rn = lambda std=1.0: torch.randn(shape, generator=g, dtype=torch.float32) * std
        if kind == "version":
            t = torch.tensor([3.0])
        elif kind == "token_rp_bucket":
            t = token_bucket_table(cfg.token_bucket_size)
        elif kind == "image_rp_bucket":
            t = image_bucket_table(cfg.image_bucket_size, (2 * cfg.image_bucket_size - 1) ** 2 + 3)
        elif kind == "image_position_idx":
            t = decoder_image_position_idx(cfg.code_image_size, cfg.image_bucket_size)
        elif kind == "emb":
            t = rn(emb_std)
        elif kind == "emb_pad":
            t = rn(emb_std)
            t[PAD].zero_()
        elif kind.startswith("tied:"):
            t = sd[kind[5:]]
        elif kind == "w":
            t = rn(w_std)
        elif kind == "b":
            t = rn(0.02)
        elif kind in ("ln_w", "bn_w", "c_attn"):
            t = 1.0 + rn(0.1)
        elif kind in ("ln_b", "bn_b"):
            t = rn(0.05)
        elif kind == "bn_rm":
            t = rn(0.05)
        elif kind == "bn_rv":
            t = 1.0 + rn(0.1).abs()
        elif kind == "bn_nbt":
            t = torch.tensor(0, dtype=torch.long)
        elif kind == "rel":
            t = rn(rel_std)
        elif kind == "conv":
            co, ci, kh, kw = shape
            t = rn(math.sqrt(2.0 / (co * kh * kw)))  # kaiming-normal fan_out (resnet.py:168-170)
        else:
            raise KeyError(kind)
        sd[name] = t
    return sd


def trie_words(n, max_len, seed, vocab=VOCAB, alphabet=12):
    """Answer candidates for a constraint trie, as the tasks insert them ([bos] + answer + [eos], tasks/mm_tasks/vqa_gen.py:
    158-167): n answers of 1..max_len tokens over a small random alphabet, so prefixes are shared and some answers are
    prefixes of others."""
    g = torch.Generator(device="cpu")
    g.manual_seed(5000 + seed)
    alpha = (torch.randperm(min(50265, vocab) - 4, generator=g)[:alphabet] + 4).tolist()
    words = []
    for _ in range(n):
        L = int(torch.randint(1, max_len + 1, (1,), generator=g))
        words.append([BOS] + [alpha[int(i)] for i in torch.randint(0, alphabet, (L,), generator=g)] + [EOS])
    return words


def prefix_tokens(lens, seed, vocab=VOCAB):
    """Forced decoder prefixes as data/mm_data/vqa_gen_dataset.py:228-231 collates them (the decoder prompt without its bos,
    right-padded with pad): one row per sentence, lens[i] random tokens (0 = no prefix for that sentence)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7000 + seed)
    out = torch.full((len(lens), max(max(lens), 1)), PAD, dtype=torch.long)
    for i, n in enumerate(lens):
        out[i, :n] = torch.randint(4, min(50265, vocab), (n,), generator=g)
    return out


def candidate_answers(n, max_len, seed, vocab=VOCAB, alphabet=12):
    """n DISTINCT answers (token tensors without bos / eos) over a small alphabet, as ans2label_dict's keys would encode."""
    seen, out = set(), []
    for w in trie_words(4 * n, max_len, seed, vocab, alphabet):
        a = tuple(w[1:-1])
        if a not in seen:
            seen.add(a)
            out.append(torch.tensor(a, dtype=torch.long))
        if len(out) == n:
            break
    return out


# ----------------------------------------------------------------------------------------------
# synthetic batches (reference `sample` layout: data/mm_data/*_dataset.py collaters)
# ----------------------------------------------------------------------------------------------
def make_batch(bsz, src_len, tgt_len, img=256, seed=0, vocab=VOCAB, n_pad=3, with_image=True,
               target_prefix_pad=0, constraint=False, dtype=torch.float32):
    """SURVEY.md 8(d) C1 recipe: row 0 full length, later rows right-padded by n_pad*row (capped).

    target_prefix_pad>0 pads the first tokens of `target` (VQA/SNLI-VE `prev_output` prompt style,
    data/mm_data/vqa_gen_dataset.py:169)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    hi = min(50265, vocab)
    src = torch.randint(4, hi, (bsz, src_len), generator=g)
    prev = torch.randint(4, hi, (bsz, tgt_len), generator=g)
    tgt = torch.randint(4, hi, (bsz, tgt_len), generator=g)
    src[:, 0] = BOS
    prev[:, 0] = BOS
    lens = []
    for b in range(bsz):
        p = min(n_pad * b, src_len - 3)
        q = min(n_pad * b, tgt_len - 2)
        if p > 0:
            src[b, src_len - p:] = PAD
        src[b, src_len - p - 1] = EOS
        if q > 0:
            prev[b, tgt_len - q:] = PAD
            tgt[b, tgt_len - q:] = PAD
        tgt[b, tgt_len - q - 1] = EOS
        lens.append(src_len - p)
    if target_prefix_pad > 0:
        tgt[:, :target_prefix_pad] = PAD
    sample = {
        "nsentences": bsz,
        "ntokens": int(tgt.ne(PAD).sum()),
        "net_input": {
            "src_tokens": src,
            "src_lengths": torch.tensor(lens, dtype=torch.long),
            "prev_output_tokens": prev,
        },
        "target": tgt,
    }
    if with_image:
        sample["net_input"]["patch_images"] = torch.randn(bsz, 3, img, img, generator=g).to(dtype)
        sample["net_input"]["patch_masks"] = torch.ones(bsz, dtype=torch.bool)
    # text-only tasks (gigaword) carry no patch_images at all (data/nlg_data/summary_dataset.py:52-61)
    if constraint:
        cm = torch.zeros(bsz, tgt_len, vocab, dtype=torch.bool)
        allowed = torch.randint(4, hi, (64,), generator=g)
        cm[:, :, allowed] = True
        cm.scatter_(2, tgt.clamp(min=0).unsqueeze(-1), True)  # the gold token is always allowed
        cm[:, :, EOS] = True
        sample["constraint_masks"] = cm
    return sample
