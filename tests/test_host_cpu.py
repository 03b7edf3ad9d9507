"""CPU-side checks: the C-ABI library exports what include/ofa_b200.h declares, the product's host logic
(state-dict contract, integer tables, arch presets, loud failure without CUDA) -- no kernel is launched here."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import synth
from tests.helpers import GOLDEN, build_product

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from musketeer_b200 import build
    return build.build()


def test_abi_exports_every_declared_symbol(lib_path):
    hdr = open(os.path.join(ROOT, "include", "ofa_b200.h")).read()
    declared = set(re.findall(r"\b(ofa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ofa_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    from musketeer_b200 import _lib
    lib = _lib.load()
    assert set(_lib.SIGNATURES) | {"ofa_last_error"} == declared
    assert lib.ofa_abi_version() == 1


def test_sass_is_blackwell_native(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA tensor loads
    assert "UTMASTG" in sass      # TMA tensor stores (GEMM / conv epilogues)
    assert "UTMAREDG" in sass     # TMA reduce-add (fp32 gradient arenas, attention dQ)
    assert "LDTM" in sass         # tcgen05.ld
    # no legacy mma.sync on the GEMM-shaped paths: the warp-level MMA users are the two attention kernels whose query side
    # cannot fill a 128-row tcgen05 tile -- the single-token decode attention (G <= 8 query rows per sentence, csrc/decode.cu)
    # and the forward / backward for T <= 16 target rows (csrc/attention_small.cu); all are streams over K / V
    for fn in sass.split("Function :")[1:]:
        if "HMMA." in fn.replace("UTCHMMA", ""):
            name = fn.split("\n", 1)[0]
            assert any(k in name for k in ("attn_decode_mma_kernel", "attn_decode_online_kernel", "attn_bwd_smallq_kernel",
                                                 "attn_fwd_smallq_kernel")), name
    assert "UBLKCP" in sass       # 1-D bulk copies (the LayerNorm backward's shared-memory row ring)
    assert "FFMA2" in sass        # packed fp32x2 arithmetic (wide-row LayerNorm / GELU forward)
    # relative-position table gradients of the attention backward: native int32 shared atomics, no float CAS loops
    bwd = sass[sass.index("attn_bwd_tc_kernel"):]
    bwd = bwd[:bwd.index("Function :", 10)] if "Function :" in bwd[10:] else bwd
    assert "ATOMS.ADD" in bwd and "ATOMS.CAST" not in bwd


@pytest.mark.parametrize("arch", ["ofa_tiny", "ofa_base"])
def test_state_dict_contract(arch):
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))[arch]
    cfg = synth.make_cfg(arch)
    from types import SimpleNamespace
    from musketeer_b200 import OFAModel
    from tests.helpers import FakeTask
    args = SimpleNamespace(**vars(cfg))
    args.no_scale_embedding = True
    model = OFAModel.build_model(args, FakeTask(cfg.vocab_size))
    msd = model.state_dict()
    assert list(msd.keys()) == [e[0] for e in spec["entries"]]
    for (k, shape, dt) in spec["entries"]:
        assert list(msd[k].shape) == shape, k
        assert str(msd[k].dtype) == dt, k
    assert sum(p.numel() for p in model.parameters()) == spec["n_params"]
    assert model.decoder.output_projection.weight is model.encoder.embed_tokens.weight
    assert model.decoder.embed_tokens.weight is model.encoder.embed_tokens.weight


def test_bucket_tables_bit_exact():
    from musketeer_b200.ofa import make_token_bucket_position, make_image_bucket_position, _rel_bucket_1d
    t = make_token_bucket_position(256)
    assert torch.equal(t, synth.token_bucket_table(256))
    b = make_image_bucket_position(42, 83 * 83 + 3)
    assert torch.equal(b, synth.image_bucket_table(42, 83 * 83 + 3))
    r = _rel_bucket_1d(t)
    i = torch.arange(1024)[:, None]
    j = torch.arange(1024)[None, :]
    assert torch.equal(r[(i - j) + 1023], t)      # the bucket is a function of i - j only


def test_arch_presets():
    from types import SimpleNamespace
    from musketeer_b200 import ARCHS
    want = {"ofa_tiny": (256, 4, 4, 4, "resnet50"), "ofa_medium": (512, 4, 4, 8, "resnet101"),
            "ofa_base": (768, 6, 6, 12, "resnet101"), "ofa_large": (1024, 12, 12, 16, "resnet152"),
            "ofa_huge": (1280, 24, 12, 16, "resnet152")}
    for name, (d, le, ld, h, rn) in want.items():
        a = SimpleNamespace()
        ARCHS[name](a)
        assert (a.encoder_embed_dim, a.encoder_layers, a.decoder_layers, a.encoder_attention_heads, a.resnet_type) == \
            (d, le, ld, h, rn)
        assert a.encoder_ffn_embed_dim == 4 * d and a.share_all_embeddings and a.attn_scale_factor == 2


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing without its kernels."""
    from musketeer_b200 import ops, _lib
    x = torch.randn(4, 64)
    with pytest.raises(_lib.OfaKernelError):
        ops.layer_norm(x, torch.ones(64), torch.zeros(64))
    with pytest.raises(_lib.OfaKernelError):
        ops.linear(x, torch.randn(8, 64))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "musketeer_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f


def test_grad_accumulator_keeps_running_sum_over_micro_steps():
    """update_freq > 1 (trainer.py:752-773): a second micro-step that runs without clearing .grad must ADD to the first one's
    gradients although .grad aliases the accumulator's arenas (the arenas were overwritten and .grad doubled).
    The producers are emulated on the CPU by writing into the accumulator's targets the way the kernels do."""
    import torch
    from musketeer_b200 import ops
    lin = torch.nn.Linear(64, 64)            # weight -> fp32 arena (GEMM reduce-add); bias -> bf16/param-dtype buffer
    emb = torch.nn.Embedding(16, 64)         # both: GEMM part (tied projection) + scatter-add part
    model = torch.nn.ModuleList([lin, emb])

    def micro_step(c):
        with ops.grad_accumulation(model) as acc:
            acc.target32(lin.weight).add_(c)                         # TMA reduce-add of a wgrad GEMM
            b, accum = acc.target(lin.bias)                          # colsum kernel: overwrite or accumulate
            b.copy_(b + 2 * c if accum else torch.full_like(b, 2 * c))
            acc.target32(emb.weight).add_(3 * c)
            e, accum = acc.target(emb.weight)
            if not accum:
                e.zero_()
            e.add_(4 * c)

    micro_step(1.0)
    assert float(lin.weight.grad[0, 0]) == 1 and float(lin.bias.grad[0]) == 2 and float(emb.weight.grad[0, 0]) == 7
    micro_step(3.0)                                                  # no zero_grad in between
    assert float(lin.weight.grad[0, 0]) == 4, float(lin.weight.grad[0, 0])
    assert float(lin.bias.grad[0]) == 8 and float(emb.weight.grad[0, 0]) == 28
    with ops.grad_accumulation(model) as acc:                        # a step that touches only one parameter keeps the others
        acc.target32(lin.weight).add_(1.0)
    assert float(lin.weight.grad[0, 0]) == 5 and float(lin.bias.grad[0]) == 8 and float(emb.weight.grad[0, 0]) == 28
    for p in model.parameters():
        p.grad = None
    micro_step(2.0)                                                  # cleared gradients start from zero again
    assert float(lin.weight.grad[0, 0]) == 2 and float(lin.bias.grad[0]) == 4 and float(emb.weight.grad[0, 0]) == 14
    lin.weight.grad = torch.ones_like(lin.weight)                    # a foreign gradient tensor is added to, not replaced
    micro_step(1.0)
    assert float(lin.weight.grad[0, 0]) == 2


def test_flatten_trie_matches_the_trie_walk():
    """Host side of constrained decoding: the CSR form of utils/trie.py's Trie (root = node 0, children in
    insertion order) answers get_next_layer for every prefix exactly like the object trie; dead prefixes have no node."""
    from musketeer_b200.sequence_generator import flatten_trie
    from oracle import ofa_oracle as oo, synth
    trie = oo.Trie(2)
    words = synth.trie_words(n=40, max_len=4, seed=3, vocab=4099)
    for w in words:
        trie.insert(w)
    (ptr, tok, child), index = flatten_trie(trie, "cpu", return_index=True)
    ptr, tok, child = ptr.tolist(), tok.tolist(), child.tolist()
    assert ptr[0] == 0 and len(ptr) == max(child) + 2 and len(tok) == len(child) == ptr[-1]
    assert len(index) == len(ptr) - 1 and sorted(index.values()) == list(range(len(ptr) - 1))

    def walk(prefix):
        node = 0
        for t in prefix:
            nxt = [child[e] for e in range(ptr[node], ptr[node + 1]) if tok[e] == t]
            if not nxt:
                return None
            node = nxt[0]
        return node

    seen = set()
    for w in words:
        for n in range(1, len(w) + 1):
            pre = tuple(w[:n])
            if pre in seen:
                continue
            seen.add(pre)
            node = walk(pre)
            assert node is not None
            assert [tok[e] for e in range(ptr[node], ptr[node + 1])] == list(trie.get_next_layer(list(pre)))
    assert walk([0, 4098, 4097]) is None                    # a prefix outside the trie: the kernels fall back to [eos]
    assert trie.get_next_layer([0, 4098, 4097]) == [2]


def test_generator_argument_checks():
    """Options the hot path does not carry fail at the call, naming the option (no silent fallback)."""
    from musketeer_b200.sequence_generator import SequenceGenerator
    from oracle import synth
    from tests.helpers import build_product
    cfg = synth.make_cfg("ofa_micro", vocab_size=4099)
    model, task = build_product(cfg, synth.synth_state_dict(cfg, seed=0), device="cpu")
    with pytest.raises(NotImplementedError):
        SequenceGenerator([model], task.target_dictionary, beam_size=9)
    with pytest.raises(NotImplementedError):
        SequenceGenerator([model, model], task.target_dictionary)
    gen = SequenceGenerator([model], task.target_dictionary, beam_size=2, constraint_range="10,20")
    sample = synth.make_batch(2, 9, 2, img=64, seed=1, vocab=4099)
    with pytest.raises(NotImplementedError):
        gen.generate([model], sample, prefix_tokens=torch.full((2, 1), 5))
    with pytest.raises(NotImplementedError):
        gen.generate([model], sample, constraints=torch.zeros(2, 1))
