"""The fairseq boundary on CPU, with the fairseq / omegaconf stand-ins of oracle/ref_shim on the path
(fairseq itself is un-vendored and not installable here).  Every check runs in a fresh interpreter so that
musketeer_b200 binds to the shim's base classes at import time, exactly as it binds to real fairseq in a training
environment."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "ref_shim")

# run_scripts/musketeer/train_musketeer.sh:124-176, the model- and criterion-level flags with the script's values
SCRIPT_FLAGS = """--arch=ofa_base --criterion=adjust_label_smoothed_cross_entropy --label-smoothing=0.1
--encoder-normalize-before --decoder-normalize-before --share-decoder-input-output-embed --share-all-embeddings
--layernorm-embedding --patch-layernorm-embedding --code-layernorm-embedding --resnet-drop-path-rate=0.0
--encoder-drop-path-rate=0.1 --decoder-drop-path-rate=0.1 --dropout=0.1 --attention-dropout=0.0 --add-type-embedding
--scale-attn --scale-fc --scale-heads --use-rdrop --disable-entangle --code-image-size=320 --sample-patch-num=196
--drop-worst-ratio=0.2 --drop-worst-after=6000""".split()


def _run(body):
    code = "import sys\nsys.path.insert(0, %r)\nsys.path.insert(0, %r)\n" % (SHIM, ROOT) + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_registries_and_base_classes():
    _run("""
        import fairseq.models as fm, fairseq.criterions as fcr
        import musketeer_b200.plugin as plugin
        from musketeer_b200 import OFAModel, AdjustLabelSmoothedCrossEntropyCriterion
        from musketeer_b200.ofa import TransformerEncoder, TransformerDecoder
        assert plugin.REGISTERED
        assert fm.MODEL_REGISTRY["ofa"] is OFAModel
        for arch in ("ofa_tiny", "ofa_medium", "ofa_base", "ofa_large", "ofa_huge"):
            assert fm.ARCH_MODEL_REGISTRY[arch] is OFAModel and callable(fm.ARCH_CONFIG_REGISTRY[arch]), arch
        assert fcr.CRITERION_REGISTRY["adjust_label_smoothed_cross_entropy"] is AdjustLabelSmoothedCrossEntropyCriterion
        assert issubclass(OFAModel, fm.FairseqEncoderDecoderModel)
        assert issubclass(TransformerEncoder, fm.FairseqEncoder)
        assert issubclass(TransformerDecoder, fm.FairseqIncrementalDecoder)     # sequence_generator.py:776-781 gates on this
        assert issubclass(AdjustLabelSmoothedCrossEntropyCriterion, fcr.FairseqCriterion)
        assert callable(AdjustLabelSmoothedCrossEntropyCriterion.reduce_metrics)
        assert AdjustLabelSmoothedCrossEntropyCriterion.logging_outputs_can_be_summed()
    """)


def test_script_flags_parse_and_model_builds():
    """argparse round trip of the flag list of train_musketeer.sh through OFAModel.add_args + the criterion's flags, the arch
    function, build_model, and the criterion constructor -- the path train.py walks (train.py:95-112)."""
    out = _run("""
        import argparse
        import fairseq.models as fm
        import musketeer_b200.plugin
        from musketeer_b200.options import add_criterion_args
        from musketeer_b200.synthetic import Task
        flags = %r
        p = argparse.ArgumentParser(allow_abbrev=False)
        p.add_argument("--arch"); p.add_argument("--criterion")
        g = p.add_argument_group("Model-specific configuration", argument_default=argparse.SUPPRESS)   # fairseq options.py idiom
        fm.ARCH_MODEL_REGISTRY["ofa_base"].add_args(g)
        add_criterion_args(p)
        args = p.parse_args(flags)
        assert args.scale_attn and args.scale_fc and args.scale_heads and args.disable_entangle and args.use_rdrop
        assert args.dropout == 0.1 and args.encoder_drop_path_rate == 0.1 and args.sample_patch_num == 196
        assert not hasattr(args, "encoder_embed_dim")            # left to the architecture function
        fm.ARCH_CONFIG_REGISTRY[args.arch](args)
        assert (args.encoder_embed_dim, args.encoder_layers, args.decoder_layers, args.resnet_type) == (768, 6, 6, "resnet101")
        # every reference flag is known to the parser
        import re
        known = {s for a in p._actions for s in a.option_strings}
        for f in ["--pooler-dropout", "--regression_head", "--encoder-prompt-dim", "--quant-noise-pq", "--min-params-to-wrap",
                  "--relu-dropout", "--freeze-resnet", "--sync-bn", "--scale-resids", "--resnet-model-path", "--bitfit"]:
            assert f in known, f
        # build (tiny dims keep the CPU test fast; same flag set)
        args.encoder_embed_dim = args.decoder_embed_dim = 256
        args.encoder_ffn_embed_dim = args.decoder_ffn_embed_dim = 512
        args.encoder_attention_heads = args.decoder_attention_heads = 4
        args.encoder_layers = args.decoder_layers = 2
        args.resnet_type = "resnet50"
        args.patch_image_size = 384
        task = Task(4099)
        model = fm.ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)
        assert isinstance(model, fm.FairseqEncoderDecoderModel) and isinstance(model.decoder, fm.FairseqIncrementalDecoder)
        assert model.encoder.layers[1].drop_path_rate > 0 and model.encoder.layers[0].dropout_p == 0.1
        assert model.decoder.output_projection.weight is model.encoder.embed_tokens.weight
        from musketeer_b200 import AdjustLabelSmoothedCrossEntropyCriterion as C
        crit = C(task, False, args.label_smoothing, use_rdrop=args.use_rdrop, drop_worst_ratio=args.drop_worst_ratio,
                 drop_worst_after=args.drop_worst_after, sample_patch_num=args.sample_patch_num)
        assert crit.padding_idx == 1 and crit.task is task
        h = model.half()                      # --fp16 (train_musketeer.sh:174): runs as bf16
        import torch
        assert next(h.parameters()).dtype == torch.bfloat16
        print("ok", sum(p.numel() for p in model.parameters()))
    """ % (SCRIPT_FLAGS,))
    assert out.startswith("ok")


def test_upgrade_state_dict_named():
    """Checkpoint compatibility (models/ofa/ofa.py:216-318, unify_transformer.py:1033-1072,1605-1659, layer / attention
    upgrades): vocabulary growth, dropped classification heads, missing buffers, legacy layer_norms.* and in_proj_* keys."""
    _run("""
        import torch
        import musketeer_b200.plugin
        from musketeer_b200.synthetic import build_model, Task
        m, task = build_model("ofa_tiny", "cpu", torch.float32, vocab=300, encoder_layers=1, decoder_layers=1)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        # a checkpoint written with a smaller vocabulary and without the bucket buffers, plus a stray head and legacy keys
        for k in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "decoder.output_projection.weight"):
            sd[k] = sd[k][:290]
        for k in ("encoder.token_rp_bucket", "decoder.image_rp_bucket", "decoder.image_position_idx"):
            del sd[k]
        sd["classification_heads.snli.dense.weight"] = torch.zeros(4, 4)
        sd["classification_heads.snli.out_proj.weight"] = torch.zeros(3, 4)
        sd["encoder.layers.0.layer_norms.0.weight"] = sd.pop("encoder.layers.0.self_attn_layer_norm.weight")
        sd["encoder.layers.0.layer_norms.0.bias"] = sd.pop("encoder.layers.0.self_attn_layer_norm.bias")
        a = "decoder.layers.0.self_attn."
        sd[a + "in_proj_weight"] = torch.cat([sd.pop(a + "q_proj.weight"), sd.pop(a + "k_proj.weight"), sd.pop(a + "v_proj.weight")])
        sd[a + "in_proj_bias"] = torch.cat([sd.pop(a + "q_proj.bias"), sd.pop(a + "k_proj.bias"), sd.pop(a + "v_proj.bias")])
        small = sd["encoder.embed_image_positions.weight"]
        sd["encoder.embed_image_positions.weight"] = small[:100]
        ref = m.state_dict()
        m.upgrade_state_dict_named(sd, "")
        assert set(sd.keys()) == set(ref.keys()), set(sd.keys()) ^ set(ref.keys())
        m.load_state_dict(sd, strict=True)
        assert sd["encoder.embed_tokens.weight"].shape[0] == 300
        assert torch.equal(sd["encoder.embed_tokens.weight"][:290], ref["encoder.embed_tokens.weight"][:290])
        assert torch.equal(sd[a + "k_proj.weight"], ref[a + "k_proj.weight"])
        assert torch.equal(sd["encoder.layers.0.self_attn_layer_norm.weight"], ref["encoder.layers.0.self_attn_layer_norm.weight"])
        assert sd["encoder.embed_image_positions.weight"].shape == ref["encoder.embed_image_positions.weight"].shape
    """)


def test_ddp_wrapper_is_installed_behind_fairseq_models():
    """trainer.py:254-266 calls fairseq.models.DistributedFairseqModel(cfg, model, process_group, device) at run time: with the
    plugin imported that name hands this package's OFAModel to DistributedOFAModel for --ddp-backend=no_c10d."""
    _run("""
        import os, types, torch, torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
        dist.init_process_group("gloo", rank=0, world_size=1)
        import fairseq.models as fm
        import musketeer_b200.plugin
        from musketeer_b200.dp import DistributedOFAModel
        from musketeer_b200.synthetic import build_model
        m, task = build_model("ofa_tiny", "cpu", torch.float32, vocab=300, encoder_layers=1, decoder_layers=1)
        w = fm.DistributedFairseqModel(types.SimpleNamespace(ddp_backend="no_c10d"), m, process_group=None, device="cpu")
        assert isinstance(w, DistributedOFAModel) and w.module is m
        assert hasattr(w, "no_sync") and hasattr(w, "all_reduce_grads")
        assert w.encoder is m.encoder and w.max_decoder_positions() == m.max_decoder_positions()
        assert list(w.state_dict().keys()) == list(m.state_dict().keys())
        try:
            fm.DistributedFairseqModel(types.SimpleNamespace(ddp_backend="pytorch_ddp"), torch.nn.Linear(2, 2), None, "cpu")
            raise SystemExit("other backends must be delegated to fairseq")
        except ValueError:
            pass
        dist.destroy_process_group()
    """)
