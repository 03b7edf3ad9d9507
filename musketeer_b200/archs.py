"""Architecture presets: same names and defaults as the reference (models/ofa/ofa.py:370-486,
models/ofa/unify_transformer.py:1680-1745).  Each fills only attributes the caller left unset."""


def _d(args, name, value):
    setattr(args, name, getattr(args, name, value))


def base_architecture(args):
    _d(args, "encoder_embed_path", None)
    _d(args, "encoder_embed_dim", 512)
    _d(args, "encoder_ffn_embed_dim", 2048)
    _d(args, "encoder_layers", 6)
    _d(args, "encoder_attention_heads", 8)
    _d(args, "encoder_normalize_before", False)
    _d(args, "encoder_learned_pos", False)
    _d(args, "decoder_embed_path", None)
    _d(args, "decoder_embed_dim", args.encoder_embed_dim)
    _d(args, "decoder_ffn_embed_dim", args.encoder_ffn_embed_dim)
    _d(args, "decoder_layers", 6)
    _d(args, "decoder_attention_heads", 8)
    _d(args, "decoder_normalize_before", False)
    _d(args, "decoder_learned_pos", False)
    _d(args, "attention_dropout", 0.0)
    _d(args, "activation_dropout", 0.0)
    _d(args, "activation_fn", "relu")
    _d(args, "dropout", 0.1)
    _d(args, "adaptive_softmax_cutoff", None)
    _d(args, "adaptive_softmax_dropout", 0)
    _d(args, "share_decoder_input_output_embed", False)
    _d(args, "share_all_embeddings", False)
    _d(args, "no_token_positional_embeddings", False)
    _d(args, "adaptive_input", False)
    _d(args, "no_cross_attention", False)
    _d(args, "cross_self_attention", False)
    _d(args, "decoder_output_dim", args.decoder_embed_dim)
    _d(args, "decoder_input_dim", args.decoder_embed_dim)
    _d(args, "encoder_prompt", False)
    _d(args, "decoder_prompt", False)
    _d(args, "no_scale_embedding", False)
    _d(args, "layernorm_embedding", False)
    _d(args, "tie_adaptive_weights", False)
    _d(args, "checkpoint_activations", False)
    _d(args, "offload_activations", False)
    _d(args, "encoder_layers_to_keep", None)
    _d(args, "decoder_layers_to_keep", None)
    _d(args, "encoder_layerdrop", 0)
    _d(args, "decoder_layerdrop", 0)
    _d(args, "quant_noise_pq", 0)
    _d(args, "quant_noise_pq_block_size", 8)
    _d(args, "quant_noise_scalar", 0)


def ofa_large_architecture(args):
    _d(args, "encoder_embed_path", None)
    _d(args, "encoder_embed_dim", 1024)
    _d(args, "encoder_ffn_embed_dim", 4 * 1024)
    _d(args, "encoder_layers", 12)
    _d(args, "encoder_attention_heads", 16)
    _d(args, "encoder_normalize_before", True)
    _d(args, "encoder_learned_pos", True)
    _d(args, "decoder_embed_path", None)
    _d(args, "decoder_embed_dim", args.encoder_embed_dim)
    _d(args, "decoder_ffn_embed_dim", args.encoder_ffn_embed_dim)
    _d(args, "decoder_layers", 12)
    _d(args, "decoder_attention_heads", 16)
    _d(args, "decoder_normalize_before", True)
    _d(args, "decoder_learned_pos", True)
    _d(args, "attention_dropout", 0.0)
    _d(args, "relu_dropout", 0.0)
    _d(args, "dropout", 0.0)
    _d(args, "max_target_positions", 1024)
    _d(args, "max_source_positions", 1024)
    _d(args, "adaptive_softmax_cutoff", None)
    _d(args, "adaptive_softmax_dropout", 0)
    _d(args, "share_decoder_input_output_embed", True)
    _d(args, "share_all_embeddings", True)
    _d(args, "decoder_output_dim", args.decoder_embed_dim)
    _d(args, "decoder_input_dim", args.decoder_embed_dim)
    _d(args, "no_scale_embedding", True)
    _d(args, "layernorm_embedding", True)
    _d(args, "activation_fn", "gelu")
    _d(args, "pooler_activation_fn", "tanh")
    _d(args, "pooler_dropout", 0.0)
    _d(args, "pooler_classifier", "mlp")
    _d(args, "resnet_drop_path_rate", 0.0)
    _d(args, "encoder_drop_path_rate", 0.0)
    _d(args, "decoder_drop_path_rate", 0.0)
    _d(args, "resnet_type", "resnet152")
    _d(args, "token_bucket_size", 256)
    _d(args, "image_bucket_size", 42)
    _d(args, "freeze_encoder_embedding", False)
    _d(args, "freeze_decoder_embedding", False)
    _d(args, "add_type_embedding", True)
    _d(args, "attn_scale_factor", 2)
    _d(args, "code_image_size", 128)
    _d(args, "patch_layernorm_embedding", True)
    _d(args, "code_layernorm_embedding", True)
    _d(args, "entangle_position_embedding", False)
    _d(args, "disable_entangle", False)
    _d(args, "sync_bn", False)
    _d(args, "scale_attn", False)
    _d(args, "scale_fc", False)
    _d(args, "scale_heads", False)
    _d(args, "scale_resids", False)
    _d(args, "orig_patch_image_size", 256)


def _sized(dim, enc_layers, dec_layers, heads, resnet):
    def arch(args):
        _d(args, "encoder_embed_dim", dim)
        _d(args, "encoder_ffn_embed_dim", 4 * dim)
        _d(args, "encoder_layers", enc_layers)
        _d(args, "encoder_attention_heads", heads)
        _d(args, "decoder_layers", dec_layers)
        _d(args, "decoder_attention_heads", heads)
        _d(args, "resnet_type", resnet)
        ofa_large_architecture(args)
    return arch


ofa_base_architecture = _sized(768, 6, 6, 12, "resnet101")
ofa_huge_architecture = _sized(1280, 24, 12, 16, "resnet152")
ofa_medium_architecture = _sized(512, 4, 4, 8, "resnet101")
ofa_tiny_architecture = _sized(256, 4, 4, 4, "resnet50")

ARCHS = {
    "ofa_large": ofa_large_architecture,
    "ofa_base": ofa_base_architecture,
    "ofa_huge": ofa_huge_architecture,
    "ofa_medium": ofa_medium_architecture,
    "ofa_tiny": ofa_tiny_architecture,
}
