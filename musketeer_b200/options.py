"""Command-line surface of the `ofa` model: every flag the reference registers in `TransformerModel.add_args`
(models/ofa/unify_transformer.py:150-334) and `OFAModel.add_args` (models/ofa/ofa.py:45-73), as one table, so that
`train.py --arch ofa_base --scale-attn ...` parses against this package exactly as against the reference
(run_scripts/musketeer/train_musketeer.sh:124-176).  Flags whose feature is outside the hot-path scope still PARSE; the
model constructor raises NotImplementedError when one of them is switched on (musketeer_b200/ofa.py::_unsupported)."""

ACTIVATIONS = ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]     # fairseq.utils.get_available_activation_fns
DEFAULT_MIN_PARAMS_TO_WRAP = int(1e8)

F, I, S, B = "float", "int", "str", "flag"

# (option strings, kind, extra argparse keywords)
MODEL_FLAGS = [
    # unify_transformer.py:150-334
    (("--activation-fn",), S, {"choices": ACTIVATIONS}),
    (("--dropout",), F, {}), (("--attention-dropout",), F, {}), (("--activation-dropout", "--relu-dropout"), F, {}),
    (("--encoder-embed-path",), S, {}), (("--encoder-embed-dim",), I, {}), (("--encoder-ffn-embed-dim",), I, {}),
    (("--encoder-layers",), I, {}), (("--encoder-attention-heads",), I, {}), (("--encoder-normalize-before",), B, {}),
    (("--encoder-learned-pos",), B, {}), (("--bitfit",), B, {"default": False}), (("--freeze-encoder",), B, {}),
    (("--adapter",), B, {}), (("--adapter-dim",), I, {}),
    (("--encoder-prompt",), B, {}), (("--encoder-prompt-type",), S, {"choices": ["prefix"]}),
    (("--encoder-prompt-projection",), B, {}), (("--encoder-prompt-length",), I, {}), (("--encoder-prompt-dim",), I, {}),
    (("--decoder-embed-path",), S, {}), (("--decoder-embed-dim",), I, {}), (("--decoder-ffn-embed-dim",), I, {}),
    (("--decoder-layers",), I, {}), (("--decoder-attention-heads",), I, {}), (("--decoder-learned-pos",), B, {}),
    (("--decoder-normalize-before",), B, {}), (("--decoder-output-dim",), I, {}), (("--freeze-decoder",), B, {}),
    (("--decoder-prompt",), B, {}), (("--decoder-prompt-type",), S, {"choices": ["prefix"]}),
    (("--decoder-prompt-length",), I, {}), (("--decoder-prompt-projection",), B, {}), (("--decoder-prompt-dim",), I, {}),
    (("--share-decoder-input-output-embed",), B, {}), (("--share-all-embeddings",), B, {}),
    (("--no-token-positional-embeddings",), B, {"default": False}),
    (("--adaptive-softmax-cutoff",), S, {}), (("--adaptive-softmax-dropout",), F, {}),
    (("--layernorm-embedding",), B, {}), (("--no-scale-embedding",), B, {}),
    (("--checkpoint-activations",), B, {}), (("--offload-activations",), B, {}),
    (("--no-cross-attention",), B, {"default": False}), (("--cross-self-attention",), B, {"default": False}),
    (("--encoder-layerdrop",), F, {"default": 0}), (("--decoder-layerdrop",), F, {"default": 0}),
    (("--encoder-layers-to-keep",), S, {"default": None}), (("--decoder-layers-to-keep",), S, {"default": None}),
    (("--quant-noise-pq",), F, {"default": 0}), (("--quant-noise-pq-block-size",), I, {"default": 8}),
    (("--quant-noise-scalar",), F, {"default": 0}),
    (("--min-params-to-wrap",), I, {"default": DEFAULT_MIN_PARAMS_TO_WRAP}),
    (("--resnet-drop-path-rate",), F, {}), (("--encoder-drop-path-rate",), F, {}), (("--decoder-drop-path-rate",), F, {}),
    (("--token-bucket-size",), I, {}), (("--image-bucket-size",), I, {}), (("--attn-scale-factor",), F, {}),
    (("--freeze-resnet",), B, {}), (("--freeze-encoder-embedding",), B, {}), (("--freeze-decoder-embedding",), B, {}),
    (("--add-type-embedding",), B, {}), (("--interpolate-position",), B, {}),
    (("--resnet-type",), S, {"choices": ["resnet50", "resnet101", "resnet152"]}), (("--resnet-model-path",), S, {}),
    (("--code-image-size",), I, {}), (("--patch-layernorm-embedding",), B, {}), (("--code-layernorm-embedding",), B, {}),
    (("--entangle-position-embedding",), B, {}), (("--disable-entangle",), B, {}), (("--sync-bn",), B, {}),
    (("--scale-attn",), B, {}), (("--scale-fc",), B, {}), (("--scale-heads",), B, {}), (("--scale-resids",), B, {}),
    # ofa.py:45-73
    (("--pooler-dropout",), F, {}), (("--pooler-classifier",), S, {"choices": ["mlp", "linear"]}),
    (("--pooler-activation-fn",), S, {"choices": ACTIVATIONS}), (("--spectral-norm-classification-head",), B, {}),
    (("--regression_head",), B, {}),
]

# criterions/label_smoothed_cross_entropy.py:14-53 (a fairseq dataclass there; fairseq derives the flags from it)
CRITERION_FLAGS = [
    (("--label-smoothing",), F, {"default": 0.0}), (("--report-accuracy",), B, {"default": False}),
    (("--ignore-prefix-size",), I, {"default": 0}), (("--ignore-eos",), B, {"default": False}),
    (("--drop-worst-ratio",), F, {"default": 0.0}), (("--drop-worst-after",), I, {"default": 0}),
    (("--use-rdrop",), B, {"default": False}), (("--reg-alpha",), F, {"default": 1.0}),
    (("--sample-patch-num",), I, {"default": 196}), (("--constraint-range",), S, {"default": None}),
]


def _add(parser, table):
    kinds = {F: float, I: int, S: str}
    for names, kind, extra in table:
        if kind == B:
            parser.add_argument(*names, action="store_true", **extra)
        else:
            parser.add_argument(*names, type=kinds[kind], **extra)


def add_model_args(parser):
    _add(parser, MODEL_FLAGS)


def add_criterion_args(parser):
    _add(parser, CRITERION_FLAGS)
