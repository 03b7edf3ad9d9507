"""The fairseq base classes and registries the OFA plugin binds to.

With fairseq importable (a real training environment, or the stand-in package the tests put on sys.path) the names below
ARE fairseq's: `OFAModel` is then a `FairseqEncoderDecoderModel`, its decoder a `FairseqIncrementalDecoder` (the reference
generator gates incremental decoding on that: models/sequence_generator.py:776-781), the criterion a `FairseqCriterion`
(real `register_criterion` rejects anything else), and `register_model("ofa")` / `register_model_architecture` /
`register_criterion` write into fairseq's registries.

fairseq is an un-vendored, un-installable dependency of the reference (it is absent on the GPU box), so
without it the same names are minimal stand-ins with the same constructor signatures and the few inherited methods the
hot path relies on; registrations land in the local dictionaries below, which tests inspect."""
import torch.nn as nn

MODEL_REGISTRY, ARCH_MODEL_REGISTRY, ARCH_CONFIG_REGISTRY, CRITERION_REGISTRY = {}, {}, {}, {}

try:
    from fairseq.models import (FairseqEncoder, FairseqIncrementalDecoder, FairseqEncoderDecoderModel,  # noqa: F401
                                register_model, register_model_architecture)
    from fairseq.criterions import FairseqCriterion, register_criterion  # noqa: F401
    from fairseq import metrics  # noqa: F401
    HAVE_FAIRSEQ = True
except ImportError:
    HAVE_FAIRSEQ = False
    metrics = None

    class FairseqEncoder(nn.Module):
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

        def forward_torchscript(self, net_input):
            return self.forward(**{k: v for k, v in net_input.items() if k != "prev_output_tokens"})

    class FairseqIncrementalDecoder(nn.Module):
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary
            self.onnx_trace = False
            self.adaptive_softmax = None

    class FairseqEncoderDecoderModel(nn.Module):
        def __init__(self, encoder, decoder):
            super().__init__()
            self.encoder, self.decoder = encoder, decoder

        def get_targets(self, sample, net_output):
            return sample["target"]

        def set_num_updates(self, num_updates):
            pass

        def prepare_for_inference_(self, cfg=None):
            self.eval()

        def make_generation_fast_(self, **kwargs):
            self.eval()

    class FairseqCriterion(nn.modules.loss._Loss):
        def __init__(self, task):
            super().__init__()
            self.task = task
            if hasattr(task, "target_dictionary"):
                d = task.target_dictionary
                self.padding_idx = d.pad() if d is not None else -100

    def register_model(name, dataclass=None):
        def deco(cls):
            if name in MODEL_REGISTRY:
                raise ValueError("Cannot register duplicate model ({})".format(name))
            MODEL_REGISTRY[name] = cls
            return cls
        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            if model_name not in MODEL_REGISTRY:
                raise ValueError("Cannot register model architecture for unknown model type ({})".format(model_name))
            ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
            ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn
        return deco

    def register_criterion(name, dataclass=None):
        def deco(cls):
            CRITERION_REGISTRY[name] = cls
            cls.__dataclass = dataclass
            return cls
        return deco
