"""fairseq.models subset: registries and the base classes the OFA files subclass.
Restates upstream fairseq behaviour (fairseq/models/{__init__,fairseq_model,fairseq_encoder,
fairseq_decoder,fairseq_incremental_decoder}.py)."""
from typing import Dict, List, Optional, Tuple
import torch
from torch import nn, Tensor
import torch.nn.functional as F
from fairseq import utils

MODEL_REGISTRY = {}
ARCH_MODEL_REGISTRY = {}
ARCH_CONFIG_REGISTRY = {}

def register_model(name, dataclass=None):
    def deco(cls):
        MODEL_REGISTRY[name] = cls
        return cls
    return deco

def register_model_architecture(model_name, arch_name):
    def deco(fn):
        ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
        ARCH_CONFIG_REGISTRY[arch_name] = fn
        return fn
    return deco

class BaseFairseqModel(nn.Module):
    def __init__(self):
        super().__init__()
    def get_targets(self, sample, net_output):
        return sample["target"]
    def get_normalized_probs(self, net_output, log_probs, sample=None):
        return self.get_normalized_probs_scriptable(net_output, log_probs, sample)
    def get_normalized_probs_scriptable(self, net_output, log_probs, sample=None):
        if hasattr(self, "decoder"):
            return self.decoder.get_normalized_probs(net_output, log_probs, sample)
        raise NotImplementedError
    def set_num_updates(self, num_updates):
        for m in self.modules():
            if hasattr(m, "set_num_updates") and m != self:
                m.set_num_updates(num_updates)
    def upgrade_state_dict_named(self, state_dict, name):
        pass

class FairseqEncoder(nn.Module):
    def __init__(self, dictionary):
        super().__init__()
        self.dictionary = dictionary
    def forward_torchscript(self, net_input: Dict[str, Tensor]):
        encoder_input = {k: v for k, v in net_input.items() if k != "prev_output_tokens"}
        return self.forward(**encoder_input)

class FairseqDecoder(nn.Module):
    def __init__(self, dictionary):
        super().__init__()
        self.dictionary = dictionary
        self.onnx_trace = False
        self.adaptive_softmax = None
    def get_normalized_probs(self, net_output, log_probs, sample=None):
        return self.get_normalized_probs_scriptable(net_output, log_probs, sample)
    def get_normalized_probs_scriptable(self, net_output, log_probs, sample=None):
        logits = net_output[0]
        if log_probs:
            return utils.log_softmax(logits, dim=-1)
        return utils.softmax(logits, dim=-1)

class FairseqIncrementalDecoder(FairseqDecoder):
    def reorder_incremental_state_scripting(self, incremental_state, new_order):
        for module in self.modules():
            if hasattr(module, "reorder_incremental_state"):
                result = module.reorder_incremental_state(incremental_state, new_order)
                if result is not None:
                    incremental_state = result

class FairseqEncoderDecoderModel(BaseFairseqModel):
    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
    def max_positions(self):
        return (self.encoder.max_positions(), self.decoder.max_positions())
    def max_decoder_positions(self):
        return self.decoder.max_positions()
