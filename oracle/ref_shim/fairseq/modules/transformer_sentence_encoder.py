"""init_bert_params (upstream fairseq/modules/transformer_sentence_encoder.py): N(0, 0.02) for
Linear / Embedding weights (+ MultiheadAttention q/k/v), zero biases, zero padding row."""
import torch
from torch import nn

def init_bert_params(module):
    def normal_(data):
        data.copy_(data.cpu().normal_(mean=0.0, std=0.02).to(data.device))
    if isinstance(module, nn.Linear):
        normal_(module.weight.data)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        normal_(module.weight.data)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    if module.__class__.__name__ == "MultiheadAttention" and hasattr(module, "q_proj"):
        normal_(module.q_proj.weight.data)
        normal_(module.k_proj.weight.data)
        normal_(module.v_proj.weight.data)
