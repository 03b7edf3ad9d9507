"""fairseq.modules subset. LayerNorm == torch.nn.LayerNorm (upstream falls back to it without apex)."""
import torch
from torch import nn
from .fairseq_dropout import FairseqDropout

def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True, export=False):
    return nn.LayerNorm(normalized_shape, eps, elementwise_affine)

class LayerDropModuleList(nn.ModuleList):
    def __init__(self, p, modules=None):
        super().__init__(modules)
        self.p = p

class AdaptiveSoftmax(nn.Module): pass
class BaseLayer(nn.Module): pass
class SinusoidalPositionalEmbedding(nn.Module): pass
class GradMultiply(torch.autograd.Function): pass
