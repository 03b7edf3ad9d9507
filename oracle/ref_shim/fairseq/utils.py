"""fairseq.utils subset (upstream fairseq/utils.py; restated from its documented behaviour)."""
import torch
import torch.nn.functional as F

def get_available_activation_fns():
    return ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]

def gelu(x):
    # upstream: torch.nn.functional.gelu(x.float()).type_as(x)  (erf form, fp32 math)
    return F.gelu(x.float()).type_as(x)

def get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return gelu
    if activation == "tanh":
        return torch.tanh
    if activation == "linear":
        return lambda x: x
    raise RuntimeError("--activation-fn {} not supported".format(activation))

def new_arange(x, *size):
    if len(size) == 0:
        size = x.size()
    return torch.arange(size[-1], device=x.device).expand(*size).contiguous()

def softmax(x, dim, onnx_trace=False):
    return F.softmax(x, dim=dim, dtype=torch.float32)

def log_softmax(x, dim, onnx_trace=False):
    return F.log_softmax(x, dim=dim, dtype=torch.float32)

def fill_with_neg_inf(t):
    return t.float().fill_(float("-inf")).type_as(t)

def item(tensor):
    if hasattr(tensor, "item"):
        return tensor.item()
    if hasattr(tensor, "__getitem__"):
        return tensor[0]
    return tensor

def eval_str_list(x, type=float):
    if x is None:
        return None
    if isinstance(x, str):
        x = eval(x)
    try:
        return list(map(type, x))
    except TypeError:
        return [type(x)]

def strip_pad(tensor, pad):
    return tensor[tensor.ne(pad)]

def get_perplexity(loss, round=2, base=2):
    return float("inf") if loss is None else round_(base ** loss, round)

def round_(x, n):
    return __builtins__["round"](x, n) if isinstance(__builtins__, dict) else x

def move_to_cuda(sample):
    return sample
