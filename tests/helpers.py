"""Shared helpers for the parity tests: rebuild a golden case (weights + batches) from its recipe."""
import copy
import json
import os
import random

import torch

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    fx["case"] = json.loads(fx["recipe"])
    return fx


def build_case(case, emb_std=None):
    cfg = synth.make_cfg(case["arch"], **case["cfg"])
    kw = {} if emb_std is None else {"emb_std": emb_std}
    sd = synth.synth_state_dict(cfg, seed=0, **kw)
    samples = []
    for t in case.get("tasks", []):
        s = synth.make_batch(**t)
        if "patch_masks" in case:
            s["net_input"]["patch_masks"] = torch.tensor(case["patch_masks"])
        samples.append(s)
    return cfg, sd, samples


def tie(sd, requires_grad=True):
    out = {k: v.clone().requires_grad_(requires_grad and v.is_floating_point()) for k, v in sd.items()}
    out["decoder.embed_tokens.weight"] = out["encoder.embed_tokens.weight"]
    out["decoder.output_projection.weight"] = out["encoder.embed_tokens.weight"]
    return out


ZERO_GRAD_SUFFIXES = ("k_proj.bias", "pos_k_linear.bias")  # mathematically zero (softmax shift invariance)


class FakeDictionary:
    """Stands in for the fairseq Dictionary the reference task hands to build_model (tasks/ofa_task.py:93-116)."""
    def __init__(self, n): self.n = n
    def __len__(self): return self.n
    def pad(self): return 1
    def eos(self): return 2
    def bos(self): return 0
    def unk(self): return 3
    def __eq__(self, o): return isinstance(o, FakeDictionary) and o.n == self.n
    def __ne__(self, o): return not self.__eq__(o)


class FakeTask:
    def __init__(self, n):
        d = FakeDictionary(n)
        self.source_dictionary = self.target_dictionary = self.tgt_dict = self.src_dict = d


def build_product(cfg, sd, dtype=torch.float32, device="cuda"):
    """Instantiate musketeer_b200.OFAModel for an oracle config and load the synthetic weights."""
    from types import SimpleNamespace
    from musketeer_b200 import OFAModel
    args = SimpleNamespace(**vars(cfg))
    args.no_scale_embedding = True
    args.activation_fn = "gelu"
    task = FakeTask(cfg.vocab_size)
    model = OFAModel.build_model(args, task)
    model.load_state_dict(sd, strict=True)
    model = model.to(device=device, dtype=dtype)
    return model, task


def to_device(obj, device, float_dtype=None):
    if isinstance(obj, torch.Tensor):
        if obj.is_floating_point() and float_dtype is not None:
            return obj.to(device=device, dtype=float_dtype)
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device, float_dtype) for k, v in obj.items()}
    if isinstance(obj, list):
        return [to_device(v, device, float_dtype) for v in obj]
    return obj
