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
