"""Runs the UNMODIFIED reference hot-path files through `oracle/ref_shim`.

TEST INFRASTRUCTURE ONLY.  The reference tree is looked up in $MUSKETEER_REF, /root/reference (the build container) and
baseline/_ref (a byte-for-byte copy of the files this path needs, made by tools/install_reference.py; git-ignored, it
travels to the GPU box with the snapshot).  It is what pins the oracle restatement (`oracle/ofa_oracle.py`), produces
`tests/golden/`, and is timed as the reference arm of bench.py.
"""
import importlib
import importlib.util
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "ref_shim")


def _find_ref():
    for c in (os.environ.get("MUSKETEER_REF"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "models", "ofa")):
            return c
    return os.environ.get("MUSKETEER_REF", "/root/reference")


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "models", "ofa"))


class _FakeEvent:
    # models/ofa/ofa.py:109-166 creates CUDA events and synchronises unconditionally
    def __init__(self, enable_timing=False): pass
    def record(self): pass
    def elapsed_time(self, other): return 0.0


class _KLDiv181(torch.autograd.Function):
    """F.kl_div(input, target, reduction='sum') with the semantics of torch 1.8.1, the version the reference
    pins (README.md:7-9).  torch 2.11's composite kl_div returns NaN when input and log(target) are both -inf
    (constraint-masked vocabulary entries under R-Drop, criterions/label_smoothed_cross_entropy.py:74-78,233);
    1.8.1 (aten/src/ATen/native/Loss.cpp kl_div + tools/autograd kl_div_target_backward, restated from memory
    of that source) zeroes entries with target == 0 in the forward and in both gradients."""
    @staticmethod
    def forward(ctx, inp, tgt):
        ctx.save_for_backward(inp, tgt)
        pos = tgt > 0
        out = torch.where(pos, tgt * (torch.where(pos, tgt, torch.ones_like(tgt)).log()
                                     - torch.where(pos, inp, torch.zeros_like(inp))), torch.zeros_like(tgt))
        return out.sum()

    @staticmethod
    def backward(ctx, g):
        inp, tgt = ctx.saved_tensors
        pos = tgt > 0
        safe_t = torch.where(pos, tgt, torch.ones_like(tgt))
        safe_i = torch.where(pos, inp, torch.zeros_like(inp))
        gi = torch.where(pos, -tgt, torch.zeros_like(tgt)) * g
        gt = torch.where(pos, safe_t.log() + 1 - safe_i, torch.zeros_like(tgt)) * g
        return gi, gt


class _FProxy:
    """torch.nn.functional with kl_div replaced by the torch-1.8.1 behaviour (see _KLDiv181)."""
    def __init__(self, F):
        self._F = F
    def __getattr__(self, k):
        return getattr(self._F, k)
    def kl_div(self, inp, tgt, reduction="mean", log_target=False):
        assert reduction == "sum" and not log_target
        return _KLDiv181.apply(inp, tgt)


_loaded = {}


def load():
    """Import the reference modules; returns a namespace with OFAModel, criterion, generator."""
    if _loaded:
        return _loaded["ns"]
    assert available(), "reference tree not found at %s" % REF
    os.environ["MUSKETEER_REF"] = REF        # the shim's fairseq.search loads models/search.py from the same tree
    for p in (_SHIM, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.cuda.Event = _FakeEvent
        torch.cuda.synchronize = lambda *a, **k: None
    ofa = importlib.import_module("models.ofa.ofa")
    ut = importlib.import_module("models.ofa.unify_transformer")
    spec = importlib.util.spec_from_file_location(
        "_ref_lsce", os.path.join(REF, "criterions", "label_smoothed_cross_entropy.py"))
    lsce = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lsce)
    lsce.F = _FProxy(lsce.F)
    sg = importlib.import_module("models.sequence_generator")
    ns = types.SimpleNamespace(ofa=ofa, ut=ut, lsce=lsce, sg=sg)
    _loaded["ns"] = ns
    return ns


class FakeDictionary:
    def __init__(self, n): self.n = n
    def __len__(self): return self.n
    def pad(self): return 1
    def eos(self): return 2
    def bos(self): return 0
    def unk(self): return 3
    def __eq__(self, o): return isinstance(o, FakeDictionary) and o.n == self.n
    def __contains__(self, x): return x == "<mask>"


class FakeTask:
    def __init__(self, n):
        self.d = FakeDictionary(n)
        self.source_dictionary = self.target_dictionary = self.tgt_dict = self.src_dict = self.d


def build_model(cfg, state_dict=None):
    from . import synth
    ns = load()
    args = synth.to_ref_args(cfg)
    task = FakeTask(cfg.vocab_size)
    model = ns.ofa.OFAModel.build_model(args, task)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=True)
    return model, task


def build_criterion(task, label_smoothing=0.1, use_rdrop=False, reg_alpha=1.0, sample_patch_num=0,
                    ignore_prefix_size=0, ignore_eos=False, drop_worst_ratio=0.0, drop_worst_after=0,
                    constraint_range=None, sentence_avg=False):
    ns = load()
    return ns.lsce.AdjustLabelSmoothedCrossEntropyCriterion(
        task, sentence_avg, label_smoothing, ignore_prefix_size=ignore_prefix_size, ignore_eos=ignore_eos,
        drop_worst_ratio=drop_worst_ratio, drop_worst_after=drop_worst_after, use_rdrop=use_rdrop,
        reg_alpha=reg_alpha, sample_patch_num=sample_patch_num, constraint_range=constraint_range)


def build_generator(model, task, **kw):
    ns = load()
    return ns.sg.SequenceGenerator([model], task.target_dictionary, **kw)
