"""fairseq `--user-dir` entry point (the reference's is ofa_module/__init__.py:1-5): importing this module registers

  * model `ofa` = musketeer_b200.OFAModel                                   (models/ofa/ofa.py:25)
  * architectures ofa_tiny / ofa_medium / ofa_base / ofa_large / ofa_huge     (models/ofa/ofa.py:370-486)
  * criterion `adjust_label_smoothed_cross_entropy` + its config dataclass    (criterions/label_smoothed_cross_entropy.py:129-131)

in fairseq's registries when fairseq is importable (else in the local registries of musketeer_b200._fairseq_compat, so the
same code path is exercised on machines without fairseq), and installs the overlapped data-parallel wrapper behind
`fairseq.models.DistributedFairseqModel`.  INTEGRATION.md shows the reference-side edit."""
from . import _fairseq_compat as fc
from .archs import ARCHS
from .criterion import AdjustLabelSmoothedCrossEntropyCriterion, AdjustLabelSmoothedCrossEntropyCriterionConfig
from .dp import install_fairseq_ddp_wrapper
from .ofa import OFAModel

MODEL_NAME = "ofa"
CRITERION_NAME = "adjust_label_smoothed_cross_entropy"
_done = False


def register():
    global _done
    if _done:
        return
    fc.register_model(MODEL_NAME)(OFAModel)
    for name, fn in ARCHS.items():
        fc.register_model_architecture(MODEL_NAME, name)(fn)
    fc.register_criterion(CRITERION_NAME, dataclass=AdjustLabelSmoothedCrossEntropyCriterionConfig)(
        AdjustLabelSmoothedCrossEntropyCriterion)
    install_fairseq_ddp_wrapper()
    _done = True


register()
REGISTERED = fc.HAVE_FAIRSEQ      # True: the registrations above went into fairseq's own registries
