"""musketeer_b200: the OFA encoder-decoder forward/backward hot path of amazon-science/musketeer on B200 (sm_100a).

    from musketeer_b200 import OFAModel, AdjustLabelSmoothedCrossEntropyCriterion

The CUDA library (libofa_b200.so, built by `python -m musketeer_b200.build`) is loaded on first kernel call; a missing
library is an error, never a fallback."""
from .archs import ARCHS, ofa_base_architecture, ofa_huge_architecture, ofa_large_architecture, \
    ofa_medium_architecture, ofa_tiny_architecture  # noqa: F401
from .criterion import AdjustLabelSmoothedCrossEntropyCriterion  # noqa: F401
from .ofa import OFAModel  # noqa: F401
