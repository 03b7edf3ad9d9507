"""fairseq `--user-dir` entry: registers this package's OFAModel under the reference's names
(`@register_model("ofa")`, archs ofa_tiny/medium/base/large/huge: models/ofa/ofa.py:25,370-486) and the criterion
`adjust_label_smoothed_cross_entropy` (criterions/label_smoothed_cross_entropy.py:129-131).

fairseq is not installed in the build container, so registration is conditional; INTEGRATION.md shows
the two-line change that points the reference's `models/__init__.py` at this package."""
from .archs import ARCHS
from .ofa import OFAModel
from .criterion import AdjustLabelSmoothedCrossEntropyCriterion

try:  # pragma: no cover - exercised only where fairseq exists
    from fairseq.models import register_model, register_model_architecture, FairseqEncoderDecoderModel
    from fairseq.criterions import register_criterion

    @register_model("ofa")
    class FairseqOFAModel(OFAModel, FairseqEncoderDecoderModel):
        pass

    for _name, _fn in ARCHS.items():
        register_model_architecture("ofa", _name)(_fn)
    register_criterion("adjust_label_smoothed_cross_entropy")(AdjustLabelSmoothedCrossEntropyCriterion)
    REGISTERED = True
except ImportError:
    REGISTERED = False
