"""egovlp_b200: B200-native implementation of the EgoVLP dual-encoder pretraining hot path.

    from egovlp_b200.model.model import FrozenInTime, sim_matrix
    from egovlp_b200.model.loss import EgoNCE, NormSoftmaxLoss, MaxMarginRankingLoss

`install_as_reference_model()` aliases the package under the reference's module names (`model.model`,
`model.loss`, `model.video_transformer`, `model.metric`) so the reference's run/ and trainer/ scripts import it
unchanged.
"""
import importlib
import sys

__version__ = "0.1.0"


def install_as_reference_model():
    pkg = importlib.import_module("egovlp_b200.model")
    sys.modules["model"] = pkg
    for name in ("model", "loss", "video_transformer", "metric"):
        sys.modules["model." + name] = importlib.import_module("egovlp_b200.model." + name)
    return pkg
