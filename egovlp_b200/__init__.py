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


def install_as_reference_model(patch_optimizer=True):
    """Alias the package under the reference's module names.  With `patch_optimizer` the optimizer the reference's
    configs name -- `getattr(transformers, 'AdamW')`, run/train_egoclip.py:72-73 -- resolves to the fused
    egovlp_b200.optim.AdamW (same HF semantics; transformers 5.x removed the class).  Any other optimizer keeps
    working: the bf16 weight copies are refreshed every training forward (engine.Bf16Cache.refresh)."""
    pkg = importlib.import_module("egovlp_b200.model")
    sys.modules["model"] = pkg
    for name in ("model", "loss", "video_transformer", "metric"):
        sys.modules["model." + name] = importlib.import_module("egovlp_b200.model." + name)
    if patch_optimizer:
        import transformers
        from .optim import AdamW
        transformers.AdamW = AdamW
    return pkg
