"""Import shim for the UNMODIFIED reference (showlab/EgoVLP) mounted at /root/reference.

TEST INFRASTRUCTURE ONLY.  This module is used by ``oracle/make_golden.py`` inside the
build container to run the reference's own Python modules on CPU and record golden
vectors under ``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so
nothing at test/bench run time imports this file (tests that would need it skip).

What the shim does:
  * registers stub modules for packages the reference imports at module scope but the
    hot path never calls (timm.models.layers.{DropPath,to_2tuple,trunc_normal_}, decord,
    av, ffmpeg, humanize, ipdb, tensorboardX, sacred, dominate);
  * puts /root/reference on sys.path so ``model.model`` / ``model.loss`` /
    ``model.video_transformer`` import unmodified;
  * offers ``build_reference_model`` which constructs ``FrozenInTime`` without the
    pretrained files by patching ``torch.load`` / ``AutoModel.from_pretrained`` for the
    duration of the constructor (random-init DistilBERT of the same architecture);
  * offers ``cpu_egonce`` that runs ``model.loss.EgoNCE`` with ``Tensor.cuda`` patched to
    identity (reference hard-codes ``.cuda()`` at model/loss.py:35).
No reference source is copied: everything is executed from where it lies.
"""
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EGOVLP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__spec__ = None
    sys.modules[name] = mod
    return mod


_installed = False


def install():
    """Idempotently install the stubs and make the reference importable."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    import torch
    import transformers  # noqa: F401  (must precede the `timm` stub: transformers probes timm.__spec__)
    from transformers import AutoModel  # noqa: F401

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert self.p == 0.0 or not self.training
            return x

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    if "timm" not in sys.modules:
        timm = _stub("timm")
        models = _stub("timm.models")
        layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                       trunc_normal_=torch.nn.init.trunc_normal_)
        timm.models = models
        models.layers = layers
    for name in ("av", "ffmpeg", "humanize", "ipdb", "dominate", "sacred", "tensorboardX"):
        if name not in sys.modules:
            _stub(name)
    if "decord" not in sys.modules:
        bridge = types.SimpleNamespace(set_bridge=lambda *a, **k: None)
        _stub("decord", bridge=bridge, VideoReader=object, cpu=lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def modules():
    """Returns (model.model, model.video_transformer, model.loss) of the reference."""
    install()
    import model.model as mm
    import model.video_transformer as vt
    import model.loss as ml
    assert mm.__file__.startswith(REFERENCE_ROOT), mm.__file__
    return mm, vt, ml


@contextlib.contextmanager
def _patched_constructors(distilbert_cfg=None):
    import torch
    from transformers import DistilBertConfig, DistilBertModel
    mm, _, _ = modules()
    real_load, real_fp = torch.load, mm.AutoModel.from_pretrained

    def fake_load(path, *a, **k):
        if str(path).endswith("jx_vit_base_p16_224-80ecf9dd.pth"):
            return {"cls_token": torch.zeros(1, 1, 768)}
        return real_load(path, *a, **k)

    def fake_from_pretrained(*a, **k):
        cfg = distilbert_cfg or DistilBertConfig(dropout=0.0, attention_dropout=0.0)
        return DistilBertModel(cfg)

    torch.load = fake_load
    mm.AutoModel.from_pretrained = fake_from_pretrained
    try:
        yield
    finally:
        torch.load = real_load
        mm.AutoModel.from_pretrained = real_fp


def build_reference_model(num_frames=4, projection_dim=256):
    """FrozenInTime(video_params, text_params, ...) exactly as configs/pt/egoclip.json builds it,
    minus the pretrained files (model/model.py:15-95)."""
    mm, _, _ = modules()
    video_params = {"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224",
                    "num_frames": num_frames, "pretrained": True, "time_init": "zeros"}
    text_params = {"model": "distilbert-base-uncased", "pretrained": True, "input": "text"}
    with _patched_constructors():
        net = mm.FrozenInTime(video_params, text_params, projection_dim=projection_dim,
                              load_checkpoint=None, projection="minimal")
    return net


@contextlib.contextmanager
def cuda_is_identity():
    import torch
    real = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = real


def cpu_egonce(x, sim_v, sim_n, **kw):
    _, _, ml = modules()
    with cuda_is_identity():
        return ml.EgoNCE(**kw)(x, sim_v, sim_n)
