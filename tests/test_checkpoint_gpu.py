"""Checkpoint I/O and feature extraction on the GPU: the reference trainer's checkpoint format
(base/base_trainer.py:399-421) round-trips through the CUDA model + fused AdamW and continues the same trajectory,
`FrozenInTime(load_checkpoint=...)` inflates a 4-frame checkpoint into a 16-frame model (model/model.py:88-95,145-187),
and the dense feature loops of run/test_nlq.py:60-109 are batch-size invariant."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_tower(seed=11):
    from egovlp_b200 import synthetic as syn
    from egovlp_b200.model.video_transformer import SpaceTimeTransformer
    sd = syn.seeded_state_dict(syn.TINY_DIMS, seed=seed, text=False, proj=False)
    net = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=4,
                               time_init="zeros", num_classes=0)
    net.load_state_dict({k[len("video_model."):]: v for k, v in sd.items()})
    return net.cuda()


def _step(net, opt, video, text, verb, noun):
    from egovlp_b200.model.loss import EgoNCE
    opt.zero_grad(set_to_none=True)
    loss = EgoNCE().fused(text, net(video), verb, noun)
    loss.backward()
    opt.step()
    return loss.item()


def test_trainer_checkpoint_roundtrip_continues_identically(tmp_path):
    from egovlp_b200 import synthetic as syn
    from egovlp_b200.optim import AdamW
    video = syn.synthetic_video(8, 4, seed=2, img=32).cuda()
    text = torch.randn(8, 128, generator=torch.Generator().manual_seed(4)).cuda()
    verb, noun = [t.cuda() for t in syn.synthetic_tags(8, seed=3)]
    net = _tiny_tower()
    opt = AdamW(net.parameters(), lr=1e-3)
    for _ in range(3):
        _step(net, opt, video, text, verb, noun)
    path = str(tmp_path / "checkpoint-epoch1.pth")
    torch.save({"arch": type(net).__name__, "epoch": 1, "state_dict": net.state_dict(), "optimizer": opt.state_dict(),
                "monitor_best": 0.0, "config": {"arch": {}, "optimizer": {"type": "AdamW"}}}, path)
    cont = [_step(net, opt, video, text, verb, noun) for _ in range(3)]

    ckpt = torch.load(path, map_location="cuda:0")                       # base_trainer.py:432
    net2 = _tiny_tower(seed=99)                                          # different init: everything must come from the file
    net2.load_state_dict({("module." + k)[7:]: v for k, v in ckpt["state_dict"].items()})   # undo_dp branch (:451-457)
    opt2 = AdamW(net2.parameters(), lr=1e-3)
    _step(net2, opt2, video, text, verb, noun)                           # builds state + pointer table, then is overwritten
    net2.load_state_dict(ckpt["state_dict"])
    opt2.load_state_dict(ckpt["optimizer"])                              # :476
    resumed = [_step(net2, opt2, video, text, verb, noun) for _ in range(3)]
    # not bitwise: split-K weight gradients are accumulated with fp32 atomics, whose order varies run to run
    assert resumed == pytest.approx(cont, rel=1e-5), (resumed, cont)
    for (n1, p1), (_, p2) in zip(net.named_parameters(), net2.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=1e-4, atol=2e-6, msg=n1)


@pytest.fixture(scope="module")
def full_models(tmp_path_factory):
    """A 4-frame FrozenInTime saved in trainer format, and a 16-frame one built from it via load_checkpoint."""
    from egovlp_b200 import synthetic as syn
    from egovlp_b200.model.model import FrozenInTime
    vp = {"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 4, "pretrained": True,
          "time_init": "zeros"}
    tp = {"model": "distilbert-base-uncased", "pretrained": True, "input": "text"}
    m4 = FrozenInTime(video_params=dict(vp), text_params=dict(tp), projection="minimal", load_checkpoint="")
    m4.load_state_dict(syn.seeded_state_dict(syn.model_dims(num_frames=4), seed=0))
    path = str(tmp_path_factory.mktemp("ckpt") / "egovlp_4f.pth")
    torch.save({"state_dict": {"module." + k: v for k, v in m4.state_dict().items()}}, path)     # saved under DDP
    os.environ.setdefault("LOCAL_RANK", "0")
    m16 = FrozenInTime(video_params=dict(vp, num_frames=16), text_params=dict(tp), projection="minimal",
                       load_checkpoint=path, load_temporal_fix="zeros")
    for m in (m4, m16):
        m.cuda().eval()
        m.set_device(torch.device("cuda:0"))
    return m4, m16


def test_load_checkpoint_inflates_temporal_embed(full_models):
    from egovlp_b200 import synthetic as syn
    m4, m16 = full_models
    te4, te16 = m4.video_model.temporal_embed, m16.video_model.temporal_embed
    assert te16.shape[1] == 16 and torch.equal(te16[:, :4], te4) and torch.count_nonzero(te16[:, 4:]) == 0
    video = syn.synthetic_video(3, 4, seed=7).cuda()
    with torch.no_grad():
        assert torch.equal(m16.compute_video(video), m4.compute_video(video))       # first 4 temporal slots are used


def test_dense_features_are_batch_invariant_and_match_oracle(full_models):
    from egovlp_b200 import features, synthetic as syn
    from oracle import reference_port as rp
    m4, _ = full_models
    frames = syn.synthetic_video(1, 38, seed=5)[0]                       # 38 frames -> 9 windows of 4 (2 frames dropped)
    ref_loop = features.dense_video_features(m4, frames, 4, batch=4, reference_tail=True)
    big = features.dense_video_features(m4, frames, 4, batch=64)
    assert big.shape == (9, 256)
    assert torch.equal(ref_loop[:8], big[:8]) and torch.count_nonzero(ref_loop[8]) == 0   # reference leaves the tail at 0
    sd = {k: v.detach().cpu() for k, v in m4.state_dict().items()}
    want = rp.compute_video(frames[:36].reshape(9, 4, 3, 224, 224)[[0, 8]], sd)
    got = big[[0, 8]]
    assert ((got - want).norm() / want.norm()).item() < 1e-2
    text = {k: v.cuda() for k, v in syn.synthetic_text(2, 9, seed=1, ragged=True).items()}
    tok = features.text_features(m4, text, token=True)
    n_words = int(text["attention_mask"][0].sum())
    assert tok.shape == (n_words - 2, 256)
    want_tok = rp.compute_text_tokens({k: v.cpu() for k, v in text.items()}, sd)[0][1:n_words - 1]
    assert ((tok.float().cpu() - want_tok).norm() / want_tok.norm()).item() < 1e-2
    sent = features.text_features(m4, text)
    want_sent = rp.compute_text({k: v.cpu() for k, v in text.items()}, sd)
    assert ((sent.float().cpu() - want_sent).norm() / want_sent.norm()).item() < 1e-2
