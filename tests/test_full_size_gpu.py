"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes (configs[2] per-GPU shape: 64 clips x 16
frames x 3 x 224^2, L = 16, G = 512 for the loss; configs[3] max-margin + dual softmax at 4096^2), where the fp32
oracle cannot run in seconds:
  * clips are independent units: the embeddings of a batch do not depend on the other clips in it (bit-exact under
    batch permutation and batch splitting) -- the property data-parallel sharding relies on;
  * the full-size loss kernels agree with the fp32 oracle (the [G, G] part is cheap on the CPU even at G = 512);
  * dual softmax at 4096^2: rows of the first softmax and columns of the result are distributions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    from egovlp_b200 import synthetic as syn
    from egovlp_b200.model.model import FrozenInTime
    m = FrozenInTime(video_params={"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 16,
                                   "pretrained": True, "time_init": "zeros"},
                     text_params={"model": "distilbert-base-uncased", "pretrained": True, "input": "text"},
                     projection="minimal", load_checkpoint="")
    m.load_state_dict(syn.seeded_state_dict(syn.model_dims(num_frames=16), seed=0))
    return m.cuda().eval()


def test_clips_are_independent_units_at_b64_t16(net):
    from egovlp_b200 import synthetic as syn
    B = 64
    video = syn.synthetic_video(B, 16, seed=3).cuda()
    text = {k: v.cuda() for k, v in syn.synthetic_text(B, 16, seed=3, ragged=True).items()}
    with torch.no_grad():
        t, v = net({"video": video, "text": text})
        assert t.shape == (B, 256) and v.shape == (B, 256) and torch.isfinite(t).all() and torch.isfinite(v).all()
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
        tp, vp = net({"video": video[perm], "text": {k: x[perm] for k, x in text.items()}})
        assert torch.equal(vp, v[perm]) and torch.equal(tp, t[perm])
        halves = [net({"video": video[i:i + 32], "text": {k: x[i:i + 32] for k, x in text.items()}}) for i in (0, 32)]
        assert torch.equal(torch.cat([h[1] for h in halves]), v) and torch.equal(torch.cat([h[0] for h in halves]), t)
        one = net.compute_video(video[5:6])
        assert torch.equal(one, v[5:6])
    # distinct clips give distinct, non-degenerate embeddings
    vn = torch.nn.functional.normalize(v.float(), dim=1)
    off = (vn @ vn.t() - torch.eye(B, device=v.device)).abs().max().item()
    assert off < 0.9999


def test_full_size_loss_g512_vs_oracle():
    from egovlp_b200 import synthetic as syn
    from egovlp_b200.model.loss import EgoNCE, MaxMarginRankingLoss, AdaptiveMaxMarginRankingLoss
    from egovlp_b200.model.model import sim_matrix
    from oracle import reference_port as rp
    G = 512
    g = torch.Generator().manual_seed(17)
    t, v = torch.randn(G, 256, generator=g), torch.randn(G, 256, generator=g)
    verb, noun = syn.synthetic_tags(G, seed=17)
    w = torch.rand(G, generator=g)
    tr, vr = t.clone().requires_grad_(True), v.clone().requires_grad_(True)
    want = rp.egonce_loss(rp.sim_matrix(tr, vr), rp.sim_matrix(verb, verb), rp.sim_matrix(noun, noun))
    want.backward()
    tc, vc = t.cuda().requires_grad_(True), v.cuda().requires_grad_(True)
    got = EgoNCE().fused(tc, vc, verb.cuda(), noun.cuda())
    got.backward()
    torch.testing.assert_close(got.cpu(), want.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(tc.grad.cpu(), tr.grad, rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(vc.grad.cpu(), vr.grad, rtol=2e-4, atol=1e-7)
    # the trainer formulation (three similarity matrices) gives the same value
    x = sim_matrix(tc.detach(), vc.detach())
    trainer = EgoNCE()(x, sim_matrix(verb.cuda(), verb.cuda()), sim_matrix(noun.cuda(), noun.cuda()))
    torch.testing.assert_close(trainer, got.detach(), rtol=1e-6, atol=1e-6)
    xr = rp.sim_matrix(t, v)
    torch.testing.assert_close(MaxMarginRankingLoss()(x).cpu(), rp.max_margin_ranking_loss(xr), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(AdaptiveMaxMarginRankingLoss()(x, w.cuda()).cpu(),
                               rp.adaptive_max_margin_ranking_loss(xr, w), rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("G,C,B,rank", [(512, 256, 64, 3), (96, 256, 32, 2), (9, 16, 3, 1), (40, 32, 40, 0)])
def test_fused_gather_egonce_kernel_local_slice_and_modes(G, C, B, rank):
    """ONE forward kernel on column views of the packed gather buffer + ONE backward kernel emitting this rank's rows,
    against the oracle's gathered_step_loss (all G rows) -- every positives mode, ragged G / C, the zero-noun clamp."""
    from egovlp_b200 import engine, synthetic as syn
    from egovlp_b200.model.loss import EgoNCE
    from oracle import reference_port as rp
    g = torch.Generator().manual_seed(G + C)
    t, v = torch.randn(G, C, generator=g), torch.randn(G, C, generator=g)
    verb, noun = syn.synthetic_tags(G, seed=G)
    world = G // B
    sl = slice(rank * B, (rank + 1) * B)
    # the collective is replaced by a function that drops this rank's packed rows into the pre-gathered buffer
    allp = torch.cat([t, v, verb, noun], dim=1).cuda()

    def fake_gather(packed_local):
        torch.testing.assert_close(packed_local, allp[sl])
        return allp

    tl, vl = t[sl].cuda().requires_grad_(True), v[sl].cuda().requires_grad_(True)
    loss = EgoNCE().gathered(tl, vl, verb[sl].cuda(), noun[sl].cuda(), fake_gather, rank, world)
    loss.backward()
    tr, vr = t.clone().requires_grad_(True), v.clone().requires_grad_(True)
    want = rp.egonce_loss(rp.sim_matrix(tr, vr), rp.sim_matrix(verb, verb), rp.sim_matrix(noun, noun))
    want.backward()
    torch.testing.assert_close(loss.cpu(), want.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(tl.grad.cpu(), tr.grad[sl], rtol=3e-4, atol=2e-7)
    torch.testing.assert_close(vl.grad.cpu(), vr.grad[sl], rtol=3e-4, atol=2e-7)
    for kw, ref_kw in (({"noun": True, "verb": False}, {"noun": True, "verb": False}),
                       ({"noun": False, "verb": True}, {"noun": False, "verb": True})):
        got = EgoNCE(**kw).fused(t.cuda(), v.cuda(), verb.cuda(), noun.cuda())
        ref = rp.egonce_loss(rp.sim_matrix(t, v), rp.sim_matrix(verb, verb), rp.sim_matrix(noun, noun), **ref_kw)
        torch.testing.assert_close(got.cpu(), ref, rtol=2e-5, atol=2e-5)
    # InfoNCE = diagonal positives only (mode 0) through the same kernels
    loss0, _ = __import__("egovlp_b200.ops", fromlist=["x"]).egonce_fused_fwd(t.cuda(), v.cuda(), None, None, 20.0, 0)
    torch.testing.assert_close(loss0.cpu(), rp.norm_softmax_loss(rp.sim_matrix(t, v)), rtol=2e-5, atol=2e-5)
    # a second launch reuses the workspace (the ticket word is left at zero)
    again = EgoNCE().fused(t.cuda(), v.cuda(), verb.cuda(), noun.cuda())
    torch.testing.assert_close(again.cpu(), want.detach(), rtol=2e-5, atol=2e-5)


def test_dual_softmax_4096_properties():
    from egovlp_b200 import ops
    g = torch.Generator().manual_seed(23)
    a = torch.nn.functional.normalize(torch.randn(4096, 256, generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(4096, 256, generator=g), dim=1)
    sim = (a @ b.t()).cuda()
    out = ops.dual_softmax(sim)
    assert out.shape == sim.shape and torch.isfinite(out).all() and (out >= 0).all()
    torch.testing.assert_close(out.sum(0), torch.ones(4096, device="cuda"), rtol=1e-4, atol=1e-4)   # softmax over dim 0
    ref = torch.softmax(torch.softmax(sim.double() / 500, 1) * sim.double(), 0)
    torch.testing.assert_close(out.double(), ref, rtol=1e-4, atol=1e-9)
