"""Pin the CPU oracle (oracle/reference_port.py) against golden vectors recorded from the UNMODIFIED
reference (oracle/make_golden.py).  fp32 vs fp32, so tolerances are rounding-level."""
import torch

from conftest import load_golden, weights_of
from oracle import reference_port as rp
from egovlp_b200 import synthetic as syn


def close(a, b, rtol=2e-5, atol=2e-6):
    torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)


def test_video_tower_tiny_forward_and_grads():
    g = load_golden("video_tiny.npz")
    sd = syn.seeded_state_dict(syn.TINY_DIMS, seed=int(g["seed"]), text=False, proj=False)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = rp.video_tower(g["video"], p, heads=2)
    close(out, g["out"], rtol=1e-4, atol=1e-5)
    (out * g["probe"]).sum().backward()
    checked = 0
    for k, v in g.items():
        if k.startswith("g:"):
            close(p[k[2:]].grad, v, rtol=2e-4, atol=2e-5)
            checked += 1
    assert checked >= 10


def test_distilbert_tiny():
    g = load_golden("distilbert_tiny.npz")
    sd = syn.seeded_state_dict(syn.TINY_DIMS, seed=int(g["seed"]), video=False, proj=False)
    out = rp.distilbert_forward(g["input_ids"], g["attention_mask"], sd, heads=2)
    close(out, g["out"], rtol=1e-4, atol=1e-5)


def test_losses_and_sim():
    g = load_golden("losses.npz")
    close(rp.sim_matrix(g["a"], g["b"]), g["x"])
    close(rp.sim_matrix(g["verb"], g["verb"]), g["sim_v"])
    sn = rp.sim_matrix(g["noun"], g["noun"])
    close(sn, g["sim_n"])
    assert torch.all(sn[-1] == 0)                        # zero multi-hot row -> zero similarities, no NaN
    x = g["x"].clone().requires_grad_(True)
    l = rp.egonce_loss(x, g["sim_v"], g["sim_n"])
    l.backward()
    close(l, g["egonce"]); close(x.grad, g["egonce_dx"], atol=1e-6)
    close(rp.egonce_loss(g["x"], g["sim_v"], g["sim_n"], noun=True, verb=False), g["egonce_noun_only"])
    close(rp.egonce_loss(g["x"], g["sim_v"], g["sim_n"], noun=False, verb=True), g["egonce_verb_only"])
    close(rp.egonce_loss(g["x"], g["sim_v"], g["sim_n"], temperature=0.07), g["egonce_t007"])
    x = g["x"].clone().requires_grad_(True)
    l = rp.norm_softmax_loss(x); l.backward()
    close(l, g["infonce"]); close(x.grad, g["infonce_dx"], atol=1e-6)
    x = g["x"].clone().requires_grad_(True)
    l = rp.max_margin_ranking_loss(x); l.backward()
    close(l, g["maxmargin"]); close(x.grad, g["maxmargin_dx"], atol=1e-6)
    close(rp.max_margin_ranking_loss(g["x"], fix_norm=False), g["maxmargin_nofix"])
    s, pred = rp.egomcq_predict(g["mcq_text"], g["mcq_video"])
    close(s, g["mcq_scores"])
    assert torch.equal(pred, g["mcq_pred"])
    assert int(pred[2]) != 3                              # tie resolves to the lower index


def test_full_size_cfg1_forward_backward():
    """BASELINE configs[0]: B=2,T=4,L=8 full-size model, weights re-derived from the seed."""
    g = load_golden("full_cfg1.npz")
    dims = syn.model_dims(num_frames=16)
    p = {k: v.requires_grad_(True) for k, v in syn.seeded_state_dict(dims, seed=int(g["seed"])).items()}
    data = {"video": syn.synthetic_video(2, 4, seed=0), "text": syn.synthetic_text(2, 8, seed=0, ragged=True)}
    t, v = rp.frozen_in_time_forward(data, p)
    close(t, g["text_emb"], rtol=2e-4, atol=2e-5)
    close(v, g["video_emb"], rtol=2e-4, atol=2e-5)
    loss = rp.norm_softmax_loss(rp.sim_matrix(t, v))
    close(loss, g["infonce"], rtol=1e-4)
    loss.backward()
    n = 0
    for k, ref in g.items():
        if k.startswith("g:"):
            name = k[2:]
            if name.endswith("[:8]"):
                got = p[name[:-4]].grad
                got = got.reshape(got.shape[0], -1)[:8]
            else:
                got = p[name].grad
            scale = ref.abs().max().item() + 1e-12
            assert (got - ref).abs().max().item() <= 2e-3 * scale, name
            n += 1
        elif k.startswith("n:"):
            got = p[k[2:]].grad.norm()
            assert abs(got.item() - ref.item()) <= 2e-3 * ref.item() + 1e-7, k  # k_lin.bias grads are analytically 0
    assert n >= 12


def test_retrieval_oracle_vs_reference_golden():
    """EPIC-MIR side: AdaptiveMaxMarginRankingLoss, utils/nDCG.py, utils/mAP.py (golden from the reference itself,
    incl. its own known-answer example utils/nDCG.py:141-164)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "retrieval.npz"))
    x = torch.from_numpy(z["amm_x"]).requires_grad_(True)
    w = torch.from_numpy(z["amm_w"])
    l = rp.adaptive_max_margin_ranking_loss(x, w); l.backward()
    close(l, torch.from_numpy(z["amm"])); close(x.grad, torch.from_numpy(z["amm_dx"]), atol=1e-7)
    close(rp.adaptive_max_margin_ranking_loss(x.detach(), w, fix_norm=False), torch.from_numpy(z["amm_nofix"]))
    sim, rel = z["rk_sim"], z["rk_rel"]
    kc = rp.k_counts_of(rel)
    assert np.array_equal(kc, z["rk_kcounts"])
    np.testing.assert_allclose(rp.dcg(sim, rel, kc), z["rk_dcg"], rtol=1e-12)
    np.testing.assert_allclose(rp.dcg(rel, rel, kc), z["rk_idcg"], rtol=1e-12)
    np.testing.assert_allclose(rp.ndcg(sim, rel), z["rk_ndcg"], rtol=1e-12)
    np.testing.assert_allclose(rp.ndcg(sim, rel, reduction=None), z["rk_ndcg_vec"], rtol=1e-12)
    np.testing.assert_allclose(rp.average_precision(sim, rel).mean(), z["rk_map"], rtol=1e-12)
    with np.errstate(invalid="ignore", divide="ignore"):        # gallery items without any relevant query: NaN, as numpy
        np.testing.assert_allclose(rp.ndcg(sim.T, rel.T), z["rk_ndcg_t"], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(rp.average_precision(sim.T, rel.T).mean(), z["rk_map_t"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(rp.ndcg(z["ka_sim"], z["ka_rel"]), z["ka_ndcg"], rtol=1e-12)
    np.testing.assert_allclose(rp.average_precision(z["ka_sim"], z["ka_rel"]).mean(), z["ka_map"], rtol=1e-12)
