"""EPIC-Kitchens MIR side of the path on the GPU, through the C-ABI: AdaptiveMaxMarginRankingLoss
and the ranking metrics (egovlp_rank_metrics behind the utils/nDCG.py / utils/mAP.py / model/metric.py mirrors),
against golden values recorded from the unmodified reference and against the numpy oracle at larger sizes.
Metrics are accumulated in fp64 on both sides: tolerance 1e-9 relative; ranking itself is exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(GOLDEN, "retrieval.npz"))


def test_adaptive_max_margin_vs_reference_golden(z):
    from egovlp_b200.model.loss import AdaptiveMaxMarginRankingLoss
    x = torch.from_numpy(z["amm_x"]).cuda().requires_grad_(True)
    w = torch.from_numpy(z["amm_w"]).cuda()
    loss = AdaptiveMaxMarginRankingLoss(margin=0.4, fix_norm=True)(x, w)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), torch.from_numpy(z["amm"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(x.grad.cpu(), torch.from_numpy(z["amm_dx"]), rtol=1e-5, atol=1e-7)
    nofix = AdaptiveMaxMarginRankingLoss(margin=0.4, fix_norm=False)(x.detach(), w)
    torch.testing.assert_close(nofix.cpu(), torch.from_numpy(z["amm_nofix"]), rtol=1e-5, atol=1e-6)
    with pytest.raises(AttributeError):
        AdaptiveMaxMarginRankingLoss()(x)


def test_ndcg_map_vs_reference_golden(z):
    from egovlp_b200.utils import nDCG, mAP
    sim, rel = z["rk_sim"], z["rk_rel"]
    kc = nDCG.calculate_k_counts(rel)
    assert isinstance(kc, np.ndarray) and np.array_equal(kc, z["rk_kcounts"])
    np.testing.assert_allclose(nDCG.calculate_DCG(sim, rel, kc), z["rk_dcg"], rtol=1e-9)
    np.testing.assert_allclose(nDCG.calculate_IDCG(rel, kc), z["rk_idcg"], rtol=1e-9)
    np.testing.assert_allclose(nDCG.calculate_nDCG(sim, rel), z["rk_ndcg"], rtol=1e-9)
    np.testing.assert_allclose(nDCG.calculate_nDCG(sim, rel, kc, IDCG=z["rk_idcg"]), z["rk_ndcg"], rtol=1e-9)
    np.testing.assert_allclose(nDCG.calculate_nDCG(sim, rel, reduction=None), z["rk_ndcg_vec"], rtol=1e-9)
    np.testing.assert_allclose(mAP.calculate_mAP(sim, rel), z["rk_map"], rtol=1e-9)
    simT, relT = np.ascontiguousarray(sim.T), np.ascontiguousarray(rel.T)
    np.testing.assert_allclose(nDCG.calculate_nDCG(simT, relT), z["rk_ndcg_t"], equal_nan=True)   # NaN rows, as numpy
    np.testing.assert_allclose(mAP.calculate_mAP(simT, relT), z["rk_map_t"], equal_nan=True)
    ap_t = mAP.calculate_AP(simT, relT)
    ref_t = rp.average_precision(simT, relT)
    np.testing.assert_allclose(ap_t, ref_t, rtol=1e-9, equal_nan=True)
    # the reference's own known-answer example (utils/nDCG.py:141-164)
    np.testing.assert_allclose(nDCG.calculate_nDCG(z["ka_sim"], z["ka_rel"]), z["ka_ndcg"], rtol=1e-9)
    np.testing.assert_allclose(mAP.calculate_mAP(z["ka_sim"], z["ka_rel"]), z["ka_map"], rtol=1e-9)
    # torch in -> torch out, nothing leaves the device
    out = nDCG.calculate_nDCG(torch.from_numpy(sim).cuda(), torch.from_numpy(rel).cuda(), reduction=None)
    assert out.is_cuda and out.dtype == torch.float64


def test_mir_metrics_flow_vs_reference_golden(z):
    from egovlp_b200.model.metric import mir_metrics_core
    res = mir_metrics_core(z["mir_sims"], torch.from_numpy(z["mir_idx"]), z["mir_video_id"], z["mir_text_id"],
                           z["mir_relevancy"])
    for k in ("nDCG_V2T", "nDCG_T2V", "nDCG_AVG", "mAP_V2T", "mAP_T2V", "mAP_AVG"):
        np.testing.assert_allclose(res[k], float(z["mir_" + k]), rtol=1e-6, err_msg=k)   # sims are fp32: (s+1)/2 rounding


@pytest.mark.parametrize("cols", [1, 2, 37, 511, 512, 513, 5000, 16384])
@pytest.mark.parametrize("rel_dtype", [np.float32, np.float64])
def test_rank_metrics_vs_oracle_with_ties(cols, rel_dtype):
    """Quantised similarities (many exact ties), both tie rules, explicit and implicit k_counts."""
    from egovlp_b200 import ops
    rng = np.random.default_rng(cols)
    rows = 19
    sim = (rng.integers(0, max(2, cols // 3), size=(rows, cols)) / 64.0).astype(np.float32)
    rel = rng.choice([0.0, 0.0, 0.25, 1.0], size=(rows, cols)).astype(rel_dtype)
    rel[0] = 0                                                   # a query with nothing relevant
    kc_custom = (rng.random((rows, cols)) < 0.3).astype(np.int32)
    s_d, r_d = torch.from_numpy(sim).cuda(), torch.from_numpy(rel).cuda()
    with np.errstate(invalid="ignore", divide="ignore"):
        want_dcg = rp.dcg(sim, rel.astype(np.float64), rp.k_counts_of(rel))
        want_dcg_kc = rp.dcg(sim, rel.astype(np.float64), kc_custom)
        want_ap = rp.average_precision(sim, rel.astype(np.float64))
    dcg, _ = ops.rank_metrics(s_d, r_d, None, tie_mode=1, want_ap=False)
    np.testing.assert_allclose(dcg.cpu().numpy(), want_dcg, rtol=1e-9, atol=1e-12)
    dcg_kc, _ = ops.rank_metrics(s_d, r_d, torch.from_numpy(kc_custom).cuda(), tie_mode=1, want_ap=False)
    np.testing.assert_allclose(dcg_kc.cpu().numpy(), want_dcg_kc, rtol=1e-9, atol=1e-12)
    _, ap = ops.rank_metrics(s_d, r_d, None, tie_mode=0, want_dcg=False)
    np.testing.assert_allclose(ap.cpu().numpy(), want_ap, rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.isnan(ap[0].item())


def test_rank_metrics_limits():
    from egovlp_b200 import ops
    from egovlp_b200._lib import EgovlpError
    s = torch.zeros(2, 16385, device="cuda")
    with pytest.raises(EgovlpError, match="16384"):
        ops.rank_metrics(s, s.clone())
    d, a = ops.rank_metrics(torch.zeros(0, 8, device="cuda"), torch.zeros(0, 8, device="cuda"))
    assert d.shape == (0,) and a.shape == (0,)
