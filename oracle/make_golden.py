"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_shim.py) on CPU.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The fixtures are small: tiny-model cases store their weights; the full-size case stores only
outputs, its weights being re-derivable from egovlp_b200.synthetic.seeded_state_dict(seed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from egovlp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TINY = syn.TINY_DIMS


def npz(name, **arrs):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def video_tiny():
    _, vt, _ = ref_shim.modules()
    sd = syn.seeded_state_dict(TINY, seed=3, text=False, proj=False)
    net = vt.SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=4,
                                  time_init="zeros", num_classes=0)
    net.pre_logits = torch.nn.Identity()
    missing = net.load_state_dict({k[len("video_model."):]: v for k, v in sd.items()}, strict=True)
    print("video_tiny load:", missing)
    net.eval()
    video = syn.synthetic_video(2, 3, seed=5, img=32)            # T=3 < num_frames=4 (quirk 5)
    out = net(video)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out * probe).sum().backward()
    grads = {n: p.grad for n, p in net.named_parameters()}
    npz("video_tiny.npz", video=video, out=out, probe=probe, seed=3,
        **{"g:video_model." + n: g for n, g in grads.items() if g is not None and (
            "blocks.1.timeattn.qkv" in n or "blocks.0.attn.proj.weight" in n or "blocks.0.mlp.fc1.bias" in n
            or n in ("cls_token", "temporal_embed", "pos_embed", "patch_embed.proj.weight", "norm.weight",
                     "blocks.0.norm3.weight", "blocks.1.norm1.bias"))})


def distilbert_tiny():
    from transformers import DistilBertConfig, DistilBertModel
    sd = syn.seeded_state_dict(TINY, seed=4, video=False, proj=False)
    cfg = DistilBertConfig(vocab_size=120, dim=128, n_layers=2, n_heads=2, hidden_dim=256, max_position_embeddings=32,
                           dropout=0.0, attention_dropout=0.0)
    net = DistilBertModel(cfg).eval()
    print("distilbert_tiny load:", net.load_state_dict({k[len("text_model."):]: v for k, v in sd.items()}, strict=True))
    text = syn.synthetic_text(5, 9, seed=1, ragged=True, vocab=120)
    out = net(**text).last_hidden_state
    npz("distilbert_tiny.npz", input_ids=text["input_ids"], attention_mask=text["attention_mask"], out=out, seed=4)


def losses():
    mm, _, ml = ref_shim.modules()
    g = torch.Generator().manual_seed(21)
    G = 9
    a = torch.randn(G, 16, generator=g)
    b = torch.randn(G, 16, generator=g)
    verb, noun = syn.synthetic_tags(G, seed=2, n_verb=6, n_noun=10)       # small vocab -> several positives
    x = mm.sim_matrix(a, b)
    sv, sn = mm.sim_matrix(verb, verb), mm.sim_matrix(noun, noun)
    out = dict(a=a, b=b, verb=verb, noun=noun, x=x, sim_v=sv, sim_n=sn)
    x_req = x.clone().requires_grad_(True)
    l = ref_shim.cpu_egonce(x_req, sv, sn)
    l.backward()
    out["egonce"], out["egonce_dx"] = l, x_req.grad
    out["egonce_noun_only"] = ref_shim.cpu_egonce(x, sv, sn, noun=True, verb=False)
    out["egonce_verb_only"] = ref_shim.cpu_egonce(x, sv, sn, noun=False, verb=True)
    out["egonce_t007"] = ref_shim.cpu_egonce(x, sv, sn, temperature=0.07)
    x_req = x.clone().requires_grad_(True)
    l = ml.NormSoftmaxLoss()(x_req)
    l.backward()
    out["infonce"], out["infonce_dx"] = l, x_req.grad
    x_req = x.clone().requires_grad_(True)
    l = ml.MaxMarginRankingLoss(margin=0.2, fix_norm=True)(x_req)
    l.backward()
    out["maxmargin"], out["maxmargin_dx"] = l, x_req.grad
    out["maxmargin_nofix"] = ml.MaxMarginRankingLoss(margin=0.2, fix_norm=False)(x)
    # EgoMCQ scoring exactly as trainer/trainer_egoclip.py:214 + model/metric.py:227 per query
    Q, K = 6, 5
    t = torch.randn(Q, 16, generator=g)
    v = torch.randn(Q, K, 16, generator=g)
    v[2, 3] = v[2, 1]                                   # exact tie -> argmax must return the lower index
    scores = torch.stack([mm.sim_matrix(t[q:q + 1], v[q])[0] for q in range(Q)])
    out.update(mcq_text=t, mcq_video=v, mcq_scores=scores, mcq_pred=torch.argmax(scores, dim=1))
    npz("losses.npz", **out)


def full_cfg1():
    """BASELINE.json configs[0]: FrozenInTime on 2x4x3x224x224 video + 8-token text (model built for 16 frames)."""
    mm, _, ml = ref_shim.modules()
    dims = syn.model_dims(num_frames=16)
    sd = syn.seeded_state_dict(dims, seed=0)
    net = ref_shim.build_reference_model(num_frames=16)
    print("full load:", net.load_state_dict(sd, strict=True))
    net.train()                                         # dropout p=0 everywhere via the patched DistilBertConfig
    video = syn.synthetic_video(2, 4, seed=0)
    text = syn.synthetic_text(2, 8, seed=0, ragged=True)
    text_emb, video_emb = net({"video": video, "text": text})
    x = mm.sim_matrix(text_emb, video_emb)
    loss = ml.NormSoftmaxLoss()(x)
    loss.backward()
    sel = {}
    for n, p in net.named_parameters():
        if n in ("video_model.blocks.11.mlp.fc2.bias", "video_model.blocks.0.timeattn.qkv.bias",
                 "video_model.blocks.5.attn.proj.bias", "video_model.norm.weight", "video_model.cls_token",
                 "video_model.temporal_embed", "txt_proj.1.bias", "vid_proj.0.bias",
                 "text_model.transformer.layer.5.output_layer_norm.weight",
                 "text_model.transformer.layer.0.attention.q_lin.bias"):
            sel["g:" + n] = p.grad
        elif n in ("video_model.blocks.0.mlp.fc1.weight", "video_model.blocks.7.timeattn.qkv.weight",
                   "video_model.patch_embed.proj.weight", "text_model.transformer.layer.2.ffn.lin1.weight"):
            sel["g:" + n + "[:8]"] = p.grad.reshape(p.shape[0], -1)[:8]
    gnorm = {"n:" + n: p.grad.norm() for n, p in net.named_parameters() if p.grad is not None}
    npz("full_cfg1.npz", text_emb=text_emb, video_emb=video_emb, sim=x, infonce=loss, seed=0, **sel, **gnorm)


def retrieval():
    """EPIC-MIR side: AdaptiveMaxMarginRankingLoss, utils/nDCG.py, utils/mAP.py and the whole
    model/metric.py:mir_metrics flow (run on tiny annotation files written to a temp dir with the paths it expects)."""
    import pickle
    import tempfile
    mm, _, ml = ref_shim.modules()
    import model.metric as ref_metric
    from utils import nDCG as ref_ndcg, mAP as ref_map
    g = torch.Generator().manual_seed(21)
    out = {}
    n = 12
    x = torch.randn(n, n, generator=g) * 0.3
    w = torch.rand(n, generator=g)
    x_req = x.clone().requires_grad_(True)
    l = ml.AdaptiveMaxMarginRankingLoss(margin=0.4, fix_norm=True)(x_req, w)
    l.backward()
    out.update(amm_x=x, amm_w=w, amm=l, amm_dx=x_req.grad,
               amm_nofix=ml.AdaptiveMaxMarginRankingLoss(margin=0.4, fix_norm=False)(x, w))
    # ranking metrics on tie-free similarities; graded relevancy in {0, .25, .5, 1}, every row has a 1
    rng = np.random.default_rng(5)
    R, Cc = 9, 700
    sim = rng.permutation(R * Cc).reshape(R, Cc).astype(np.float32) / (R * Cc)
    rel = rng.choice([0.0, 0.0, 0.0, 0.25, 0.5, 1.0], size=(R, Cc))
    rel[np.arange(R), rng.integers(0, Cc, R)] = 1.0
    kc = ref_ndcg.calculate_k_counts(rel)
    out.update(rk_sim=sim, rk_rel=rel, rk_kcounts=kc, rk_dcg=ref_ndcg.calculate_DCG(sim, rel, kc),
               rk_idcg=ref_ndcg.calculate_IDCG(rel, kc), rk_ndcg=ref_ndcg.calculate_nDCG(sim, rel),
               rk_ndcg_vec=ref_ndcg.calculate_nDCG(sim, rel, reduction=None),
               rk_ndcg_t=ref_ndcg.calculate_nDCG(sim.T, rel.T), rk_map=ref_map.calculate_mAP(sim, rel),
               rk_map_t=ref_map.calculate_mAP(sim.T, rel.T))
    # the reference's own known-answer example (utils/nDCG.py:141-164)
    ka_sim = np.array([[1.0, 0.7, 0.4, 0.0], [0.3, 0.9, 0.6, 0.1], [0.2, 0.5, 0.8, 0.4]], dtype=np.float32)
    ka_rel = np.array([[1.0, 0.5, 0.0, 0.0], [0.0, 1.0, 0.5, 0.0], [0.0, 0.0, 1.0, 0.5]])
    out.update(ka_sim=ka_sim, ka_rel=ka_rel, ka_ndcg=ref_ndcg.calculate_nDCG(ka_sim, ka_rel),
               ka_map=ref_map.calculate_mAP(ka_sim, ka_rel))
    # full mir_metrics flow: N videos in shuffled loader order, Nt unique sentences
    N, Nt = 40, 17
    video_id = np.array([f"P{i:02d}_{i % 7}" for i in range(N)], dtype=object)
    text_rows = np.sort(rng.choice(N, Nt, replace=False))
    text_id = video_id[text_rows]
    relevancy = rng.choice([0.0, 0.0, 0.3, 1.0], size=(N, Nt))
    relevancy[text_rows, np.arange(Nt)] = 1.0
    relevancy[np.arange(N), rng.integers(0, Nt, N)] = 1.0
    idx_arr = torch.from_numpy(rng.permutation(N))
    t_emb = torch.randn(N, 16, generator=g)
    v_emb = torch.randn(N, 16, generator=g)
    sims = mm.sim_matrix(t_emb, v_emb).numpy()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "dataset/epic-kitchens/epic-kitchens-100-annotations-master/retrieval_annotations")
        os.makedirs(os.path.join(base, "relevancy"))
        import pandas as pd
        pd.DataFrame({"narration_id": video_id}).to_csv(os.path.join(base, "EPIC_100_retrieval_test.csv"), index=False)
        pd.DataFrame({"narration_id": text_id}).to_csv(os.path.join(base, "EPIC_100_retrieval_test_sentence.csv"), index=False)
        with open(os.path.join(base, "relevancy/caption_relevancy_EPIC_100_retrieval_test.pkl"), "wb") as f:
            pickle.dump(relevancy, f)
        os.chdir(tmp)
        try:
            res = ref_metric.mir_metrics(sims, idx_arr)
        finally:
            os.chdir(cwd)
    print("mir_metrics (reference):", res)
    out.update(mir_sims=sims, mir_idx=idx_arr, mir_video_id=video_id.astype(str), mir_text_id=text_id.astype(str),
               mir_relevancy=relevancy, **{"mir_" + k: v for k, v in res.items()})
    npz("retrieval.npz", **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["video_tiny", "distilbert_tiny", "losses", "full_cfg1", "retrieval"]
    for w in which:
        globals()[w]()
