"""CPU restatement (torch fp32, functional) of the EgoVLP hot path.  TEST INFRASTRUCTURE ONLY.

This file is the oracle the CUDA path is checked against.  It may be imported by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs, never by the product package ``egovlp_b200``.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the UNMODIFIED
reference executed in the build container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; checked in ``tests/test_oracle_golden.py``).

Every function names the reference lines it restates (paths relative to
/root/reference).  Weights are passed as a flat ``state_dict``-style mapping using the
reference's key names, so one mapping drives the reference module,
this oracle and the CUDA model.

Third-party arithmetic: the text tower lives in HuggingFace ``transformers`` (reference
pins 4.2.1 in environment.yml:60; not vendored, not installable offline).  Its DistilBERT
graph is restated in ``distilbert_forward`` from the published architecture and pinned
against transformers 5.5.0's ``DistilBertModel`` (same graph).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# video tower  (model/video_transformer.py)
# ----------------------------------------------------------------------------------------------

def _linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def _softmax_av(q, k, v):
    """softmax(q k^T) v over the last two dims — video_transformer.py:29-33 (`attn`)."""
    return torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v


def divided_attention_core(qkv, heads, frames, patches, mode, scale_q=True):
    """Attention core of VarAttention.forward (video_transformer.py:104-133) without einops.

    qkv: [B, 1+frames*patches, 3*D] = the qkv Linear's output (token 0 = CLS, then frame-major patches).
    mode 'time'  : each (b, head, patch) attends over its `frames` tokens (+ CLS key/value)
    mode 'space' : each (b, head, frame) attends over its `patches` tokens (+ CLS key/value)
    The CLS query attends over all 1+frames*patches keys in both modes (:112).
    Returns the head-merged [B, S, D] tensor that feeds `proj`.
    """
    B, S, D3 = qkv.shape
    D = D3 // 3
    d = D // heads
    qkv = qkv.reshape(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)                  # 3,B,h,S,d (:104)
    q, k, v = qkv[0] * (d ** -0.5 if scale_q else 1.0), qkv[1], qkv[2]            # :106 (scale on q incl. CLS)
    cls_out = _softmax_av(q[:, :, :1], k, v)                                      # :112  [B,h,1,d]

    def group(t):  # [B,h,frames*patches,d] -> groups
        t = t.reshape(B, heads, frames, patches, d)
        return t.transpose(2, 3) if mode == "time" else t                          # :114

    qg, kg, vg = group(q[:, :, 1:]), group(k[:, :, 1:]), group(v[:, :, 1:])
    n_groups = qg.shape[2]
    kc = k[:, :, :1].unsqueeze(2).expand(B, heads, n_groups, 1, d)                # :117-118
    vc = v[:, :, :1].unsqueeze(2).expand(B, heads, n_groups, 1, d)
    og = _softmax_av(qg, torch.cat([kc, kg], dim=3), torch.cat([vc, vg], dim=3))   # :120-124
    if mode == "time":
        og = og.transpose(2, 3)                                                   # :127
    out = torch.cat([cls_out, og.reshape(B, heads, frames * patches, d)], dim=2)   # :130
    return out.permute(0, 2, 1, 3).reshape(B, S, D)                               # :133


def divided_attention(x, p, prefix, heads, frames, patches, mode):
    """VarAttention.forward (video_transformer.py:100-137): qkv Linear -> attention core -> proj Linear."""
    qkv = _linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])            # :103
    out = divided_attention_core(qkv, heads, frames, patches, mode)
    return _linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"])       # :135


def space_time_block(x, p, prefix, heads, frames, patches, eps=1e-6):
    """SpaceTimeBlock.forward (video_transformer.py:163-177)."""
    D = x.shape[-1]

    def ln(t, name):
        return F.layer_norm(t, (D,), p[prefix + name + ".weight"], p[prefix + name + ".bias"], eps)

    t_out = divided_attention(ln(x, "norm3"), p, prefix + "timeattn.", heads, frames, patches, "time")   # :166
    t_res = x + t_out                                                                                  # :167
    s_out = divided_attention(ln(t_res, "norm1"), p, prefix + "attn.", heads, frames, patches, "space")  # :168
    s_res = x + s_out                      # residual from the block INPUT (:171)
    h = _linear(ln(s_res, "norm2"), p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"])
    h = F.gelu(h)                                                                                       # exact erf (:41,48)
    return s_res + _linear(h, p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])                # :175


def video_tokens(video, p, prefix="video_model."):
    """Patch embedding + CLS + positional/temporal embeddings (video_transformer.py:72-77, 302-321)."""
    B, T, C, H, W = video.shape
    w = p[prefix + "patch_embed.proj.weight"]
    D, _, ph, pw = w.shape
    gh, gw = H // ph, W // pw
    n = gh * gw
    assert T <= p[prefix + "temporal_embed"].shape[1]                                                   # :74
    patches = video.reshape(B * T, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B * T * n, C * ph * pw)
    tok = _linear(patches, w.reshape(D, -1), p[prefix + "patch_embed.proj.bias"]).reshape(B, T * n, D)  # conv k16 s16
    pos = p[prefix + "pos_embed"][0]                  # [1+n, D]
    tmp = p[prefix + "temporal_embed"][0, :T]         # [T, D]   (first T frames' embeds, :319)
    tok = tok + (pos[1:].unsqueeze(0) + tmp.unsqueeze(1)).reshape(1, T * n, D)                          # :312-320
    cls = (p[prefix + "cls_token"][0, 0] + pos[0]).expand(B, 1, D)                                      # :308-316
    return torch.cat([cls, tok], dim=1), T, n


def video_tower(video, p, heads=12, prefix="video_model.", eps=1e-6):
    """SpaceTimeTransformer.forward_features (video_transformer.py:302-333): returns [B, D] CLS feature."""
    x, T, n = video_tokens(video, p, prefix)
    depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in p if k.startswith(prefix + "blocks."))
    for i in range(depth):
        x = space_time_block(x, p, f"{prefix}blocks.{i}.", heads, T, n, eps)
    D = x.shape[-1]
    return F.layer_norm(x, (D,), p[prefix + "norm.weight"], p[prefix + "norm.bias"], eps)[:, 0]          # :330


# ----------------------------------------------------------------------------------------------
# text tower  (transformers DistilBertModel; call sites model/model.py:32-36,119-122)
# ----------------------------------------------------------------------------------------------

def distilbert_forward(input_ids, attention_mask, p, heads=12, prefix="text_model.", eps=1e-12, dropout=None):
    """DistilBERT encoder: returns last_hidden_state [B, L, D].  `dropout` = None (eval / p = 0) or a dict of
    MULTIPLIERS (mask / (1 - p)) for HF's three train-mode dropout sites (modeling_distilbert.py): "emb" [B, L, D] on the
    embedding LayerNorm output, ("att", layer) [B, H, L, L] on the attention probabilities, ("ffn", layer) [B, L, D] on
    the FFN output -- the masks themselves are the RNG's business, the arithmetic around them is the oracle's."""
    B, L = input_ids.shape
    we = p[prefix + "embeddings.word_embeddings.weight"]
    pe = p[prefix + "embeddings.position_embeddings.weight"]
    D = we.shape[1]
    d = D // heads
    x = we[input_ids] + pe[:L].unsqueeze(0)
    x = F.layer_norm(x, (D,), p[prefix + "embeddings.LayerNorm.weight"], p[prefix + "embeddings.LayerNorm.bias"], eps)
    if dropout is not None:
        x = x * dropout["emb"]
    key_bias = torch.zeros(B, 1, 1, L, dtype=x.dtype, device=x.device)
    key_bias = key_bias.masked_fill(attention_mask.reshape(B, 1, 1, L) == 0, float("-inf"))
    n_layers = 1 + max(int(k.split("transformer.layer.")[1].split(".")[0]) for k in p if "transformer.layer." in k)
    for i in range(n_layers):
        lp = f"{prefix}transformer.layer.{i}."

        def heads_of(t):
            return t.reshape(B, L, heads, d).transpose(1, 2)

        q = heads_of(_linear(x, p[lp + "attention.q_lin.weight"], p[lp + "attention.q_lin.bias"])) / math.sqrt(d)
        k = heads_of(_linear(x, p[lp + "attention.k_lin.weight"], p[lp + "attention.k_lin.bias"]))
        v = heads_of(_linear(x, p[lp + "attention.v_lin.weight"], p[lp + "attention.v_lin.bias"]))
        w = torch.softmax(q @ k.transpose(-1, -2) + key_bias, dim=-1)
        if dropout is not None:
            w = w * dropout[("att", i)]
        ctx = (w @ v).transpose(1, 2).reshape(B, L, D)
        sa = _linear(ctx, p[lp + "attention.out_lin.weight"], p[lp + "attention.out_lin.bias"])
        x = F.layer_norm(sa + x, (D,), p[lp + "sa_layer_norm.weight"], p[lp + "sa_layer_norm.bias"], eps)
        h = F.gelu(_linear(x, p[lp + "ffn.lin1.weight"], p[lp + "ffn.lin1.bias"]))
        h = _linear(h, p[lp + "ffn.lin2.weight"], p[lp + "ffn.lin2.bias"])
        if dropout is not None:
            h = h * dropout[("ffn", i)]
        x = F.layer_norm(h + x, (D,), p[lp + "output_layer_norm.weight"], p[lp + "output_layer_norm.bias"], eps)
    return x


# ----------------------------------------------------------------------------------------------
# dual encoder  (model/model.py)
# ----------------------------------------------------------------------------------------------

def compute_text(text, p, heads=12, dropout=None):
    """FrozenInTime.compute_text (model/model.py:117-126): CLS -> ReLU -> Linear."""
    h = distilbert_forward(text["input_ids"], text["attention_mask"], p, heads, dropout=dropout)[:, 0]
    return _linear(torch.relu(h), p["txt_proj.1.weight"], p["txt_proj.1.bias"])


def compute_text_tokens(text, p, heads=12, dropout=None):
    """FrozenInTime.compute_text_tokens (model/model.py:128-138)."""
    h = distilbert_forward(text["input_ids"], text["attention_mask"], p, heads, dropout=dropout)
    return _linear(torch.relu(h), p["txt_proj.1.weight"], p["txt_proj.1.bias"])


def compute_video(video, p, heads=12):
    """FrozenInTime.compute_video (model/model.py:140-143)."""
    return _linear(video_tower(video, p, heads), p["vid_proj.0.weight"], p["vid_proj.0.bias"])


def frozen_in_time_forward(data, p, heads=12, video_only=False, return_embeds=True):
    """FrozenInTime.forward (model/model.py:100-115); returns (text, video) in that order."""
    if video_only:
        return compute_video(data["video"], p, heads)
    t = compute_text(data["text"], p, heads)
    v = compute_video(data["video"], p, heads)
    return (t, v) if return_embeds else sim_matrix(t, v)


def sim_matrix(a, b, eps=1e-8):
    """model/model.py:189-197: cosine similarity with the norm clamped at eps."""
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    return an @ bn.t()


# ----------------------------------------------------------------------------------------------
# losses  (model/loss.py)
# ----------------------------------------------------------------------------------------------

def norm_softmax_loss(x, temperature=0.05):
    """NormSoftmaxLoss.forward (model/loss.py:13-25)."""
    i = torch.log_softmax(x / temperature, dim=1).diagonal().mean()
    j = torch.log_softmax(x.t() / temperature, dim=1).diagonal().mean()
    return -i - j


def egonce_loss(x, sim_v, sim_n, temperature=0.05, noun=True, verb=True):
    """EgoNCE.forward (model/loss.py:34-53).  Positives: diagonal plus pairs whose
    verb AND noun similarity products are > 0 (variants via the noun/verb flags, :36-41).
    Note the column term reuses the un-transposed mask (:50), restated as is."""
    eye = torch.eye(x.shape[0], dtype=x.dtype, device=x.device)
    if noun and verb:
        mask = sim_v * sim_n + eye
    elif noun:
        mask = sim_n + eye
    else:
        mask = sim_v + eye
    pos = (mask > 0).to(x.dtype)
    i = torch.log((torch.softmax(x / temperature, dim=1) * pos).sum(1)).mean()
    j = torch.log((torch.softmax(x.t() / temperature, dim=1) * pos).sum(1)).mean()
    return -i - j


def max_margin_ranking_loss(x, margin=0.2, fix_norm=True):
    """MaxMarginRankingLoss.forward (model/loss.py:63-90)."""
    n = x.shape[0]
    d = x.diagonal().unsqueeze(1)
    h = torch.relu(margin - (d - x)) + torch.relu(margin - (d - x.t()))   # [i,j]: rows, and columns of x
    if fix_norm:
        off = 1.0 - torch.eye(n, dtype=x.dtype, device=x.device)
        return (h * off).sum() / (2 * n * (n - 1))
    return h.sum() / (2 * n * n)


def adaptive_max_margin_ranking_loss(x, weight, margin=0.4, fix_norm=True):
    """AdaptiveMaxMarginRankingLoss.forward (model/loss.py:100-133): anchor i's margin is weight[i] * margin in both
    the row (x_ij) and the column (x_ji) terms."""
    n = x.shape[0]
    d = x.diagonal().unsqueeze(1)
    m = weight.reshape(n, 1) * margin
    h = torch.relu(m - (d - x)) + torch.relu(m - (d - x.t()))
    if fix_norm:
        off = 1.0 - torch.eye(n, dtype=x.dtype, device=x.device)
        return (h * off).sum() / (2 * n * (n - 1))
    return h.sum() / (2 * n * n)


# ----------------------------------------------------------------------------------------------
# EPIC-Kitchens MIR ranking metrics (numpy, float64 accumulation like the reference)
# ----------------------------------------------------------------------------------------------

def _rank_desc(sim, tie_hi):
    """Column order of every row, best first.  tie_hi=True: stable ascending argsort reversed (utils/nDCG.py:32);
    tie_hi=False: stable argsort of -sim (utils/mAP.py:25).  (The reference's default argsort kind is not stable;
    on tie-free similarities all variants agree.)"""
    import numpy as np
    if tie_hi:
        return np.argsort(sim, axis=1, kind="stable")[:, ::-1]
    return np.argsort(-sim, axis=1, kind="stable")


def k_counts_of(relevancy):
    """utils/nDCG.py:47-75 calculate_k_counts: rank i counts iff i < #(relevancy[row] > 0)."""
    import numpy as np
    k = (relevancy > 0).sum(axis=1, keepdims=True)
    return (np.arange(relevancy.shape[1])[None, :] < k).astype(int)


def dcg(sim, relevancy, k_counts):
    """utils/nDCG.py:3-45 calculate_DCG."""
    import numpy as np
    ranks = _rank_desc(np.asarray(sim), True)
    rows = np.arange(sim.shape[0])[:, None]
    num = np.asarray(relevancy, dtype=np.float64)[rows, ranks] * k_counts
    return (num / np.log2(np.arange(sim.shape[1]) + 2.0)[None, :]).sum(axis=1)


def ndcg(sim, relevancy, k_counts=None, idcg=None, reduction="mean"):
    """utils/nDCG.py:96-139 calculate_nDCG (IDCG = DCG of the relevancy ranked by itself, :78-94)."""
    import numpy as np
    if k_counts is None:
        k_counts = k_counts_of(relevancy)
    d = dcg(sim, relevancy, k_counts)
    if idcg is None:
        idcg = dcg(relevancy, relevancy, k_counts)
    return np.mean(d / idcg) if reduction == "mean" else d / idcg


def average_precision(sim, relevancy):
    """utils/mAP.py:4-44 calculate_mAP, per query (the reference returns the mean of this vector)."""
    import numpy as np
    order = _rank_desc(np.asarray(sim), False)
    rows = np.arange(sim.shape[0])[:, None]
    rel = np.asarray(relevancy, dtype=np.float64)[rows, order]
    cum = np.cumsum(rel, axis=1)
    cum[rel != 1] = 0
    with np.errstate(invalid="ignore", divide="ignore"):
        return (cum / (np.arange(rel.shape[1]) + 1.0)).sum(axis=1) / (rel == 1).sum(axis=1)


def dual_softmax(sim):
    """run/test_epic.py:137-143: s = softmax(s/500, dim=1) * s ; s = softmax(s, dim=0)."""
    s = torch.softmax(sim / 500.0, dim=1) * sim
    return torch.softmax(s, dim=0)


def egomcq_predict(text_embeds, video_embeds):
    """trainer/trainer_egoclip.py:204-215 + model/metric.py:227: per query cosine of 1 text vs
    K candidate videos, prediction = argmax (ties -> lowest index).
    text_embeds [Q, C]; video_embeds [Q, K, C] -> (scores [Q, K], pred [Q] int64)."""
    t = text_embeds / text_embeds.norm(dim=1, keepdim=True).clamp_min(1e-8)
    v = video_embeds / video_embeds.norm(dim=2, keepdim=True).clamp_min(1e-8)
    s = torch.einsum("qc,qkc->qk", t, v)
    return s, torch.argmax(s, dim=1)


def gathered_step_loss(text_local, video_local, verb_local, noun_local, temperature=0.05):
    """Single-process equivalent of trainer/trainer_egoclip.py:125-135 for a list of per-rank
    local tensors: concatenate (== all_gather + cat), sim_matrix x3, EgoNCE."""
    t, v = torch.cat(text_local), torch.cat(video_local)
    vb, nn_ = torch.cat(verb_local), torch.cat(noun_local)
    return egonce_loss(sim_matrix(t, v), sim_matrix(vb, vb), sim_matrix(nn_, nn_), temperature)
