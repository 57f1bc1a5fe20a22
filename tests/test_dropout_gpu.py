"""Train-mode dropout of the DistilBERT text tower (HF modeling_distilbert.py; the reference keeps
`text_model.train()`, model/model.py:36).  RNG parity with torch is impossible, so the test separates the two halves:
the MASKS are checked as an RNG (rate, scaling, determinism, site independence), and the ARITHMETIC around them is
checked against the fp32 oracle fed with the very masks the kernels drew (extracted through the C-ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dropout_kernel_is_a_proper_mask():
    from egovlp_b200 import ops
    n, p, seed = 1 << 20, 0.1, 1234567891011
    ones = torch.ones(n, device="cuda")
    y, y16 = ops.dropout(ones, p, seed, 3, y32=torch.empty_like(ones), y16=torch.empty(n, device="cuda", dtype=torch.bfloat16))
    kept = y != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 3e-3                     # 1M draws: sigma = 3e-4
    torch.testing.assert_close(y[kept], torch.full_like(y[kept], 1 / (1 - p)))
    assert torch.equal(y16, y.bfloat16())
    y2, _ = ops.dropout(ones, p, seed, 3, y32=torch.empty_like(ones))
    assert torch.equal(y, y2)                                                    # same (seed, site) -> same mask
    other_site, _ = ops.dropout(ones, p, seed, 4, y32=torch.empty_like(ones))
    other_seed, _ = ops.dropout(ones, p, seed + 1, 3, y32=torch.empty_like(ones))
    for o in (other_site, other_seed):                                           # independent streams
        agree = ((o != 0) == kept).float().mean().item()
        assert abs(agree - (0.81 + 0.01)) < 5e-3
    # no visible structure along the element index: every 4096-block keeps ~90 %
    blocks = kept.view(-1, 4096).float().mean(1)
    assert (blocks - 0.9).abs().max().item() < 0.03
    x, add = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    z, _ = ops.dropout(x, p, seed, 3, add=add, y32=torch.empty_like(x))
    torch.testing.assert_close(z, x * y + add)
    same, _ = ops.dropout(x, 0.0, seed, 3, y32=torch.empty_like(x))
    assert torch.equal(same, x)


def _extract_attention_multiplier(B, L, H, p, seed, site):
    """q = k = 0 gives uniform probabilities 1/L; v = identity (one-hot of the key index, L <= 64) makes the output
    row i equal the dropped probability row -> multiplier[b, h, i, j] = out * L."""
    from egovlp_b200 import ops
    D = 64 * H
    qkv = torch.zeros(B * L, 3 * D, device="cuda", dtype=torch.bfloat16)
    eye = torch.zeros(L, 64, device="cuda")
    eye[torch.arange(L), torch.arange(L)] = 1.0
    for h in range(H):
        qkv.view(B, L, 3 * D)[:, :, 2 * D + 64 * h: 2 * D + 64 * (h + 1)] = eye.bfloat16()
    out = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
    ops.text_attn_fwd(qkv, torch.ones(B, L, dtype=torch.int64, device="cuda"), out, B, L, H, p, seed, site)
    probs = out.float().view(B, L, H, 64)[..., :L].permute(0, 2, 1, 3) * L      # [B, H, L(i), L(j)]
    return torch.where(probs > 0.5, torch.full_like(probs, 1 / (1 - p)), torch.zeros_like(probs))


def test_attention_dropout_matches_torch_with_the_same_mask():
    from egovlp_b200 import ops
    B, L, H, p, seed, site = 3, 48, 2, 0.25, 987654321, 5
    D = 64 * H
    mult = _extract_attention_multiplier(B, L, H, p, seed, site)
    assert abs((mult > 0).float().mean().item() - (1 - p)) < 0.02
    g = torch.Generator().manual_seed(2)
    qkv = (torch.randn(B * L, 3 * D, generator=g) * 0.5).cuda().bfloat16()
    lens = torch.tensor([L, L - 7, 5])
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int64).cuda()
    dout = torch.randn(B * L, D, generator=g).cuda().bfloat16()
    out = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
    ops.text_attn_fwd(qkv, mask, out, B, L, H, p, seed, site)
    dqkv = torch.empty_like(qkv)
    ops.text_attn_bwd(qkv, mask, dout, dqkv, B, L, H, 1.0, p, seed, site)
    x = qkv.float().requires_grad_(True)
    q, k, v = (x.view(B, L, 3, H, 64)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    bias = torch.zeros(B, 1, 1, L, device="cuda").masked_fill(mask.view(B, 1, 1, L) == 0, float("-inf"))
    w = torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1) * mult
    ref = (w @ v).permute(0, 2, 1, 3).reshape(B * L, D)
    ref.backward(dout.float())
    torch.testing.assert_close(out.float(), ref.detach(), rtol=2e-2, atol=2e-3)        # bf16 outputs
    torch.testing.assert_close(dqkv.float(), x.grad, rtol=2e-2, atol=4e-3)


def test_text_tower_with_dropout_vs_oracle_fed_with_the_drawn_masks():
    from egovlp_b200 import engine, ops, synthetic as syn
    from oracle import reference_port as rp
    from test_model_gpu import rel, cos
    dims = syn.TINY_DIMS
    sd = {k: v for k, v in syn.seeded_state_dict(dims, seed=4, video=False, proj=True).items() if not k.startswith("vid_proj")}
    text = syn.synthetic_text(5, 9, seed=1, ragged=True, vocab=120)
    B, L, D, H, p_hid, p_att = 5, 9, dims["text_dim"], dims["text_heads"], 0.1, 0.2
    order = ["text_model.embeddings.word_embeddings.weight", "text_model.embeddings.position_embeddings.weight",
             "text_model.embeddings.LayerNorm.weight", "text_model.embeddings.LayerNorm.bias"]
    for i in range(dims["text_layers"]):
        lp = f"text_model.transformer.layer.{i}."
        for lin in ("attention.q_lin", "attention.k_lin", "attention.v_lin", "attention.out_lin"):
            order += [lp + lin + ".weight", lp + lin + ".bias"]
        order += [lp + "sa_layer_norm.weight", lp + "sa_layer_norm.bias", lp + "ffn.lin1.weight", lp + "ffn.lin1.bias",
                  lp + "ffn.lin2.weight", lp + "ffn.lin2.bias", lp + "output_layer_norm.weight", lp + "output_layer_norm.bias"]
    order += ["txt_proj.1.weight", "txt_proj.1.bias"]
    p_gpu = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    ids, mask = text["input_ids"].cuda(), text["attention_mask"].cuda()
    cache = engine.Bf16Cache()

    def run(seed_for_torch):
        torch.manual_seed(seed_for_torch)
        return engine.TextTowerFn.apply(ids, mask, H, 1e-12, False, cache, (p_hid, p_att), *[p_gpu[k] for k in order])

    det = engine.TextTowerFn.apply(ids, mask, H, 1e-12, False, cache, None, *[p_gpu[k] for k in order])
    a, b, c = run(7), run(7), run(8)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, det)     # reproducible, seeded, active
    # the seed the forward drew, and from it the multipliers of every site
    torch.manual_seed(7)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    ones = torch.ones(B * L * D, device="cuda")
    drop = {"emb": ops.dropout(ones, p_hid, seed, 0, y32=torch.empty_like(ones))[0].view(B, L, D).cpu()}
    for i in range(dims["text_layers"]):
        drop[("att", i)] = _extract_attention_multiplier(B, L, H, p_att, seed, 1 + 2 * i).cpu()
        drop[("ffn", i)] = ops.dropout(ones, p_hid, seed, 2 + 2 * i, y32=torch.empty_like(ones))[0].view(B, L, D).cpu()
    p_cpu = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = rp.compute_text(text, p_cpu, heads=H, dropout=drop)
    assert rel(a, want) < 1e-2, rel(a, want)
    assert rel(det, want) > 5e-2                                                       # and it is not the p = 0 output
    probe = torch.randn(want.shape, generator=torch.Generator().manual_seed(3))
    (want * probe).sum().backward()
    (a * probe.cuda()).sum().backward()
    for k in order:
        ref = p_cpu[k].grad
        if ref is None or k.endswith("k_lin.bias"):
            continue
        assert cos(p_gpu[k].grad, ref) > 0.995, (k, cos(p_gpu[k].grad, ref))
