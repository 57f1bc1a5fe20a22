"""Tensor-level wrappers over the C-ABI (include/egovlp_b200.h).  torch is used only for device memory and
the current stream; every computation happens in libegovlp_b200.so."""
import ctypes as C

import torch

from ._lib import GemmEpilogue, call, lib

BF16, F32 = torch.bfloat16, torch.float32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype, f"{name}: expected cuda {dtype}, got {t.device} {t.dtype}"


_prof = None         # bench.py's live roofline probe: {kind: [(work, start event, end event), ...]}


def profile(enable):
    """CUDA events (on the launching stream) around every GEMM / divided-attention / LayerNorm launch.
    profile(True) starts recording; profile(False) returns {kind: (work, milliseconds, launches)} where `work` is the
    algorithmic FLOPs (kind 'gemm') or algorithmic HBM bytes (the HBM-bound kinds) of the recorded launches."""
    global _prof
    if enable:
        _prof = {}
        return None
    rec, _prof = _prof or {}, None
    torch.cuda.synchronize()
    return {k: (sum(r[0] for r in v), sum(r[1].elapsed_time(r[2]) for r in v), len(v)) for k, v in rec.items()}


def profile_gemm(enable):
    """GEMM-only view of `profile` (tools/): profile_gemm(False) -> (flops, milliseconds, launches)."""
    res = profile(enable)
    return None if enable else res.get("gemm", (0.0, 0.0, 0))


class _Probe:
    """`with _Probe(kind, work):` records one launch when profiling is on (no-op otherwise)."""
    __slots__ = ("kind", "work", "e0")

    def __init__(self, kind, work):
        self.kind, self.work, self.e0 = kind, work, None

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.setdefault(self.kind, []).append((self.work, self.e0, e1))
        return False


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, residual=None, aux=None, out2=None, act=0, alpha=1.0,
         col_scale=1.0, col_scale_ncols=0, accumulate=False, split_k=1, res_row_mod=0, colsum=None, colsum_a=None):
    """out = epi(A @ B^T).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); 2-D, last-dim contiguous.
    out: bf16 or fp32 [M,N]; accumulate=True -> fp32 atomic add into `out` (required for split_k>1)."""
    _chk(a, BF16, "a"); _chk(b, BF16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    (K, M) = a.shape if a_mn else a.shape[::-1]
    (Kb, N) = b.shape if b_mn else b.shape[::-1]
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (BF16, F32)
    e = GemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.aux = aux.data_ptr() if aux is not None else None
    e.out = out.data_ptr()
    e.out2 = out2.data_ptr() if out2 is not None else None
    e.ldr = residual.stride(0) if residual is not None else 0
    e.ldaux = aux.stride(0) if aux is not None else 0
    e.ldo = out.stride(0)
    e.ldo2 = out2.stride(0) if out2 is not None else 0
    if bias is not None:
        _chk(bias, F32, "bias"); assert bias.numel() == N
    if residual is not None:
        _chk(residual, F32, "residual")
        assert residual.shape == ((res_row_mod or M), N) and residual.stride(1) == 1
    if aux is not None:
        _chk(aux, BF16, "aux"); assert aux.shape == (M, N)
    if out2 is not None:
        _chk(out2, BF16, "out2"); assert out2.shape == (M, N)
    if accumulate:
        assert out.dtype == F32
        e.out_mode = 2
    else:
        e.out_mode = 0 if out.dtype == BF16 else 1
    e.act, e.alpha, e.col_scale, e.col_scale_ncols = act, alpha, col_scale, col_scale_ncols
    e.res_row_mod = res_row_mod
    if colsum is not None:
        _chk(colsum, F32, "colsum"); assert colsum.numel() == N
    e.colsum = colsum.data_ptr() if colsum is not None else None
    if colsum_a is not None:                 # wgrad form only: column sums of A accumulated from the smem tiles
        _chk(colsum_a, F32, "colsum_a"); assert a_mn and b_mn and colsum_a.numel() == M
    e.colsum_a = colsum_a.data_ptr() if colsum_a is not None else None
    with _Probe("gemm", 2.0 * M * N * K):
        call("egovlp_gemm_bf16", _ptr(a), int(a_mn), C.c_longlong(a.stride(0)), _ptr(b), int(b_mn),
             C.c_longlong(b.stride(0)), M, N, K, C.byref(e), split_k, _stream())
    return out


def layernorm_fwd(x, gamma, beta, eps, *, add=None, sum_out=None, y16=None, y32=None, mean=None, rstd=None):
    """x fp32 [rows, D] (row stride free).  Returns nothing; writes the provided outputs."""
    _chk(x, F32, "x")
    rows, D = x.shape
    assert x.stride(1) == 1
    for t in (add, sum_out, y32):
        assert t is None or (t.dtype == F32 and t.is_contiguous() and t.shape == (rows, D))
    assert y16 is None or (y16.dtype == BF16 and y16.is_contiguous() and y16.shape == (rows, D))
    nbytes = rows * D * (4 + 4 * (add is not None) + 4 * (sum_out is not None) + 2 * (y16 is not None) + 4 * (y32 is not None))
    with _Probe("layernorm_fwd", nbytes):
        call("egovlp_layernorm_fwd", _ptr(x), C.c_longlong(x.stride(0)), _ptr(add), _ptr(sum_out), _ptr(gamma),
             _ptr(beta), _ptr(y16), _ptr(y32), _ptr(mean), _ptr(rstd), rows, D, C.c_float(eps), _stream())


def layernorm_bwd(dy, x, gamma, mean, rstd, *, add1=None, add2=None, dx=None, dx16=None, dgamma=None, dbeta=None,
                  colsum_dx=None):
    """dx = LNbwd(dy) + add1 + add2.  dy / add1 / add2 may each be fp32 or bf16."""
    _chk(x, F32, "x")
    rows, D = x.shape
    assert dy.is_cuda and dy.dtype in (F32, BF16) and dy.shape == (rows, D) and dy.stride(1) == 1 and x.stride(1) == 1
    for t in (add1, add2):
        assert t is None or (t.dtype in (F32, BF16) and t.is_contiguous() and t.shape == (rows, D))
    assert dx is None or (dx.dtype == F32 and dx.shape == (rows, D) and dx.stride(1) == 1)
    assert dx16 is None or (dx16.dtype == BF16 and dx16.is_contiguous() and dx16.shape == (rows, D))
    nbytes = rows * D * (dy.element_size() + 4 + sum(t.element_size() for t in (add1, add2) if t is not None)
                         + 4 * (dx is not None) + 2 * (dx16 is not None))
    with _Probe("layernorm_bwd", nbytes):
        call("egovlp_layernorm_bwd", _ptr(dy), int(dy.dtype == BF16), C.c_longlong(dy.stride(0)), _ptr(x),
             C.c_longlong(x.stride(0)), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(add1),
             int(add1 is not None and add1.dtype == BF16), _ptr(add2), int(add2 is not None and add2.dtype == BF16),
             _ptr(dx), C.c_longlong(dx.stride(0) if dx is not None else D), _ptr(dx16), _ptr(dgamma), _ptr(dbeta),
             _ptr(colsum_dx), rows, D, _stream())


def cast_bf16(src, dst=None):
    _chk(src, F32, "src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    call("egovlp_cast_f32_to_bf16", _ptr(src), _ptr(dst), C.c_longlong(src.numel()), _stream())
    return dst


class _CastDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("numel", C.c_longlong)]


def build_cast_table(pairs):
    """Device-side descriptor / chunk tables for `cast_multi`: pairs = [(fp32 src, bf16 dst), ...] (contiguous)."""
    chunk = lib().egovlp_adamw_chunk_elems()
    descs = (_CastDesc * len(pairs))()
    ct, co = [], []
    for i, (src, dst) in enumerate(pairs):
        _chk(src, F32, "src"); _chk(dst, BF16, "dst")
        assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
        descs[i] = _CastDesc(src.data_ptr(), dst.data_ptr(), src.numel())
        n = (src.numel() + chunk - 1) // chunk
        ct += [i] * n
        co += list(range(n))
    dev = pairs[0][0].device
    raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
    return raw, torch.tensor(ct, dtype=torch.int32, device=dev), torch.tensor(co, dtype=torch.int32, device=dev)


def cast_multi(raw, chunk_tensor, chunk_offset):
    """fp32 -> bf16 for every (src, dst) pair of a table from `build_cast_table`, one launch."""
    call("egovlp_cast_multi_f32_to_bf16", _ptr(raw), _ptr(chunk_tensor), _ptr(chunk_offset), chunk_tensor.numel(),
         _stream())


def colsum_accum(dy, out):
    """out[n] += sum_m dy[m,n]; dy bf16/fp32 [M,N]."""
    assert dy.dim() == 2 and dy.stride(1) == 1 and out.dtype == F32 and out.numel() == dy.shape[1]
    call("egovlp_colsum_accum", _ptr(dy), int(dy.dtype == F32), C.c_longlong(dy.stride(0)), _ptr(out), dy.shape[0],
         dy.shape[1], _stream())


def divided_attn_fwd(qkv, B, T, N, H, mode):
    """qkv bf16 [B*S, 3*64*H] (q pre-scaled) -> (out bf16 [B*S, D], lse fp32 [B, H, S]).  mode: 0 time, 1 space."""
    _chk(qkv, BF16, "qkv")
    S, D = 1 + T * N, 64 * H
    assert qkv.is_contiguous() and qkv.shape == (B * S, 3 * D)
    out = torch.empty(B * S, D, dtype=BF16, device=qkv.device)
    lse = torch.empty(B, H, S, dtype=F32, device=qkv.device)
    n_ws = lib().egovlp_divided_attn_workspace_floats(B, T, N, H, mode)
    assert n_ws > 0, "unsupported attention geometry"
    ws = torch.empty(n_ws, dtype=F32, device=qkv.device)
    # algorithmic bytes: read q|k|v once (3 x 2 D), write the output (2 D) and the lse (4 H) per token
    with _Probe("attn_time_fwd" if mode == 0 else "attn_space_fwd", B * S * (8 * D + 4 * H)):
        call("egovlp_divided_attn_fwd", _ptr(qkv), _ptr(out), _ptr(lse), _ptr(ws), B, T, N, H, mode, _stream())
    return out, lse


def divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, mode, q_scale, dqkv=None):
    _chk(qkv, BF16, "qkv"); _chk(out, BF16, "out"); _chk(dout, BF16, "dout"); _chk(lse, F32, "lse")
    assert qkv.is_contiguous() and out.is_contiguous() and dout.is_contiguous() and lse.is_contiguous()
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    ws = torch.empty(B * H * 3 * 64, dtype=F32, device=qkv.device)
    S, D = 1 + T * N, 64 * H
    # algorithmic bytes: read q|k|v, out, dout and the lse once, write dq|dk|dv
    with _Probe("attn_time_bwd" if mode == 0 else "attn_space_bwd", B * S * (16 * D + 4 * H)):
        call("egovlp_divided_attn_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(ws), B, T, N, H,
             mode, C.c_float(q_scale), _stream())
    return dqkv


# ---------------------------------------------------------------- video front end
def patch_im2col(video, patches, P):
    _chk(video, F32, "video")
    B, T, Cc, H, W = video.shape
    assert video.is_contiguous() and patches.dtype == BF16 and patches.is_contiguous()
    call("egovlp_patch_im2col", _ptr(video), _ptr(patches), B, T, Cc, H, W, P, _stream())


def patch_im2col_u8(video, patches, P, mean, std):
    assert video.is_cuda and video.dtype == torch.uint8 and video.is_contiguous()
    B, T, Cc, H, W = video.shape
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    call("egovlp_patch_im2col_u8", _ptr(video), _ptr(patches), B, T, Cc, H, W, P, m3, s3, _stream())


def video_pos_table(cls_token, pos_embed, temporal_embed, conv_bias, table, T, N, D):
    call("egovlp_video_pos_table", _ptr(cls_token), _ptr(pos_embed), _ptr(temporal_embed), _ptr(conv_bias),
         _ptr(table), T, N, D, _stream())


def video_embed_bwd(dx, tmp, dcls, dpos, dtemporal, dbias, B, T, N, D):
    call("egovlp_video_embed_bwd", _ptr(dx), _ptr(tmp), _ptr(dcls), _ptr(dpos), _ptr(dtemporal), _ptr(dbias), B, T, N,
         D, _stream())


# ---------------------------------------------------------------- text tower pieces
def text_embed_fwd(ids, word, pos, out, B, L, D):
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    call("egovlp_text_embed_fwd", _ptr(ids), _ptr(word), _ptr(pos), _ptr(out), B, L, D, _stream())


def text_embed_bwd(ids, dsum, dword, dpos, B, L, D):
    call("egovlp_text_embed_bwd", _ptr(ids), _ptr(dsum), _ptr(dword), _ptr(dpos), B, L, D, _stream())


def text_attn_fwd(qkv, mask, out, B, L, H, p_drop=0.0, seed=0, site=0):
    assert mask.dtype == torch.int64 and mask.is_contiguous()
    call("egovlp_text_attn_fwd", _ptr(qkv), _ptr(mask), _ptr(out), B, L, H, C.c_float(p_drop), C.c_ulonglong(seed),
         C.c_uint(site), _stream())


def text_attn_bwd(qkv, mask, dout, dqkv, B, L, H, q_scale, p_drop=0.0, seed=0, site=0):
    call("egovlp_text_attn_bwd", _ptr(qkv), _ptr(mask), _ptr(dout), _ptr(dqkv), B, L, H, C.c_float(q_scale),
         C.c_float(p_drop), C.c_ulonglong(seed), C.c_uint(site), _stream())


def dropout(x, p, seed, site, add=None, y32=None, y16=None):
    """y = dropout_p(x) (+ add) with the (seed, site) Philox mask; returns (y32, y16) (whichever were given)."""
    _chk(x, F32, "x")
    assert x.is_contiguous() and (y32 is not None or y16 is not None)
    call("egovlp_dropout", _ptr(x), _ptr(add), _ptr(y32), _ptr(y16), C.c_longlong(x.numel()), C.c_float(p),
         C.c_ulonglong(seed), C.c_uint(site), _stream())
    return y32, y16


def relu_rows_fwd(x, row_stride, out, rows, D):
    call("egovlp_relu_rows_fwd", _ptr(x), C.c_longlong(row_stride), _ptr(out), rows, D, _stream())


def relu_rows_bwd(x, row_stride, dh, dx, rows, D):
    call("egovlp_relu_rows_bwd", _ptr(x), C.c_longlong(row_stride), _ptr(dh), _ptr(dx), rows, D, _stream())


# ---------------------------------------------------------------- similarity / losses (fp32)
def rownorm_fwd(a, eps=1e-8):
    _chk(a, F32, "a")
    a = a.contiguous()
    an, norm = torch.empty_like(a), torch.empty(a.shape[0], dtype=F32, device=a.device)
    call("egovlp_rownorm_fwd", _ptr(a), _ptr(an), _ptr(norm), a.shape[0], a.shape[1], C.c_float(eps), _stream())
    return an, norm


def rownorm_bwd(dan, an, norm, eps=1e-8):
    da = torch.empty_like(an)
    call("egovlp_rownorm_bwd", _ptr(dan.contiguous()), _ptr(an), _ptr(norm), _ptr(da), an.shape[0], an.shape[1],
         C.c_float(eps), _stream())
    return da


def sgemm(a, b, *, trans_a=False, trans_b=True, out=None, alpha=1.0, beta=0.0):
    """fp32 C = op(a) @ op(b)^T-style product on CUDA cores.  Default: a[M,K] @ b[N,K]^T.
    trans_a: a is [K,M]; trans_b=False: b is [K,N]."""
    _chk(a, F32, "a"); _chk(b, F32, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    sam, sak = (a.stride(1), a.stride(0)) if trans_a else (a.stride(0), a.stride(1))
    sbn, sbk = (b.stride(0), b.stride(1)) if trans_b else (b.stride(1), b.stride(0))
    if out is None:
        out = torch.empty(M, N, dtype=F32, device=a.device)
    call("egovlp_sgemm_f32", _ptr(a), C.c_longlong(sam), C.c_longlong(sak), _ptr(b), C.c_longlong(sbn),
         C.c_longlong(sbk), _ptr(out), C.c_longlong(out.stride(0)), M, N, K, C.c_float(alpha), C.c_float(beta), _stream())
    return out


def positives_mask_from_tags(verb, noun, mode):
    """uint8 [G,G] positives mask (diag | shared verb & shared noun) from multi-hot vectors."""
    G = (verb if verb is not None else noun).shape[0]
    dev = (verb if verb is not None else noun).device
    mask = torch.empty(G, G, dtype=torch.uint8, device=dev)
    vb = nb = None
    nv = nn_ = 0
    if verb is not None:
        nv = verb.shape[1]
        vb = torch.empty(G, (nv + 31) // 32, dtype=torch.int32, device=dev)
        call("egovlp_pack_multihot", _ptr(verb.contiguous().float()), _ptr(vb), G, nv, _stream())
    if noun is not None:
        nn_ = noun.shape[1]
        nb = torch.empty(G, (nn_ + 31) // 32, dtype=torch.int32, device=dev)
        call("egovlp_pack_multihot", _ptr(noun.contiguous().float()), _ptr(nb), G, nn_, _stream())
    call("egovlp_mask_from_bits", _ptr(vb), nv, _ptr(nb), nn_, _ptr(mask), G, mode, _stream())
    return mask


def positives_mask_from_sims(sim_v, sim_n, G, mode, device=None):
    """uint8 [G,G] positives mask; mode 0 (identity) needs `device` since it has no input tensor to take it from."""
    src = sim_v if sim_v is not None else sim_n
    dev = src.device if src is not None else torch.device(device)
    assert dev.type == "cuda", f"positives mask: CUDA tensors expected, got {dev} (there is no CPU path)"
    mask = torch.empty(G, G, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        call("egovlp_mask_from_sims", _ptr(sim_v), _ptr(sim_n), _ptr(mask), G, mode, _stream())
    return mask


def nce_fwd(x, mask, inv_temp):
    _chk(x, F32, "x")
    G = x.shape[0]
    assert x.shape == (G, G) and x.is_contiguous() and mask.shape == (G, G)
    stats = torch.empty(4 * G, dtype=F32, device=x.device)
    loss = torch.empty((), dtype=F32, device=x.device)
    call("egovlp_nce_fwd", _ptr(x), _ptr(mask), G, C.c_float(inv_temp), _ptr(stats), _ptr(loss), _stream())
    return loss, stats


def nce_bwd(x, mask, stats, inv_temp, gscale):
    dx = torch.empty_like(x)
    call("egovlp_nce_bwd", _ptr(x), _ptr(mask), _ptr(stats), x.shape[0], C.c_float(inv_temp), _ptr(gscale), _ptr(dx),
         _stream())
    return dx


def pack_rows4(a, b, c, d):
    """[a | b | c | d] along dim 1 (fp32, row-major) in one launch: the send buffer of the packed all-gather."""
    for t in (a, b, c, d):
        _chk(t, F32, "pack_rows4 input"); assert t.dim() == 2 and t.is_contiguous() and t.shape[0] == a.shape[0]
    out = torch.empty(a.shape[0], a.shape[1] + b.shape[1] + c.shape[1] + d.shape[1], dtype=F32, device=a.device)
    call("egovlp_pack_rows4", _ptr(a), a.shape[1], _ptr(b), b.shape[1], _ptr(c), c.shape[1], _ptr(d), d.shape[1], _ptr(out),
         a.shape[0], _stream())
    return out


_fused_ws = {}


def egonce_fused_supported(G, C):
    return G <= lib().egovlp_egonce_fused_max_g() and C <= 256


def _rows_view(t, name):
    _chk(t, F32, name)
    assert t.dim() == 2 and t.stride(1) == 1, f"{name}: rows with unit inner stride expected"
    return t


def egonce_fused_fwd(text, video, verb, noun, inv_temp, mode, eps=1e-8):
    """ONE kernel: normalise rows -> cosine similarities (smem only) -> positives from tag bits -> masked LSEs -> loss.
    text / video [G, C], verb [G, nv] / noun [G, nn] (or None per `mode`): row-strided fp32 views, read in place.
    -> (loss, saved = (norm_text, norm_video, tag_bits, stats))."""
    text, video = _rows_view(text, "text"), _rows_view(video, "video")
    G, Cc = text.shape
    assert video.shape == (G, Cc)
    nv = verb.shape[1] if verb is not None else 0
    nn_ = noun.shape[1] if noun is not None else 0
    if verb is not None: _rows_view(verb, "verb")
    if noun is not None: _rows_view(noun, "noun")
    dev = text.device
    na, nb = torch.empty(G, dtype=F32, device=dev), torch.empty(G, dtype=F32, device=dev)
    bits = torch.empty(G, max(1, (nv + 31) // 32 + (nn_ + 31) // 32), dtype=torch.int32, device=dev)
    stats = torch.empty(4 * G, dtype=F32, device=dev)
    loss = torch.empty((), dtype=F32, device=dev)
    key = (dev.index, G)
    ws = _fused_ws.get(key)
    if ws is None:                      # zero once: the kernel leaves its ticket word zero after every launch
        ws = _fused_ws[key] = torch.zeros(lib().egovlp_egonce_fused_workspace_floats(G), dtype=F32, device=dev)
    call("egovlp_egonce_fused_fwd", _ptr(text), C.c_longlong(text.stride(0)), _ptr(video), C.c_longlong(video.stride(0)),
         _ptr(verb), C.c_longlong(verb.stride(0) if verb is not None else 0), nv, _ptr(noun),
         C.c_longlong(noun.stride(0) if noun is not None else 0), nn_, G, Cc, C.c_float(inv_temp), int(mode), C.c_float(eps),
         _ptr(na), _ptr(nb), _ptr(bits), _ptr(stats), _ptr(ws), _ptr(loss), _stream())
    return loss, (na, nb, bits, stats)


def egonce_fused_bwd(text, video, saved, n_verb, n_noun, inv_temp, mode, gscale, row0, n_local, eps=1e-8):
    """-> (d_text, d_video) [n_local, C] of rows [row0, row0 + n_local) only."""
    na, nb, bits, stats = saved
    G, Cc = text.shape
    d_text = torch.empty(n_local, Cc, dtype=F32, device=text.device)
    d_video = torch.empty(n_local, Cc, dtype=F32, device=text.device)
    call("egovlp_egonce_fused_bwd", _ptr(text), C.c_longlong(text.stride(0)), _ptr(video), C.c_longlong(video.stride(0)),
         _ptr(na), _ptr(nb), _ptr(bits), n_verb, n_noun, _ptr(stats), G, Cc, C.c_float(inv_temp), int(mode), C.c_float(eps),
         _ptr(gscale), row0, n_local, _ptr(d_text), _ptr(d_video), _stream())
    return d_text, d_video


def maxmargin_fwd(x, margin, fix_norm, row_weight=None):
    _chk(x, F32, "x")
    if row_weight is not None:
        _chk(row_weight, F32, "row_weight")
        assert row_weight.shape == (x.shape[0],) and row_weight.is_contiguous()
    loss = torch.empty((), dtype=F32, device=x.device)
    call("egovlp_maxmargin_fwd", _ptr(x), _ptr(row_weight), x.shape[0], C.c_float(margin), int(fix_norm), _ptr(loss),
         _stream())
    return loss


def maxmargin_bwd(x, margin, fix_norm, gscale, row_weight=None):
    dx = torch.empty_like(x)
    call("egovlp_maxmargin_bwd", _ptr(x), _ptr(row_weight), x.shape[0], C.c_float(margin), int(fix_norm), _ptr(gscale),
         _ptr(dx), _stream())
    return dx


def rank_metrics(sim, rel, k_counts=None, tie_mode=0, want_dcg=True, want_ap=True):
    """Per-query ranking metrics (egovlp_rank_metrics): sim fp32 [R, C], rel fp32/fp64 [R, C], optional int32
    k_counts [R, C] -> (dcg fp64 [R] | None, ap fp64 [R] | None)."""
    _chk(sim, F32, "sim")
    assert rel.is_cuda and rel.dtype in (torch.float32, torch.float64) and rel.shape == sim.shape and sim.dim() == 2
    sim, rel = sim.contiguous(), rel.contiguous()
    R, Cc = sim.shape
    if k_counts is not None:
        assert k_counts.is_cuda and k_counts.shape == sim.shape
        k_counts = k_counts.to(torch.int32).contiguous()
    dcg = torch.empty(R, dtype=torch.float64, device=sim.device) if want_dcg else None
    ap = torch.empty(R, dtype=torch.float64, device=sim.device) if want_ap else None
    if R == 0:
        return dcg, ap
    call("egovlp_rank_metrics", _ptr(sim), C.c_longlong(sim.stride(0)), _ptr(rel), int(rel.dtype == torch.float64),
         C.c_longlong(rel.stride(0)), _ptr(k_counts), R, Cc, int(tie_mode), _ptr(dcg), _ptr(ap), _stream())
    return dcg, ap


def dual_softmax(sim, temp=500.0):
    _chk(sim, F32, "sim")
    sim = sim.contiguous()
    out = torch.empty_like(sim)
    call("egovlp_dual_softmax", _ptr(sim), _ptr(out), sim.shape[0], sim.shape[1], C.c_float(temp), _stream())
    return out


def egomcq_score(text, video, eps=1e-8):
    _chk(text, F32, "text"); _chk(video, F32, "video")
    Q, K, Cc = video.shape
    scores = torch.empty(Q, K, dtype=F32, device=text.device)
    pred = torch.empty(Q, dtype=torch.int64, device=text.device)
    call("egovlp_egomcq_score", _ptr(text.contiguous()), _ptr(video.contiguous()), _ptr(scores), _ptr(pred), Q, K, Cc,
         C.c_float(eps), _stream())
    return scores, pred
