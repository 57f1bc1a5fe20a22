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


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, residual=None, aux=None, out2=None, act=0, alpha=1.0,
         col_scale=1.0, col_scale_ncols=0, accumulate=False, split_k=1):
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
        _chk(residual, F32, "residual"); assert residual.shape == (M, N)
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
    call("egovlp_layernorm_fwd", _ptr(x), C.c_longlong(x.stride(0)), _ptr(add), _ptr(sum_out), _ptr(gamma),
         _ptr(beta), _ptr(y16), _ptr(y32), _ptr(mean), _ptr(rstd), rows, D, C.c_float(eps), _stream())


def layernorm_bwd(dy, x, gamma, mean, rstd, *, add1=None, add2=None, dx=None, dx16=None, dgamma=None, dbeta=None):
    _chk(dy, F32, "dy"); _chk(x, F32, "x")
    rows, D = x.shape
    assert dy.shape == (rows, D) and dy.stride(1) == 1 and x.stride(1) == 1
    for t in (add1, add2, dx):
        assert t is None or (t.dtype == F32 and t.is_contiguous() and t.shape == (rows, D))
    assert dx16 is None or (dx16.dtype == BF16 and dx16.is_contiguous() and dx16.shape == (rows, D))
    call("egovlp_layernorm_bwd", _ptr(dy), C.c_longlong(dy.stride(0)), _ptr(x), C.c_longlong(x.stride(0)),
         _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(add1), _ptr(add2), _ptr(dx), _ptr(dx16), _ptr(dgamma),
         _ptr(dbeta), rows, D, _stream())


def cast_bf16(src, dst=None):
    _chk(src, F32, "src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    call("egovlp_cast_f32_to_bf16", _ptr(src), _ptr(dst), C.c_longlong(src.numel()), _stream())
    return dst


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
    call("egovlp_divided_attn_fwd", _ptr(qkv), _ptr(out), _ptr(lse), _ptr(ws), B, T, N, H, mode, _stream())
    return out, lse


def divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, mode, q_scale, dqkv=None):
    _chk(qkv, BF16, "qkv"); _chk(out, BF16, "out"); _chk(dout, BF16, "dout"); _chk(lse, F32, "lse")
    assert qkv.is_contiguous() and out.is_contiguous() and dout.is_contiguous() and lse.is_contiguous()
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    ws = torch.empty(B * H * 3 * 64, dtype=F32, device=qkv.device)
    call("egovlp_divided_attn_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(ws), B, T, N, H,
         mode, C.c_float(q_scale), _stream())
    return dqkv
