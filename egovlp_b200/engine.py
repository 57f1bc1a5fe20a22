"""Host-side step engine: torch.autograd.Function wrappers that sequence the CUDA kernels of the hot path.

torch supplies device buffers, the current stream and the autograd graph (so DDP / optimizers of the unchanged
reference trainer keep working); every FLOP below runs in libegovlp_b200.so through `ops`.

Numerics layout: residual stream and LayerNorm statistics in fp32, GEMM operands in bf16 (fp32 accumulation in
TMEM), attention probabilities never leave the SM, losses in fp32.  fp32 master parameters are the autograd
leaves; their bf16 GEMM copies come from `Bf16Cache` (refreshed when a parameter's version changes).
"""
import functools
import os
import threading
import weakref

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32
Q_SCALE = 0.125            # head_dim ** -0.5 for head_dim = 64 (model/video_transformer.py:87)


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


class _ZeroArena:
    """One zero-filled fp32 buffer per block backward, carved into the ~20 accumulators (split-K weight gradients, bias /
    LayerNorm gradient vectors) that each used to cost a fill launch of its own.  Views keep the buffer alive."""

    def __init__(self, n_floats, like):
        self.buf = torch.zeros(n_floats, dtype=F32, device=like.device)
        self.off = 0

    def take(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        if self.off + n > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += (n + 63) // 64 * 64                     # 256-byte aligned carve-outs
        return v


_tls = threading.local()        # per autograd thread (= per device): the arena of the block backward in progress


def _zeros(shape, like):
    arena = getattr(_tls, "arena", None)
    if arena is not None and arena.buf.device == like.device:
        v = arena.take(shape)
        if v is not None:
            return v
    return torch.zeros(shape, dtype=F32, device=like.device)


# Generation counter of "some optimizer stepped": bumped by a global torch.optim post-step hook, so that bf16 copies
# are rebuilt even when the optimizer wrote through `p.data` (transformers.AdamW 4.x, apex), which does not bump
# `p._version`.  The fused egovlp_b200.optim.AdamW refreshes the copies itself (same kernel pass) and is exempt.
_GEN = 0


def _on_optimizer_step(optimizer, args, kwargs):
    global _GEN
    if not getattr(optimizer, "_egovlp_fused", False):
        _GEN += 1


try:                                                          # torch >= 2.0
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_on_optimizer_step)
except ImportError:                                           # pragma: no cover
    pass


class _CastEntry:
    __slots__ = ("ref", "t16", "version", "ptr", "gen", "trusted")

    def __init__(self, p, t16):
        self.ref, self.t16 = weakref.ref(p), t16
        self.version, self.ptr, self.gen, self.trusted = -1, 0, -1, False

    def stamp(self, p, trusted=False):
        self.version, self.ptr, self.gen, self.trusted = p._version, p.data_ptr(), _GEN, trusted

    def current(self, p):
        return (self.ref() is p and self.version == p._version and self.ptr == p.data_ptr() and self.gen == _GEN
                and self.t16.device == p.device)


_SHADOWS = {}          # id(param) -> _CastEntry: lets the fused AdamW write the bf16 copy in its own pass


def shadow_entry(p):
    ent = _SHADOWS.get(id(p))
    return ent if ent is not None and ent.ref() is p and ent.t16.device == p.device else None


class Bf16Cache:
    """bf16 GEMM-operand copies of the fp32 master parameters.

    * `get(p)` returns the copy, re-casting when the parameter object, its storage, its version or the global
      optimizer generation changed (entries hold a weak reference, so a recycled `id()` cannot alias a dead parameter).
    * `refresh()` -- called at the top of every TRAINING forward -- re-casts every known copy in ONE launch
      (egovlp_cast_multi_f32_to_bf16) unless the fused AdamW just wrote it, so an update made through `p.data` by any
      optimizer / EMA is always seen by the next training step."""

    def __init__(self):
        self._store = {}
        self._cats = {}
        self._table = None

    def _entry(self, p, t16_factory):
        ent = self._store.get(id(p))
        if ent is None or ent.ref() is not p or ent.t16.device != p.device:
            ent = _CastEntry(p, t16_factory())
            self._store[id(p)] = ent
            _SHADOWS[id(p)] = ent
            self._table = None
        if not ent.current(p):
            ops.cast_bf16(p.detach().contiguous(), ent.t16)
            ent.stamp(p)
        return ent

    def get(self, p, shape=None):
        t = self._entry(p, lambda: torch.empty(p.shape, dtype=BF16, device=p.device)).t16
        return t.view(shape) if shape is not None else t

    def cat(self, name, params):
        """bf16 concat along dim 0 of several parameters (DistilBERT q/k/v -> one [3D, D] operand): every part is a cache
        entry whose copy is a slice of one buffer, so `refresh()` keeps the concatenation current too."""
        buf = self._cats.get(name)
        rows = sum(p.shape[0] for p in params)
        if buf is None or buf.device != params[0].device or buf.shape[0] != rows:
            buf = torch.empty((rows,) + tuple(params[0].shape[1:]), dtype=BF16, device=params[0].device)
            self._cats[name] = buf
            for p in params:
                self._store.pop(id(p), None)
        r = 0
        for p in params:
            sl = buf[r:r + p.shape[0]]
            self._entry(p, lambda sl=sl: sl)
            r += p.shape[0]
        return buf

    def refresh(self):
        """Bring every cached copy up to date with one multi-tensor cast (no-op for copies the fused AdamW just wrote)."""
        live = []
        for key, ent in list(self._store.items()):
            p = ent.ref()
            if p is None or ent.t16.device != p.device:
                del self._store[key]
                if _SHADOWS.get(key) is ent:
                    del _SHADOWS[key]
                self._table = None
                continue
            if not (ent.trusted and ent.current(p)):
                live.append((p, ent))
        if not live:
            return
        key = tuple((p.data_ptr(), ent.t16.data_ptr(), p.numel()) for p, ent in live)
        if self._table is None or self._table[0] != key:
            self._table = (key,) + ops.build_cast_table([(p.detach(), ent.t16) for p, ent in live])
        ops.cast_multi(*self._table[1:])
        for p, ent in live:
            ent.stamp(p)

    def clear(self):
        for key, ent in self._store.items():
            if _SHADOWS.get(key) is ent:
                del _SHADOWS[key]
        self._store.clear()
        self._cats.clear()
        self._table = None


# By-products of an fp32 gradient tensor handed between consecutive Functions: its bf16 twin (saves one cast pass per
# block) and its column sums (the next block's fc2 bias gradient; saves one read of the tensor per block).  An entry
# holds the fp32 tensor itself, so its storage cannot be recycled for another tensor while the entry exists, and the
# tensor's version, so a gradient that autograd accumulated into in place (a block output with two consumers) is
# recognised and simply recomputed.  One slot per device; cleared when a video forward / backward starts.
_twin = {}
_twin_lock = threading.Lock()


def _publish_twin(t32, t16, colsum=None):
    with _twin_lock:
        _twin[t32.device.index] = (t32, t32._version, t16, colsum)


def _clear_twin(device):
    with _twin_lock:
        _twin.pop(device.index, None)


def _take_twin(t32):
    with _twin_lock:
        ent = _twin.pop(t32.device.index, None)
    if ent is None:
        return None, None
    src, version, t16, colsum = ent
    same = (src.data_ptr() == t32.data_ptr() and src.numel() == t32.numel() and src._version == version
            and t32.dtype == src.dtype and t32.is_contiguous())
    return (t16, colsum) if same else (None, None)


def _byproducts_of(t32):
    """-> (bf16 twin, column sums [D]) of an fp32 [M, D] gradient, from the producer if it published them."""
    t16, colsum = _take_twin(t32)
    if t16 is None:
        t16 = ops.cast_bf16(t32.contiguous())
    if colsum is None:
        colsum = bgrad(t32)
    return t16, colsum


def _bf16_of(t32):
    t16, _ = _take_twin(t32)
    if t16 is None:
        t16 = ops.cast_bf16(t32.contiguous())
    return t16


@functools.lru_cache(maxsize=None)
def _sm_count(device_index):
    return torch.cuda.get_device_properties(device_index).multi_processor_count


@functools.lru_cache(maxsize=None)
def _split_for(n_out, n_in, k_rows, n_sm=148):
    """Split-K factor of a weight-gradient GEMM (one persistent CTA per SM, 128 x 256 tiles, 64-row k-blocks): the split
    whose (tiles x splits) units fill whole waves of the grid.  Cost model = waves x (k-blocks per unit + 4 for the
    pipeline fill and the atomic epilogue that the next unit cannot hide); the smallest split within 3 % of the best
    (fewer fp32 atomics).  round(400 / tiles) left the 54-tile qkv gradient at 2.55 waves (378 units on 148 SMs)."""
    tiles = ((n_out + 127) // 128) * ((n_in + 255) // 256)
    num_kb = (k_rows + 63) // 64
    if os.environ.get("EGOVLP_WGRAD_SPLIT", "waves") == "legacy":       # A/B knob: the round-1 rule
        return max(1, min(num_kb, round(400 / tiles)))
    costs = {}
    for s in range(1, min(num_kb, 64) + 1):
        kbs = -(-num_kb // s)
        if -(-num_kb // kbs) != s:          # the kernel drops empty splits: same as a smaller s
            continue
        if s > 1 and kbs < 2:
            break
        costs[s] = -(-tiles * s // n_sm) * (kbs + 4)
    floor = min(costs.values())
    return min(s for s, c in costs.items() if c <= 1.03 * floor)


def wgrad(dy16, x16, n_out, n_in, bias_grad=None):
    """dW[n_out, n_in] = dy^T x (contraction over token rows), fp32, split-K atomics.  `bias_grad` (fp32 [n_out],
    zero-initialised) additionally receives colsum(dy), summed from the dy tiles while they are in shared memory."""
    dw = _zeros((n_out, n_in), dy16)
    ops.gemm(dy16, x16, dw, a_mn=True, b_mn=True, accumulate=True, split_k=_split_for(n_out, n_in, dy16.shape[0], _sm_count(dy16.device.index)),
             colsum_a=bias_grad)
    return dw


def wgrad_and_bgrad(dy16, x16, n_out, n_in):
    """(dW, db) of a Linear from its bf16 output gradient: one GEMM when the fused column sums apply (n_in % 256 == 0;
    +1.0 % on the step, A/B on one box: 280.5 / 282.7 vs 278.5 / 278.9 clips/s; EGOVLP_WGRAD_COLSUM=0 disables), else
    GEMM + a column-sum pass over dy."""
    if _WGRAD_COLSUM and n_in % 256 == 0 and n_out % 8 == 0:
        db = _zeros((n_out,), dy16)
        return wgrad(dy16, x16, n_out, n_in, bias_grad=db), db
    return wgrad(dy16, x16, n_out, n_in), bgrad(dy16)


_WGRAD_COLSUM = os.environ.get("EGOVLP_WGRAD_COLSUM", "1") != "0"
# Mlp pair: fc1 saves GELU'(pre-activation) (GEMM act 3) and the fc2 input-gradient GEMM multiplies by it (act 4);
# EGOVLP_GELU_DERIV=0 = the round-1 pair (pre-activation saved, GELU' recomputed in the dgrad epilogue), for A/B runs
_ACT_FWD, _ACT_BWD = (3, 4) if os.environ.get("EGOVLP_GELU_DERIV", "1") != "0" else (1, 2)


def bgrad(dy):
    db = _zeros((dy.shape[1],), dy)
    ops.colsum_accum(dy, db)
    return db


# ----------------------------------------------------------------------------------------------------------
# video tower
# ----------------------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """VideoPatchEmbed + cls/pos/temporal embedding assembly (model/video_transformer.py:72-77, 304-321)."""

    @staticmethod
    def forward(ctx, video, cls_token, pos_embed, temporal_embed, w, b, cache, norm=None):
        B, T, C, H, W = video.shape
        D, _, P, _ = w.shape
        _clear_twin(w.device)
        N = (H // P) * (W // P)
        S = 1 + T * N
        K = C * P * P
        patches = _empty((B * S, K), BF16, w)
        if video.dtype == torch.uint8:          # raw frames: dataset normalisation fused into the unfold
            mean, std = norm if norm is not None else ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
            ops.patch_im2col_u8(video.contiguous(), patches, P, mean, std)
        else:
            video = video.contiguous().float()
            ops.patch_im2col(video, patches, P)
        table = _empty((S, D), F32, w)
        ops.video_pos_table(cls_token.detach().contiguous(), pos_embed.detach().contiguous(),
                            temporal_embed.detach().contiguous(), b.detach(), table, T, N, D)
        x = _empty((B * S, D), F32, w)
        ops.gemm(patches, cache.get(w, (D, K)), x, bias=b.detach(), residual=table, res_row_mod=S)
        ctx.dims = (B, T, N, D, K, temporal_embed.shape[1])
        ctx.save_for_backward(patches)
        return x.view(B, S, D)

    @staticmethod
    def backward(ctx, dx):
        (patches,) = ctx.saved_tensors
        B, T, N, D, K, F = ctx.dims
        dx = dx.contiguous().view(-1, D)
        dx16 = _bf16_of(dx)
        dw = wgrad(dx16, patches, D, K)
        dcls, dpos = _zeros((1, 1, D), dx), _zeros((1, N + 1, D), dx)
        dtemp, dbias = _zeros((1, F, D), dx), _zeros((D,), dx)
        tmp = _empty(((1 + T * N) * D,), F32, dx)
        ops.video_embed_bwd(dx, tmp, dcls, dpos, dtemp, dbias, B, T, N, D)
        P = int(round((K // 3) ** 0.5))
        return None, dcls, dpos, dtemp, dw.view(D, 3, P, P), dbias, None, None


class SpaceTimeBlockFn(torch.autograd.Function):
    """SpaceTimeBlock.forward (model/video_transformer.py:163-177) incl. both VarAttention calls and the Mlp.

    params: norm1.{w,b}, attn.qkv.{w,b}, attn.proj.{w,b}, timeattn.qkv.{w,b}, timeattn.proj.{w,b},
            norm2.{w,b}, mlp.fc1.{w,b}, mlp.fc2.{w,b}, norm3.{w,b}      (18 tensors, reference order)
    """

    @staticmethod
    def forward(ctx, x, dims, eps, cache, *p):
        (n1w, n1b, sqw, sqb, spw, spb, tqw, tqb, tpw, tpb, n2w, n2b, f1w, f1b, f2w, f2b, n3w, n3b) = p
        B, T, N, H, grad_mode = dims
        D = H * 64
        S = 1 + T * N
        M = B * S
        HID = f1w.shape[0]
        x2 = x.contiguous().view(M, D)
        # `grad_mode` = torch.is_grad_enabled() at the call site: inside Function.forward grad mode is always off and
        # needs_input_grad reflects requires_grad of the parameters even under torch.no_grad()
        train = grad_mode and any(ctx.needs_input_grad)

        def ln(inp, w, b):
            y = _empty((M, D), BF16, inp)
            mean, rstd = _empty((M,), F32, inp), _empty((M,), F32, inp)
            ops.layernorm_fwd(inp, w.detach(), b.detach(), eps, y16=y, mean=mean, rstd=rstd)
            return y, mean, rstd

        def attention(inp16, qw, qb, pw, pb, mode, resid):
            qkv = _empty((M, 3 * D), BF16, inp16)
            ops.gemm(inp16, cache.get(qw), qkv, bias=qb.detach(), col_scale=Q_SCALE, col_scale_ncols=D)
            a, lse = ops.divided_attn_fwd(qkv, B, T, N, H, mode)
            out = _empty((M, D), F32, inp16)
            ops.gemm(a, cache.get(pw), out, bias=pb.detach(), residual=resid)
            return qkv, a, lse, out

        n3, mean3, rstd3 = ln(x2, n3w, n3b)
        qkv_t, a_t, lse_t, tr = attention(n3, tqw, tqb, tpw, tpb, 0, x2)          # time_residual = x + time_output
        n1, mean1, rstd1 = ln(tr, n1w, n1b)
        qkv_s, a_s, lse_s, sr = attention(n1, sqw, sqb, spw, spb, 1, x2)          # space_residual = x + space_output
        n2, mean2, rstd2 = ln(sr, n2w, n2b)
        h = _empty((M, HID), BF16, x2)
        u = _empty((M, HID), BF16, x2) if train else None          # GELU'(pre-activation), consumed by the backward
        ops.gemm(n2, cache.get(f1w), h, bias=f1b.detach(), act=_ACT_FWD if train else 1, out2=u)   # u = GELU'(fc1 output)
        y = _empty((M, D), F32, x2)
        ops.gemm(h, cache.get(f2w), y, bias=f2b.detach(), residual=sr)
        if train:
            ctx.dims, ctx.cache = (B, T, N, H, HID), cache
            ctx.save_for_backward(x2, n3, mean3, rstd3, qkv_t, a_t, lse_t, tr, n1, mean1, rstd1, qkv_s, a_s, lse_s, sr,
                                  n2, mean2, rstd2, u, h, *p)
        return y.view(B, S, D)

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        (x2, n3, mean3, rstd3, qkv_t, a_t, lse_t, tr, n1, mean1, rstd1, qkv_s, a_s, lse_s, sr, n2, mean2, rstd2, u,
         h) = sv[:20]
        (n1w, n1b, sqw, sqb, spw, spb, tqw, tqb, tpw, tpb, n2w, n2b, f1w, f1b, f2w, f2b, n3w, n3b) = sv[20:]
        B, T, N, H, HID = ctx.dims
        cache = ctx.cache
        D = H * 64
        M = x2.shape[0]
        dy = dy.contiguous().view(M, D)
        _tls.arena = _ZeroArena(2 * D * HID + 8 * D * D + HID + 24 * D + 32 * 64, dy)
        try:
            return SpaceTimeBlockFn._backward(ctx, dy, sv, B, T, N, H, HID, D, M, cache)
        finally:
            _tls.arena = None

    @staticmethod
    def _backward(ctx, dy, sv, B, T, N, H, HID, D, M, cache):
        (x2, n3, mean3, rstd3, qkv_t, a_t, lse_t, tr, n1, mean1, rstd1, qkv_s, a_s, lse_s, sr, n2, mean2, rstd2, u,
         h) = sv[:20]
        (n1w, n1b, sqw, sqb, spw, spb, tqw, tqb, tpw, tpb, n2w, n2b, f1w, f1b, f2w, f2b, n3w, n3b) = sv[20:]
        dy16, g_f2b = _byproducts_of(dy)                     # fc2 bias gradient = colsum(dy)

        # ---- MLP:  y = sr + fc2(gelu(fc1(LN2(sr))))
        g_f2w = wgrad(dy16, h, D, HID)
        du = _empty((M, HID), BF16, dy)
        ops.gemm(dy16, cache.get(f2w), du, b_mn=True, aux=u, act=_ACT_BWD)                # (dy W2) * gelu'
        g_f1w, g_f1b = wgrad_and_bgrad(du, n2, HID, D)       # fc1 bias gradient = colsum(du), summed inside the wgrad GEMM
        dn2 = _empty((M, D), BF16, dy)                       # LayerNorm-input gradients travel as bf16
        ops.gemm(du, cache.get(f1w), dn2, b_mn=True)
        del du
        # gradients that stay inside the block (d space_residual, d time_residual) are kept in bf16 only
        dsr16 = _empty((M, D), BF16, dy)
        g_n2w, g_n2b = _zeros((D,), dy), _zeros((D,), dy)
        g_spb = _zeros((D,), dy)                             # bias grad of attn.proj = colsum(d space_residual)
        ops.layernorm_bwd(dn2, sr, n2w.detach(), mean2, rstd2, add1=dy, dx16=dsr16, dgamma=g_n2w, dbeta=g_n2b,
                          colsum_dx=g_spb)
        del dn2

        def attention_bwd(dres16, qkv, a, lse, inp16, qw, pw, mode):
            dres = dres16
            g_pw = wgrad(dres16, a, D, D)
            da = _empty((M, D), BF16, dres)
            ops.gemm(dres16, cache.get(pw), da, b_mn=True)
            dqkv = ops.divided_attn_bwd(qkv, a, da, lse, B, T, N, H, mode, Q_SCALE)
            g_qw, g_qb = wgrad_and_bgrad(dqkv, inp16, 3 * D, D)
            dinp = _empty((M, D), BF16, dres)
            ops.gemm(dqkv, cache.get(qw), dinp, b_mn=True)
            return g_qw, g_qb, g_pw, dinp

        # ---- space attention:  sr = x + proj(attn(LN1(tr)))
        g_sqw, g_sqb, g_spw, dn1 = attention_bwd(dsr16, qkv_s, a_s, lse_s, n1, sqw, spw, 1)
        dtr16 = _empty((M, D), BF16, dy)
        g_n1w, g_n1b = _zeros((D,), dy), _zeros((D,), dy)
        g_tpb = _zeros((D,), dy)                             # bias grad of timeattn.proj = colsum(d time_residual)
        ops.layernorm_bwd(dn1, tr, n1w.detach(), mean1, rstd1, dx16=dtr16, dgamma=g_n1w, dbeta=g_n1b, colsum_dx=g_tpb)
        del dn1
        # ---- time attention:  tr = x + proj(timeattn(LN3(x)))
        g_tqw, g_tqb, g_tpw, dn3 = attention_bwd(dtr16, qkv_t, a_t, lse_t, n3, tqw, tpw, 0)
        dx, dx16 = _empty((M, D), F32, dy), _empty((M, D), BF16, dy)
        g_n3w, g_n3b = _zeros((D,), dy), _zeros((D,), dy)
        dx_colsum = _zeros((D,), dy)                         # = the fc2 bias gradient of the block below
        ops.layernorm_bwd(dn3, x2, n3w.detach(), mean3, rstd3, add1=dsr16, add2=dtr16, dx=dx, dx16=dx16, dgamma=g_n3w,
                          dbeta=g_n3b, colsum_dx=dx_colsum)
        _publish_twin(dx, dx16, dx_colsum)
        S = 1 + T * N
        return (dx.view(B, S, D), None, None, None, g_n1w, g_n1b, g_sqw, g_sqb, g_spw, g_spb, g_tqw, g_tqb, g_tpw,
                g_tpb, g_n2w, g_n2b, g_f1w, g_f1b, g_f2w, g_f2b, g_n3w, g_n3b)


class ClsHeadFn(torch.autograd.Function):
    """norm(x)[:, 0] -> optional Linear projection (model/video_transformer.py:330; model/model.py:77-79,141-142).
    LayerNorm is applied to the B CLS rows only (the reference normalises all S rows, then slices)."""

    @staticmethod
    def forward(ctx, x, eps, cache, nw, nb, pw, pb):
        B, S, D = x.shape
        x = x.contiguous()
        cls_rows = x.view(B, S * D)[:, :D]                      # row stride S*D
        y16 = _empty((B, D), BF16, x)
        y32 = _empty((B, D), F32, x)
        mean, rstd = _empty((B,), F32, x), _empty((B,), F32, x)
        ops.layernorm_fwd(cls_rows, nw.detach(), nb.detach(), eps, y16=y16, y32=y32, mean=mean, rstd=rstd)
        ctx.has_proj = pw is not None
        ctx.shape, ctx.cache = (B, S, D), cache
        if pw is None:
            ctx.save_for_backward(x, mean, rstd, nw)
            return y32
        out = _empty((B, pw.shape[0]), F32, x)
        ops.gemm(y16, cache.get(pw), out, bias=pb.detach())
        ctx.save_for_backward(x, mean, rstd, nw, y16, pw)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, S, D = ctx.shape
        dout = dout.contiguous().float()
        _clear_twin(dout.device)
        g_pw = g_pb = None
        if ctx.has_proj:
            x, mean, rstd, nw, y16, pw = ctx.saved_tensors
            Pd = pw.shape[0]
            d16 = ops.cast_bf16(dout)
            g_pw, g_pb = wgrad(d16, y16, Pd, D), bgrad(dout)
            dn = _empty((B, D), F32, dout)
            ops.gemm(d16, ctx.cache.get(pw), dn, b_mn=True)
        else:
            x, mean, rstd, nw = ctx.saved_tensors
            dn = dout
        dx = _zeros((B, S, D), dout)
        g_nw, g_nb = _zeros((D,), dout), _zeros((D,), dout)
        ops.layernorm_bwd(dn, x.view(B, S * D)[:, :D], nw.detach(), mean, rstd, dx=dx.view(B, S * D)[:, :D],
                          dgamma=g_nw, dbeta=g_nb)
        return dx, None, None, g_nw, g_nb, g_pw, g_pb


# ----------------------------------------------------------------------------------------------------------
# text tower (DistilBERT) + ReLU/Linear projection
# ----------------------------------------------------------------------------------------------------------
class TextTowerFn(torch.autograd.Function):
    """DistilBertModel(...).last_hidden_state -> (CLS | all tokens) -> ReLU -> Linear  (model/model.py:117-138).

    params: word_emb, pos_emb, emb_ln.{w,b}, then per layer
            q.{w,b}, k.{w,b}, v.{w,b}, out.{w,b}, sa_ln.{w,b}, lin1.{w,b}, lin2.{w,b}, out_ln.{w,b}   (16 / layer),
            finally txt_proj.{w,b}.
    `drop` = (p_hidden, p_attention): HuggingFace DistilBERT's train-mode dropouts -- on the embedding LayerNorm output,
    on the attention probabilities and on the FFN output (modeling_distilbert.py; the reference calls
    `self.text_model.train()`, model/model.py:36).  Masks come from a counter-based Philox stream keyed by one seed
    drawn per forward from torch's CPU generator (so `torch.manual_seed` makes a run reproducible); the backward
    regenerates them.  (0, 0) = eval mode / the deterministic parity path."""

    @staticmethod
    def forward(ctx, input_ids, attention_mask, heads, eps, tokens_mode, cache, drop, *p):
        """`tokens_mode`: False / True, or a (tokens_mode, grad_mode) pair -- see SpaceTimeBlockFn.forward."""
        grad_mode = True
        if isinstance(tokens_mode, tuple):
            tokens_mode, grad_mode = tokens_mode
        word, pos, elw, elb = p[:4]
        pw, pb = p[-2:]
        layers = [p[4 + 16 * i: 4 + 16 * (i + 1)] for i in range((len(p) - 6) // 16)]
        B, L = input_ids.shape
        D = word.shape[1]
        M = B * L
        ids = input_ids.contiguous().to(torch.int64)
        mask = attention_mask.contiguous().to(torch.int64)
        dev = word
        train = grad_mode and any(ctx.needs_input_grad)
        saved = []
        p_hid, p_att = (float(drop[0]), float(drop[1])) if drop else (0.0, 0.0)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (p_hid > 0 or p_att > 0) else 0

        emb = _empty((M, D), F32, dev)
        ops.text_embed_fwd(ids, word.detach(), pos.detach(), emb, B, L, D)
        x, x16 = _empty((M, D), F32, dev), _empty((M, D), BF16, dev)
        mean, rstd = _empty((M,), F32, dev), _empty((M,), F32, dev)
        ops.layernorm_fwd(emb, elw.detach(), elb.detach(), eps, y16=x16, y32=x, mean=mean, rstd=rstd)
        if p_hid > 0:
            ops.dropout(x, p_hid, seed, 0, y32=x, y16=x16)                                # embeddings dropout (in place)
        saved += [emb, mean, rstd]
        for li, lp in enumerate(layers):
            (qw, qb, kw, kb, vw, vb, ow, ob, sw, sb, l1w, l1b, l2w, l2b, fw, fb) = lp
            wqkv = cache.cat(("text_qkv_w", li, id(qw)), (qw, kw, vw))
            bqkv = torch.cat([qb.detach(), kb.detach(), vb.detach()])
            HID = l1w.shape[0]
            qkv = _empty((M, 3 * D), BF16, dev)
            ops.gemm(x16, wqkv, qkv, bias=bqkv, col_scale=Q_SCALE, col_scale_ncols=D)
            ctxv = _empty((M, D), BF16, dev)
            ops.text_attn_fwd(qkv, mask, ctxv, B, L, heads, p_att, seed, 1 + 2 * li)
            sa = _empty((M, D), F32, dev)
            ops.gemm(ctxv, cache.get(ow), sa, bias=ob.detach(), residual=x)               # sa_output + x
            x1, x1_16 = _empty((M, D), F32, dev), _empty((M, D), BF16, dev)
            m1, r1 = _empty((M,), F32, dev), _empty((M,), F32, dev)
            ops.layernorm_fwd(sa, sw.detach(), sb.detach(), eps, y16=x1_16, y32=x1, mean=m1, rstd=r1)
            hh, u = _empty((M, HID), BF16, dev), (_empty((M, HID), BF16, dev) if train else None)
            ops.gemm(x1_16, cache.get(l1w), hh, bias=l1b.detach(), act=_ACT_FWD if train else 1, out2=u)
            ff = _empty((M, D), F32, dev)
            if p_hid > 0:
                ops.gemm(hh, cache.get(l2w), ff, bias=l2b.detach())
                ops.dropout(ff, p_hid, seed, 2 + 2 * li, add=x1, y32=ff)                   # dropout(ffn_output) + sa_output
            else:
                ops.gemm(hh, cache.get(l2w), ff, bias=l2b.detach(), residual=x1)          # ffn_output + sa_output
            xn, xn16 = _empty((M, D), F32, dev), _empty((M, D), BF16, dev)
            m2, r2 = _empty((M,), F32, dev), _empty((M,), F32, dev)
            ops.layernorm_fwd(ff, fw.detach(), fb.detach(), eps, y16=xn16, y32=xn, mean=m2, rstd=r2)
            saved += [x16, qkv, ctxv, sa, m1, r1, x1_16, u, hh, ff, m2, r2]
            x, x16 = xn, xn16
        rows, stride = (M, D) if tokens_mode else (B, L * D)
        if pw is None:                 # projection='' (nn.Identity, model/model.py:80-82): the hidden state itself
            r16 = None
            out = x if tokens_mode else x.view(B, L * D)[:, :D].clone()
        else:
            r16 = _empty((rows, D), BF16, dev)
            ops.relu_rows_fwd(x, stride, r16, rows, D)
            out = _empty((rows, pw.shape[0]), F32, dev)
            ops.gemm(r16, cache.get(pw), out, bias=pb.detach())
        if train:
            ctx.meta = (B, L, D, heads, tokens_mode, len(layers), len(saved), p_hid, p_att, seed)
            ctx.cache = cache
            ctx.has_proj = pw is not None
            ctx.save_for_backward(ids, mask, x, r16, *saved, *p[:len(p) - (0 if pw is not None else 2)])
        return out.view(B, L, -1) if tokens_mode else out

    @staticmethod
    def backward(ctx, dout):
        B, L, D, heads, tokens_mode, n_layers, n_saved, p_hid, p_att, seed = ctx.meta
        cache = ctx.cache
        sv = ctx.saved_tensors
        ids, mask, x_last, r16 = sv[:4]
        saved = list(sv[4:4 + n_saved])
        p = sv[4 + n_saved:]
        word, pos, elw, elb = p[:4]
        layers = [p[4 + 16 * i: 4 + 16 * (i + 1)] for i in range(n_layers)]
        M = B * L
        rows, stride = (M, D) if tokens_mode else (B, L * D)
        if ctx.has_proj:
            pw, pb = p[-2:]
            Pd = pw.shape[0]
            dout = dout.contiguous().float().view(rows, Pd)
            d16 = ops.cast_bf16(dout)
            g_pw, g_pb = wgrad(d16, r16, Pd, D), bgrad(dout)
            dr = _empty((rows, D), F32, dout)
            ops.gemm(d16, cache.get(pw), dr, b_mn=True)
            dx = _zeros((M, D), dout)
            ops.relu_rows_bwd(x_last, stride, dr, dx, rows, D)
        else:
            g_pw = g_pb = None
            dout = dout.contiguous().float().view(rows, D)
            if tokens_mode:
                dx = dout
            else:
                dx = _zeros((M, D), dout)
                dx.view(B, L * D)[:, :D].copy_(dout)
        grads = []
        for li in reversed(range(n_layers)):
            (qw, qb, kw, kb, vw, vb, ow, ob, sw, sb, l1w, l1b, l2w, l2b, fw, fb) = layers[li]
            x16, qkv, ctxv, sa, m1, r1, x1_16, u, hh, ff, m2, r2 = saved[3 + 12 * li: 3 + 12 * (li + 1)]
            HID = l1w.shape[0]
            dff, dff16 = _empty((M, D), F32, dout), _empty((M, D), BF16, dout)
            g_fw, g_fb = _zeros((D,), dout), _zeros((D,), dout)
            ops.layernorm_bwd(dx, ff, fw.detach(), m2, r2, dx=dff, dx16=dff16, dgamma=g_fw, dbeta=g_fb)
            dffn, dffn16 = dff, dff16                      # gradient of the FFN output (before its dropout)
            if p_hid > 0:
                dffn, dffn16 = _empty((M, D), F32, dout), _empty((M, D), BF16, dout)
                ops.dropout(dff, p_hid, seed, 2 + 2 * li, y32=dffn, y16=dffn16)
            g_l2w, g_l2b = wgrad(dffn16, hh, D, HID), bgrad(dffn)
            du = _empty((M, HID), BF16, dout)
            ops.gemm(dffn16, cache.get(l2w), du, b_mn=True, aux=u, act=_ACT_BWD)
            g_l1w, g_l1b = wgrad(du, x1_16, HID, D), bgrad(du)
            dx1 = _empty((M, D), F32, dout)
            ops.gemm(du, cache.get(l1w), dx1, b_mn=True, residual=dff)                     # + residual path
            dsa, dsa16 = _empty((M, D), F32, dout), _empty((M, D), BF16, dout)
            g_sw, g_sb = _zeros((D,), dout), _zeros((D,), dout)
            ops.layernorm_bwd(dx1, sa, sw.detach(), m1, r1, dx=dsa, dx16=dsa16, dgamma=g_sw, dbeta=g_sb)
            g_ow, g_ob = wgrad(dsa16, ctxv, D, D), bgrad(dsa)
            dctx = _empty((M, D), BF16, dout)
            ops.gemm(dsa16, cache.get(ow), dctx, b_mn=True)
            dqkv = _empty((M, 3 * D), BF16, dout)
            ops.text_attn_bwd(qkv, mask, dctx, dqkv, B, L, heads, Q_SCALE, p_att, seed, 1 + 2 * li)
            g_wqkv, g_bqkv = wgrad(dqkv, x16, 3 * D, D), bgrad(dqkv)
            wqkv = cache.cat(("text_qkv_w", li, id(qw)), (qw, kw, vw))
            dxin = _empty((M, D), F32, dout)
            ops.gemm(dqkv, wqkv, dxin, b_mn=True, residual=dsa)                            # x feeds qkv and the residual
            dx = dxin
            gq, gk, gv = g_wqkv[:D], g_wqkv[D:2 * D], g_wqkv[2 * D:]
            bq, bk, bv = g_bqkv[:D], g_bqkv[D:2 * D], g_bqkv[2 * D:]
            grads = [gq, bq, gk, bk, gv, bv, g_ow, g_ob, g_sw, g_sb, g_l1w, g_l1b, g_l2w, g_l2b, g_fw, g_fb] + grads
        emb, mean, rstd = saved[:3]
        if p_hid > 0:
            ops.dropout(dx, p_hid, seed, 0, y32=dx)                                       # embeddings dropout, backward
        demb = _empty((M, D), F32, dout)
        g_elw, g_elb = _zeros((D,), dout), _zeros((D,), dout)
        ops.layernorm_bwd(dx, emb, elw.detach(), mean, rstd, dx=demb, dgamma=g_elw, dbeta=g_elb)
        g_word, g_pos = torch.zeros_like(word), torch.zeros_like(pos)
        ops.text_embed_bwd(ids, demb, g_word, g_pos, B, L, D)
        return (None, None, None, None, None, None, None, g_word, g_pos, g_elw, g_elb, *grads, g_pw, g_pb)


# ----------------------------------------------------------------------------------------------------------
# similarity + losses
# ----------------------------------------------------------------------------------------------------------
class SimMatrixFn(torch.autograd.Function):
    """sim_matrix (model/model.py:189-197): cosine similarity with the norm clamped at eps, fp32."""

    @staticmethod
    def forward(ctx, a, b, eps):
        a, b = a.contiguous().float(), b.contiguous().float()
        an, na = ops.rownorm_fwd(a, eps)
        bn, nb = ops.rownorm_fwd(b, eps)
        ctx.eps = eps
        ctx.save_for_backward(an, na, bn, nb)
        return ops.sgemm(an, bn)

    @staticmethod
    def backward(ctx, dx):
        an, na, bn, nb = ctx.saved_tensors
        dx = dx.contiguous().float()
        dan = ops.sgemm(dx, bn, trans_b=False)                    # dX @ bn
        dbn = ops.sgemm(dx, an, trans_a=True, trans_b=False)      # dX^T @ an
        return ops.rownorm_bwd(dan, an, na, ctx.eps), ops.rownorm_bwd(dbn, bn, nb, ctx.eps), None


class NceLossFn(torch.autograd.Function):
    """EgoNCE / InfoNCE on a similarity matrix with a uint8 positives mask (model/loss.py:13-25, 34-53)."""

    @staticmethod
    def forward(ctx, x, mask, temperature):
        x = x.contiguous().float()
        loss, stats = ops.nce_fwd(x, mask, 1.0 / temperature)
        ctx.inv_temp = 1.0 / temperature
        ctx.save_for_backward(x, mask, stats)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, mask, stats = ctx.saved_tensors
        return ops.nce_bwd(x, mask, stats, ctx.inv_temp, g.contiguous().float()), None, None


class FusedEgoNceFn(torch.autograd.Function):
    """sim_matrix + EgoNCE / InfoNCE from the gathered embeddings and multi-hot tags in ONE kernel per direction
    (csrc/loss_fused.cu; reference trainer/trainer_egoclip.py:130-135).  `text` / `video` [G, C], `verb` / `noun`
    [G, n] are fp32 row-strided views (e.g. column slices of the packed all-gather buffer, read in place).  Gradients
    are produced for rows [row0, row0 + n_local) only and returned zero-padded to [G, C] when n_local < G is not the
    whole matrix -- `GatherEgoNceFn` below hands the local slice out directly instead."""

    @staticmethod
    def forward(ctx, text, video, verb, noun, temperature, mode):
        loss, saved = ops.egonce_fused_fwd(text, video, verb, noun, 1.0 / temperature, mode)
        ctx.args = (1.0 / temperature, mode, verb.shape[1] if verb is not None else 0,
                    noun.shape[1] if noun is not None else 0)
        ctx.save_for_backward(text, video, *saved)
        return loss

    @staticmethod
    def backward(ctx, g):
        text, video, *saved = ctx.saved_tensors
        inv_temp, mode, nv, nn_ = ctx.args
        dt, dv = ops.egonce_fused_bwd(text, video, saved, nv, nn_, inv_temp, mode, g.contiguous().float(), 0,
                                      text.shape[0])
        return dt, dv, None, None, None, None


class GatherEgoNceFn(torch.autograd.Function):
    """The exchange step of the data-parallel training step, fused with the loss: local embeddings + tags -> ONE packed
    all-gather (one pack launch, one ncclAllGather) -> the fused EgoNCE kernel on column views of the gathered buffer ->
    loss; the backward kernel emits d text / d video of THIS rank's rows only (the reference's AllGather_multi backward,
    trainer/trainer_egoclip.py:23-27).  `gather(packed_local) -> packed_all` is the collective (identity at world 1)."""

    @staticmethod
    def forward(ctx, text, video, verb, noun, temperature, mode, gather, rank):
        B, Ct = text.shape
        Cv, nv, nn_ = video.shape[1], verb.shape[1], noun.shape[1]
        assert Ct == Cv
        packed = ops.pack_rows4(text.contiguous().float(), video.contiguous().float(), verb.contiguous().float(),
                                noun.contiguous().float())
        allp = gather(packed)                              # [world * B, Ct + Cv + nv + nn], rank-major
        t, v = allp[:, :Ct], allp[:, Ct:Ct + Cv]
        vb, nb_ = allp[:, Ct + Cv:Ct + Cv + nv], allp[:, Ct + Cv + nv:]
        loss, saved = ops.egonce_fused_fwd(t, v, vb, nb_, 1.0 / temperature, mode)
        ctx.args = (1.0 / temperature, mode, nv, nn_, Ct, Cv, rank * B, B)
        ctx.save_for_backward(allp, *saved)
        return loss

    @staticmethod
    def backward(ctx, g):
        allp, *saved = ctx.saved_tensors
        inv_temp, mode, nv, nn_, Ct, Cv, row0, B = ctx.args
        dt, dv = ops.egonce_fused_bwd(allp[:, :Ct], allp[:, Ct:Ct + Cv], saved, nv, nn_, inv_temp, mode,
                                      g.contiguous().float(), row0, B)
        return dt, dv, None, None, None, None, None, None


class MaxMarginFn(torch.autograd.Function):
    """MaxMarginRankingLoss (model/loss.py:63-90) and, with `row_weight`, AdaptiveMaxMarginRankingLoss
    (model/loss.py:100-133); no host-side index building.  The weight gets no gradient (the reference feeds the
    dataset's relevancy, a constant)."""

    @staticmethod
    def forward(ctx, x, margin, fix_norm, row_weight=None):
        x = x.contiguous().float()
        w = None if row_weight is None else row_weight.detach().contiguous().float()
        ctx.args = (margin, fix_norm)
        ctx.has_w = w is not None
        ctx.save_for_backward(*((x, w) if w is not None else (x,)))
        return ops.maxmargin_fwd(x, margin, fix_norm, w)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        w = ctx.saved_tensors[1] if ctx.has_w else None
        return ops.maxmargin_bwd(x, ctx.args[0], ctx.args[1], g.contiguous().float(), w), None, None, None
