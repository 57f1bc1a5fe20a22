"""Divided space-time attention, forward and backward, at the step's geometry (B clips x 16 frames x 196 patches x 12 heads)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops

B, T, N, H = int(os.environ.get("B", 64)), 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
qkv[:, :D] *= 0.125
dout = torch.randn(M, D, device="cuda").bfloat16()


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for mode, name in ((0, "time"), (1, "space")):
    out, lse = ops.divided_attn_fwd(qkv, B, T, N, H, mode)
    fwd = t(lambda: ops.divided_attn_fwd(qkv, B, T, N, H, mode))
    bwd = t(lambda: ops.divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, mode, 0.125))
    gb_f, gb_b = M * (8 * D + 4 * H) / 1e9, M * (16 * D + 4 * H) / 1e9
    print(f"{name:5s} fwd {fwd:6.3f} ms ({gb_f / fwd * 1e3:5.0f} GB/s algorithmic)   bwd {bwd:6.3f} ms ({gb_b / bwd * 1e3:5.0f} GB/s)")
