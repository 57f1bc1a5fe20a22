"""Micro-benchmark of the non-GEMM kernels at the step's shapes (default B=16, T=16, N=196, H=12)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops

B, T, N, H = int(os.environ.get("B", 16)), 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
HBM = 6569.6


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


def report(name, ms, nbytes):
    print(f"{name:34s} {ms:8.3f} ms   {nbytes / ms / 1e6:8.1f} GB/s  ({nbytes / ms / 1e6 / HBM * 100:5.1f}% of measured HBM)")


qkv = (torch.randn(M, 3 * D, device="cuda")).bfloat16()
qkv[:, :D] *= 0.125
dout = torch.randn(M, D, device="cuda").bfloat16()
for mode, name in ((1, "space"), (0, "time")):
    out, lse = ops.divided_attn_fwd(qkv, B, T, N, H, mode)
    report(f"attn fwd {name}", t(lambda: ops.divided_attn_fwd(qkv, B, T, N, H, mode)), M * (3 * D + D) * 2)
    dqkv = torch.empty_like(qkv)
    report(f"attn bwd {name}", t(lambda: ops.divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, mode, 0.125, dqkv)),
           M * (3 * D + D + D + 3 * D) * 2)

x = torch.randn(M, D, device="cuda")
g, b_ = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
y16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
report("layernorm fwd", t(lambda: ops.layernorm_fwd(x, g, b_, 1e-6, y16=y16, mean=mean, rstd=rstd)), M * D * 6)
dy, a1, a2 = torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda")
dx, dx16 = torch.empty_like(x), torch.empty_like(y16)
dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
report("layernorm bwd (+2 adds, +bf16)", t(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, add1=a1, add2=a2, dx=dx,
                                                                     dx16=dx16, dgamma=dg, dbeta=db)), M * D * (4 * 5 + 2))
report("layernorm bwd (plain)", t(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dx=dx, dx16=dx16, dgamma=dg, dbeta=db)),
       M * D * (4 * 3 + 2))
o = torch.zeros(D, device="cuda")
report("colsum fp32 [M,768]", t(lambda: ops.colsum_accum(dy, o)), M * D * 4)
du = torch.randn(M, 4 * D, device="cuda").bfloat16()
o4 = torch.zeros(4 * D, device="cuda")
report("colsum bf16 [M,3072]", t(lambda: ops.colsum_accum(du, o4)), M * 4 * D * 2)
o3 = torch.zeros(3 * D, device="cuda")
report("colsum bf16 [M,2304]", t(lambda: ops.colsum_accum(qkv, o3)), M * 3 * D * 2)
