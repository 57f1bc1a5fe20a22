"""The three LayerNorm-backward shapes of a SpaceTimeBlock backward at B = 64 (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
B, T, N, H = int(os.environ.get("B", 64)), 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
x = torch.randn(M, D, device="cuda")
g = torch.ones(D, device="cuda")
mean, rstd = torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")
dy16 = torch.randn(M, D, device="cuda").bfloat16()
a32 = torch.randn(M, D, device="cuda")
a16, b16 = torch.randn(M, D, device="cuda").bfloat16(), torch.randn(M, D, device="cuda").bfloat16()
dx, dx16 = torch.empty_like(x), torch.empty_like(dy16)
dg, db, cs = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
variants = [
    ("LN2 (dy bf16, x, +fp32 add -> bf16)", 1.5 + 3 + 3 + 1.5, lambda: ops.layernorm_bwd(dy16, x, g, mean, rstd, add1=a32, dx16=dx16, dgamma=dg, dbeta=db, colsum_dx=cs)),
    ("LN1 (dy bf16, x -> bf16)", 1.5 + 3 + 1.5, lambda: ops.layernorm_bwd(dy16, x, g, mean, rstd, dx16=dx16, dgamma=dg, dbeta=db, colsum_dx=cs)),
    ("LN3 (dy bf16, x, +2 bf16 adds -> fp32 + bf16)", 1.5 + 3 + 3 + 3 + 1.5, lambda: ops.layernorm_bwd(dy16, x, g, mean, rstd, add1=a16, add2=b16, dx=dx, dx16=dx16, dgamma=dg, dbeta=db, colsum_dx=cs)),
]
for _ in range(2):
    for _, _, fn in variants:
        fn()
torch.cuda.synchronize()
if os.environ.get("TIME"):
    for name, kb_per_row, fn in variants:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:50s} {ms * 1e3:7.1f} us   {M * kb_per_row * 1024 / ms / 1e9:6.2f} TB/s")
