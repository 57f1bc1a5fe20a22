import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
B, T, N, H = 8, 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
x = torch.randn(M, D, device="cuda")
g, b_ = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
y16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
dy16 = torch.randn(M, D, device="cuda").bfloat16()
a1 = torch.randn(M, D, device="cuda")
dx, dx16 = torch.empty_like(x), torch.empty_like(y16)
dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
for _ in range(2):
    ops.layernorm_fwd(x, g, b_, 1e-6, y16=y16, mean=mean, rstd=rstd)
    ops.layernorm_bwd(dy16, x, g, mean, rstd, add1=a1, dx=dx, dx16=dx16, dgamma=dg, dbeta=db)
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16(); qkv[:, :D] *= 0.125
dout = torch.randn(M, D, device="cuda").bfloat16()
out, lse = ops.divided_attn_fwd(qkv, B, T, N, H, 0)
ops.divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, 0, 0.125)
torch.cuda.synchronize()
