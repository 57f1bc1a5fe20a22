import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
B, T, N, H = int(os.environ.get("B", 8)), 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16(); qkv[:, :D] *= 0.125
dout = torch.randn(M, D, device="cuda").bfloat16()
for mode in (1, 0):
    out, lse = ops.divided_attn_fwd(qkv, B, T, N, H, mode)
    ops.divided_attn_bwd(qkv, out, dout, lse, B, T, N, H, mode, 0.125)
torch.cuda.synchronize()
