"""Timeline of the tcgen05 space-attention backward (CTA 0, first groups): sets EGOVLP_ATTN_BWD_TRACE, launches once at
the step's shape and prints the stamped events in clock order (cycles relative to the first event).

    B=16 python tools/attn_bwd_trace.py [out.txt]
"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/attn_bwd_trace.txt"
raw = out_path + ".raw"
from egovlp_b200 import ops  # noqa: E402

B, T, N, H = int(os.environ.get("B", 16)), 16, 196, 12
S, D = 1 + T * N, 64 * H
M = B * S
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
qkv[:, :D] *= 0.125
dout = torch.randn(M, D, device="cuda").bfloat16()
o, lse = ops.divided_attn_fwd(qkv, B, T, N, H, 1)
dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.divided_attn_bwd(qkv, o, dout, lse, B, T, N, H, 1, 0.125, dqkv)
torch.cuda.synchronize()
os.environ["EGOVLP_ATTN_BWD_TRACE"] = raw
ops.divided_attn_bwd(qkv, o, dout, lse, B, T, N, H, 1, 0.125, dqkv)
torch.cuda.synchronize()
del os.environ["EGOVLP_ATTN_BWD_TRACE"]
NAMES = {1: "TMA   tile issued", 10: "MMA   S issue", 11: "MMA   dP issue", 12: "MMA   p_ready seen", 13: "MMA   dV issue",
         14: "MMA   ds_ready seen", 15: "MMA   dK/dQ issue", 16: "MMA   dV issued", 17: "MMA   dK/dQ issued", 20: "SOFT  s_full seen", 21: "SOFT  S in regs",
         22: "SOFT  P written", 23: "SOFT  dp_full seen", 24: "SOFT  dS written", 30: "DRAIN prepare start",
         31: "DRAIN prepare done", 32: "DRAIN dkv_full seen", 33: "DRAIN dK/dV stored", 34: "DRAIN dq_full seen",
         35: "DRAIN dQ stored"}
ev = sorted((int(c), int(e), int(g), int(i)) for e, g, i, c in (l.split() for l in open(raw)))
t0 = ev[0][0]
lines = [f"{c - t0:9d}  g{g} it{i}  {NAMES.get(e, e)}" for c, e, g, i in ev]
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:400]))
