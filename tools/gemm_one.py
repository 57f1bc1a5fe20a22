"""One GEMM of the step's Mlp shapes, for `ncu --set full` captures:  python tools/gemm_one.py {fc1_act3|fc1_plain|dgrad_act4|proj_res}"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops

M, D, HID = int(os.environ.get("M", 200768)), 768, 3072
which = sys.argv[1] if len(sys.argv) > 1 else "fc1_act3"
rnd = lambda *sh, dt=torch.bfloat16: (torch.randn(*sh, device="cuda") * 0.05).to(dt)
x, w1, b1 = rnd(M, D), rnd(HID, D), rnd(HID, dt=torch.float32)
h, u = torch.empty(M, HID, device="cuda", dtype=torch.bfloat16), torch.empty(M, HID, device="cuda", dtype=torch.bfloat16)
dy, w2 = rnd(M, D), rnd(D, HID)
xr, wp, bp = torch.randn(M, D, device="cuda"), rnd(D, D), rnd(D, dt=torch.float32)
yr = torch.empty_like(xr)
fn = {"fc1_act3": lambda: ops.gemm(x, w1, h, bias=b1, act=3, out2=u),
      "fc1_plain": lambda: ops.gemm(x, w1, h, bias=b1),
      "dgrad_act4": lambda: ops.gemm(dy, w2, h, b_mn=True, aux=u, act=4),
      "proj_res": lambda: ops.gemm(x, wp, yr, bias=bp, residual=xr)}[which]
for _ in range(int(os.environ.get("ITERS", 3))):
    fn()
torch.cuda.synchronize()
