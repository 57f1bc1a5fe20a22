"""A/B of the Mlp GEMM pair at the step's shape (M = 200768): fc1 with act 1 (+pre-activation) vs act 3 (+derivative),
dgrad-fc2 with act 2 (recompute GELU') vs act 4 (multiply by the stored derivative), plain versions for scale."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops

M, D, HID = int(os.environ.get("M", 200768)), 768, 3072


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


rnd = lambda *sh, dt=torch.bfloat16: (torch.randn(*sh, device="cuda") * 0.05).to(dt)
x, w1, b1 = rnd(M, D), rnd(HID, D), rnd(HID, dt=torch.float32)
h, u = torch.empty(M, HID, device="cuda", dtype=torch.bfloat16), torch.empty(M, HID, device="cuda", dtype=torch.bfloat16)
dy, w2 = rnd(M, D), rnd(D, HID)
du, cs = torch.empty_like(h), torch.zeros(HID, device="cuda")
for name, fn in (("fc1 plain", lambda: ops.gemm(x, w1, h, bias=b1)),
                 ("fc1 act1 + pre-activation", lambda: ops.gemm(x, w1, h, bias=b1, act=1, out2=u)),
                 ("fc1 act3 + derivative", lambda: ops.gemm(x, w1, h, bias=b1, act=3, out2=u)),
                 ("dgrad-fc2 plain", lambda: ops.gemm(dy, w2, du, b_mn=True)),
                 ("dgrad-fc2 act2 (GELU' recomputed) + colsum", lambda: ops.gemm(dy, w2, du, b_mn=True, aux=u, act=2, colsum=cs)),
                 ("dgrad-fc2 act4 (stored GELU') + colsum", lambda: ops.gemm(dy, w2, du, b_mn=True, aux=u, act=4, colsum=cs)),
                 ("dgrad-fc2 act4, narrow epilogue (no colsum)", lambda: ops.gemm(dy, w2, du, b_mn=True, aux=u, act=4))):
    ms = t(fn)
    print(f"{name:46s} {ms:7.3f} ms  {2.0 * M * D * HID / ms / 1e9:7.0f} TFLOP/s")

# the memory-bound residual GEMMs (proj / fc2: fp32 residual stream in, fp32 out) and the QKV projection
xr, wp, bp = torch.randn(M, D, device="cuda"), rnd(D, D), rnd(D, dt=torch.float32)
yr = torch.empty_like(xr)
wq, q3 = rnd(3 * D, D), torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
for name, fn, fl, gb in (("proj + fp32 residual -> fp32", lambda: ops.gemm(x, wp, yr, bias=bp, residual=xr), 2.0 * M * D * D, M * D * 10e-9),
                         ("fc2 + fp32 residual -> fp32", lambda: ops.gemm(h, w2, yr, bias=bp, residual=xr), 2.0 * M * D * HID, M * (D * 8 + HID * 2) * 1e-9),
                         ("qkv (q columns scaled)", lambda: ops.gemm(x, wq, q3, bias=None, col_scale=0.125, col_scale_ncols=D), 2.0 * M * D * 3 * D, M * D * 8e-9)):
    ms = t(fn)
    print(f"{name:46s} {ms:7.3f} ms  {fl / ms / 1e9:7.0f} TFLOP/s  {gb / ms * 1e3:6.0f} GB/s")

# the fc1 weight gradient with / without the bias gradient summed from its dy tiles (colsum_a)
dw = torch.zeros(HID, D, device="cuda")
db = torch.zeros(HID, device="cuda")
for name, fn in (("wgrad fc1 (split-K 7)", lambda: ops.gemm(du, x, dw, a_mn=True, b_mn=True, accumulate=True, split_k=7)),
                 ("wgrad fc1 + colsum_a", lambda: ops.gemm(du, x, dw, a_mn=True, b_mn=True, accumulate=True, split_k=7, colsum_a=db))):
    ms = t(fn)
    print(f"{name:46s} {ms:7.3f} ms  {2.0 * M * D * HID / ms / 1e9:7.0f} TFLOP/s")
