import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
M = 200768
def t(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
a = (torch.randn(M, 768, device="cuda") * 0.05).bfloat16(); w = (torch.randn(2304, 768, device="cuda") * 0.05).bfloat16()
bias = torch.randn(2304, device="cuda"); out = torch.empty(M, 2304, device="cuda", dtype=torch.bfloat16)
for mode in ("0", "1"):
    for nl in ("0", "1"):
        os.environ["EGOVLP_GEMM_1CTA"] = mode; os.environ["EGOVLP_GEMM_DEBUG_NOLOADS"] = nl
        ms = t(lambda: ops.gemm(a, w, out, bias=bias))
        print(f"qkv fwd  1cta={mode} noloads={nl}: {ms:.3f} ms  {2.0*M*2304*768/ms/1e9:.0f} TF/s")
