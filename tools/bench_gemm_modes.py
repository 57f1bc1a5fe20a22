"""Compare the CTA-pair and single-CTA tile schedulers on the step's shapes (M = 200768)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
M = int(os.environ.get("M", 200768))
def t(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
def rnd(*shape, dt=torch.bfloat16): return (torch.randn(*shape, device="cuda") * 0.05).to(dt)
rows = []
for name, N, K in [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]:
    a, w, bias = rnd(M, K), rnd(N, K), rnd(N, dt=torch.float32)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dy = rnd(M, N); dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(N, K, device="cuda", dtype=torch.float32)
    cases = {"fwd": lambda: ops.gemm(a, w, out, bias=bias), "dgrad": lambda: ops.gemm(dy, w, dx, b_mn=True),
             "wgrad8": lambda: ops.gemm(dy, a, dw, a_mn=True, b_mn=True, accumulate=True, split_k=8)}
    for cname, fn in cases.items():
        r = []
        for mode in ("0", "1"):
            os.environ["EGOVLP_GEMM_1CTA"] = mode
            r.append(t(fn))
        fl = 2.0 * M * N * K
        rows.append((f"{cname} {name}", r[0], r[1], fl))
print(f"M={M}   pair(ms)  single(ms)   pair TF/s  single TF/s")
for n, a_, b_, fl in rows:
    print(f"{n:14s} {a_:8.3f} {b_:8.3f}   {fl/a_/1e9:8.1f} {fl/b_/1e9:8.1f}")
