"""wgrad of the qkv Linear at the step's shape (tokens 200768, dy [tokens, 2304], x [tokens, 768]) with and without
the fused bias gradient (column sums of dy taken from the A tiles in smem), against GEMM + separate colsum pass."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200 import ops
M = 200768


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


for n_out, n_in, split in ((2304, 768, 8), (768, 768, 16)):
    dy = (torch.randn(M, n_out, device="cuda") * 0.05).bfloat16()
    x = (torch.randn(M, n_in, device="cuda") * 0.05).bfloat16()
    dw, db = torch.zeros(n_out, n_in, device="cuda"), torch.zeros(n_out, device="cuda")
    plain = t(lambda: ops.gemm(dy, x, dw, a_mn=True, b_mn=True, accumulate=True, split_k=split))
    fused = t(lambda: ops.gemm(dy, x, dw, a_mn=True, b_mn=True, accumulate=True, split_k=split, colsum_a=db))
    sep = t(lambda: ops.colsum_accum(dy, db))
    print(f"wgrad [{n_out}x{n_in}] split {split}: plain {plain:.3f} ms, fused colsum {fused:.3f} ms, separate colsum {sep:.3f} ms"
          f"  -> {plain + sep - fused:+.3f} ms per Linear")
