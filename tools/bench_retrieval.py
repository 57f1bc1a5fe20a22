"""EPIC-Kitchens-100 MIR evaluation at its real size (9668 videos x 3842 unique sentences): GPU ranking metrics
(egovlp_rank_metrics) vs the numpy oracle (= the reference's host algorithm) on this box's CPU."""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egovlp_b200.utils import nDCG, mAP          # noqa: E402
from oracle import reference_port as rp          # noqa: E402

NV, NT = 9668, 3842
rng = np.random.default_rng(0)
sim = rng.random((NV, NT), dtype=np.float32)
rel = rng.choice([0.0, 0.0, 0.0, 0.0, 0.3, 1.0], size=(NV, NT))
rel[rng.integers(0, NV, NT), np.arange(NT)] = 1.0
rel[np.arange(NV), rng.integers(0, NT, NV)] = 1.0

s, r = torch.from_numpy(sim).cuda(), torch.from_numpy(rel).cuda()
st, rt = s.t().contiguous(), r.t().contiguous()


def gpu():
    return (nDCG.calculate_nDCG(s, r), nDCG.calculate_nDCG(st, rt), mAP.calculate_mAP(s, r), mAP.calculate_mAP(st, rt))


gpu()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g = gpu()
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 5
t0 = time.perf_counter()
c = (rp.ndcg(sim, rel), rp.ndcg(sim.T, rel.T), rp.average_precision(sim, rel).mean(), rp.average_precision(sim.T, rel.T).mean())
cpu_s = time.perf_counter() - t0
print("gpu", g)
print("cpu", tuple(float(x) for x in c))
print(f"GPU {gpu_ms:.2f} ms per full evaluation (nDCG + mAP, both directions; {6 * 2 * NV * NT * 6 / gpu_ms / 1e6:.0f} GB/s of row reads)"
      f"   CPU numpy {cpu_s:.2f} s   speed-up {cpu_s * 1e3 / gpu_ms:.0f}x   max |diff| {max(abs(a - float(b)) for a, b in zip(g, c)):.2e}")
