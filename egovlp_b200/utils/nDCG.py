"""utils/nDCG.py of the reference on the GPU: same function names and argument meaning; the per-query ranking runs in
`egovlp_rank_metrics` (one CTA sorts one row in shared memory) instead of numpy argsort + fancy indexing.

Inputs may be numpy arrays (what model/metric.py:257-299 passes) or torch tensors; they are moved to the current
CUDA device, results come back as numpy (float64) when any input was numpy, else as CUDA tensors.  Equal
similarities rank by larger column first (a stable ascending argsort reversed, nDCG.py:32)."""
import numpy as np
import torch

from .. import ops


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    t = t.cuda() if not t.is_cuda else t
    return t.to(dtype) if dtype is not None and t.dtype != dtype else t


def _rel(x):
    t = _dev(x)
    return t if t.dtype in (torch.float32, torch.float64) else t.double()


def _out(t, *inputs):
    return t.cpu().numpy() if any(isinstance(i, np.ndarray) for i in inputs) else t


def calculate_k_counts(relevancy_matrix):
    """nDCG.py:47-75: [n1, n2] int mask, rank i counts iff i < number of relevant (> 0) items of the row."""
    rel = _rel(relevancy_matrix)
    k = (rel > 0).sum(dim=1, keepdim=True)
    kc = (torch.arange(rel.shape[1], device=rel.device)[None, :] < k).to(torch.int32)
    return _out(kc, relevancy_matrix)


def calculate_DCG(similarity_matrix, relevancy_matrix, k_counts):
    """nDCG.py:3-45: DCG per item of the first modality."""
    dcg, _ = ops.rank_metrics(_dev(similarity_matrix, torch.float32), _rel(relevancy_matrix),
                              None if k_counts is None else _dev(k_counts), tie_mode=1, want_ap=False)
    return _out(dcg, similarity_matrix, relevancy_matrix)


def calculate_IDCG(relevancy_matrix, k_counts):
    """nDCG.py:78-94."""
    return calculate_DCG(relevancy_matrix, relevancy_matrix, k_counts)


def calculate_nDCG(similarity_matrix, relevancy_matrix, k_counts=None, IDCG=None, reduction="mean"):
    """nDCG.py:96-139.  k_counts=None uses the kernel's built-in k (= calculate_k_counts) without materialising it."""
    rel = _rel(relevancy_matrix)
    kc = None if k_counts is None else _dev(k_counts)
    dcg, _ = ops.rank_metrics(_dev(similarity_matrix, torch.float32), rel, kc, tie_mode=1, want_ap=False)
    if IDCG is None:
        idcg, _ = ops.rank_metrics(rel.float(), rel, kc, tie_mode=1, want_ap=False)
    else:
        idcg = _dev(IDCG, torch.float64)
    if reduction == "mean":
        return float((dcg / idcg).mean())
    if reduction is None:
        return _out(dcg / idcg, similarity_matrix, relevancy_matrix)
