"""utils/mAP.py of the reference on the GPU (see nDCG.py in this package for the conventions).  Equal similarities
rank by smaller column first (a stable argsort of -sim, mAP.py:25)."""
import torch

from .. import ops
from .nDCG import _dev, _rel, _out


def calculate_AP(sim_mat, relevancy_matrix):
    """Average precision per query row (the vector mAP.py:42 averages)."""
    _, ap = ops.rank_metrics(_dev(sim_mat, torch.float32), _rel(relevancy_matrix), None, tie_mode=0, want_dcg=False)
    return _out(ap, sim_mat, relevancy_matrix)


def calculate_mAP(sim_mat, relevancy_matrix):
    """mAP.py:4-44."""
    _, ap = ops.rank_metrics(_dev(sim_mat, torch.float32), _rel(relevancy_matrix), None, tie_mode=0, want_dcg=False)
    return float(ap.mean())
