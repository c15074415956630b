"""Drop-in for nof/criteria/pointcloud_metrics.py:5-49.  The reference builds an Open3D KDTreeFlann (exact 1-NN search in
float64) per call and loops over the queries in Python; here the exact nearest neighbour comes from a brute-force float64
kernel (csrc/metrics.cu).  Inputs may be numpy arrays or tensors; results are python lists / floats like the reference's."""
import numpy as np
import torch

from ... import ops


def nn_correspondance(verts1, verts2):
    """for each vertex in verts2 find the nearest vertex in verts1 -> ([indices], [distances])  (:5-33)."""
    if len(verts1) == 0 or len(verts2) == 0:
        return [], []
    idx, dist = ops.nn_correspondance(verts1, verts2)
    return idx.cpu().numpy().tolist(), dist.cpu().numpy().tolist()


def eval_pts(pts1, pts2, threshold=0.2):
    """Chamfer distance and F-score (:39-49)."""
    _, d1 = ops.nn_correspondance(pts1, pts2)
    _, d2 = ops.nn_correspondance(pts2, pts1)
    if d1.shape[0] == 0 or d2.shape[0] == 0:
        return float("nan"), float("nan")                   # np.mean of an empty list in the reference
    s1, s2 = ops.dist_stats(d1, threshold), ops.dist_stats(d2, threshold)
    s = torch.stack([s1, s2]).cpu().numpy()                   # one D2H read of four numbers
    precision, recall = s[0, 1] / d1.shape[0], s[1, 1] / d2.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        fscore = np.float64(2 * precision * recall) / np.float64(precision + recall)
    cd = s[0, 0] / d1.shape[0] + s[1, 0] / d2.shape[0]
    return float(cd), float(fscore)
