"""Drop-in for nof/criteria/metrics.py:5-32 (logging-only metrics of train_kitti.py:158-162,228-258)."""
import torch

from .pointcloud_metrics import eval_pts


def abs_error(pred, gt, valid_mask=None):
    value = torch.abs(pred - gt)
    if valid_mask is not None:
        value = value[valid_mask]
    return torch.mean(value)


def acc_thres(pred, gt, valid_mask=None):
    error = torch.abs(pred - gt)
    if valid_mask is not None:
        error = error[valid_mask]
    acc = error < 0.2
    return torch.sum(acc) / acc.shape[0] * 100


def eval_points(pred_pts, gt_pts, valid_mask=None):
    if valid_mask is not None:
        pred_pts = pred_pts[valid_mask]
        gt_pts = gt_pts[valid_mask]
    return eval_pts(pred_pts, gt_pts)
