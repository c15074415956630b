"""Drop-in for nof/criteria/metrics.py:5-32 (the logging metrics of train_kitti.py:158-162, :228-258): same names and
arguments; each is one masked reduction on the device (csrc/loss.cu) / the exact nearest-neighbour kernels (csrc/metrics.cu)."""
from ... import ops
from .pointcloud_metrics import eval_pts


def abs_error(pred, gt, valid_mask=None):
    """Mean absolute range error over the selected rays."""
    return ops.masked_loss(pred, gt, valid_mask, "abs_error")


def acc_thres(pred, gt, valid_mask=None):
    """Percentage of the selected rays whose absolute range error is below 0.2 m."""
    return ops.masked_loss(pred, gt, valid_mask, "acc_thres")


def eval_points(pred_pts, gt_pts, valid_mask=None):
    """Chamfer distance and F-score of two point sets (nof/criteria/pointcloud_metrics.py), optionally row-selected."""
    if valid_mask is None:
        return eval_pts(pred_pts, gt_pts)
    return eval_pts(pred_pts[valid_mask], gt_pts[valid_mask])
