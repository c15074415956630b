"""Drop-in for nof/criteria/loss.py:7-50: the range-loss modules behind `nof_loss` (train_kitti.py:46-48, :145-146).

Same class names and call signature `loss(pred, target, valid_mask=None)` as the reference; the arithmetic (optional
boolean selection, elementwise SmoothL1 / squared / absolute difference, mean over the selected elements, and the backward of
all of it) is one reduction kernel and one elementwise kernel of libpcnerf_b200.so (csrc/loss.cu, ops.masked_loss) instead
of a chain of eager torch kernels."""
from torch import nn

from ... import ops


class NOFLoss(nn.Module):
    """Base: `kind` names the elementwise term (ops.LOSS_KINDS); subclasses only choose it."""
    kind = None

    def forward(self, pred, target, valid_mask=None):
        if self.kind is None:
            raise TypeError("NOFLoss is abstract: use NOFSmoothL1Loss, NOFMSELoss or NOFL1Loss (nof_loss[...])")
        return ops.masked_loss(pred, target, valid_mask, self.kind)


class NOFMSELoss(NOFLoss):
    kind = "mse"


class NOFL1Loss(NOFLoss):
    kind = "l1"


class NOFSmoothL1Loss(NOFLoss):
    kind = "smoothl1"
