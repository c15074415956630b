"""Drop-in for nof/criteria/loss.py:7-50 (masked-mean SmoothL1 / MSE / L1 wrappers).  These act on (N,) vectors of
rendered depths -- negligible work, kept as torch ops on the device so that they stay on the autograd tape."""
from torch import nn


class NOFLoss(nn.Module):
    def __init__(self):
        super(NOFLoss, self).__init__()
        self.loss = None

    def forward(self, pred, target, valid_mask=None):
        if valid_mask is not None:
            pred = pred[valid_mask]
            target = target[valid_mask]
        return self.loss(pred, target)


class NOFMSELoss(NOFLoss):
    def __init__(self):
        super(NOFMSELoss, self).__init__()
        self.loss = nn.MSELoss(reduction='mean')


class NOFL1Loss(NOFLoss):
    def __init__(self):
        super(NOFL1Loss, self).__init__()
        self.loss = nn.L1Loss(reduction='mean')


class NOFSmoothL1Loss(NOFLoss):
    def __init__(self):
        super(NOFSmoothL1Loss, self).__init__()
        self.loss = nn.SmoothL1Loss(reduction='mean')
