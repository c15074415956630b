"""`nof_loss`: loss-type name (hparams.loss_type, nof/nof_utils.py) -> module class, as nof/criteria/__init__.py:4-8."""
from .loss import NOFL1Loss, NOFLoss, NOFMSELoss, NOFSmoothL1Loss  # noqa: F401

nof_loss = dict(mse=NOFMSELoss, l1=NOFL1Loss, smoothl1=NOFSmoothL1Loss)
