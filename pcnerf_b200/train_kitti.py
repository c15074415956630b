"""Hot-path part of train_kitti.py: NOFSystem.forward (:88-106), the loss assembly of training_step (:117-156) and
configure_optimizers (:108-115), without the Lightning / matplotlib / np.save plumbing (that stays the caller's).
`NOFSystem` here is a plain nn.Module with the same attribute names, so it can be dropped into the reference's
LightningModule by replacing the three methods.
"""
import torch
from torch.optim import SGD, Adam
from torch.optim.lr_scheduler import MultiStepLR

from .nof.criteria import nof_loss
from .nof.networks import Embedding, NOF_coarse, NOF_fine
from .nof.render import render_rays_train, render_rays_val


def get_optimizer(hparams, parameters):
    """nof/nof_utils.py:162-173."""
    eps = 1e-8
    if hparams.optimizer == 'sgd':
        return SGD(parameters, lr=hparams.lr, momentum=hparams.momentum, weight_decay=hparams.weight_decay)
    if hparams.optimizer == 'adam':
        return Adam(parameters, lr=hparams.lr, eps=eps, weight_decay=hparams.weight_decay)
    if hparams.optimizer == 'flat_adam':            # same update, one kernel over one flat buffer (pcnerf_b200.optim)
        from .optim import FlatAdam
        return FlatAdam(parameters, lr=hparams.lr, eps=eps, weight_decay=hparams.weight_decay)
    raise ValueError('optimizer not recognized!')


def decode_batch(batch):
    """nof/nof_utils.py:202-205."""
    return batch['rays'], batch['ranges']


class NOFSystem(torch.nn.Module):
    def __init__(self, hparams):
        super(NOFSystem, self).__init__()
        self.hparams = hparams
        self.embedding_position = Embedding(in_channels=3, N_freq=hparams.L_pos)
        self.nof_coarse = NOF_coarse(feature_size=hparams.feature_size, in_channels_xy=3 + 3 * hparams.L_pos * 2,
                                     use_skip=hparams.use_skip)
        self.nof_fine = NOF_fine(feature_size=hparams.feature_size, in_channels_xy=3 + 3 * hparams.L_pos * 2,
                                 use_skip=hparams.use_skip)
        self.loss = nof_loss[hparams.loss_type]()
        self.loss2 = nof_loss[hparams.loss_type]()

    def forward(self, rays, isval, **rng):
        """train_kitti.py:88-106."""
        hp = self.hparams
        if isval is False:
            return render_rays_train(
                model=self.nof_coarse, model_fine=self.nof_fine, embedding_xy=self.embedding_position, rays=rays,
                N_samples=hp.N_samples, N_importance=hp.N_importance, use_disp=hp.use_disp, perturb=hp.perturb,
                noise_std=hp.noise_std, chunk=hp.chunk, isval=isval, sub_nerf_test_num=hp.sub_nerf_test_num,
                issegmentated=hp.use_segmentated_sample, childnerf_ratio=hp.segmentated_child_nerf_ratio,
                use_child_nerf_divide=hp.use_child_nerf_divide, use_child_nerf_loss=hp.use_child_nerf_loss, **rng)
        return render_rays_val(
            model=self.nof_coarse, model_fine=self.nof_fine, embedding_xy=self.embedding_position, rays=rays,
            N_samples=hp.N_samples, N_importance=hp.N_importance, use_disp=hp.use_disp, perturb=hp.perturb,
            noise_std=hp.noise_std, chunk=hp.chunk, isval=isval, sub_nerf_test_num=hp.sub_nerf_test_num, **rng)

    def configure_optimizers(self):
        """train_kitti.py:108-115."""
        parameters = list(self.nof_coarse.parameters()) + list(self.nof_fine.parameters())
        self.optimizer = get_optimizer(self.hparams, parameters)
        self.scheduler = MultiStepLR(self.optimizer, milestones=[5, 120, 256], gamma=self.hparams.decay_gamma)
        return [self.optimizer], [self.scheduler]

    def training_step(self, batch, batch_idx=0, **rng):
        """Loss math of train_kitti.py:117-156 (logging, plotting and the np.save history dropped)."""
        hp = self.hparams
        rays, gt_ranges = decode_batch(batch)
        results = self.forward(rays, False, **rng)
        pred_ranges_fine = results['depth_fine']
        pred_ranges = results['depth']
        if hp.use_child_nerf_divide == 1:
            loss_range = torch.zeros(1, device=rays.device)
            loss_range_fine = torch.zeros(1, device=rays.device)
            sub_nerf = rays[:, 9]
            for i in range(hp.sub_nerf_test_num):
                sel = (sub_nerf > (i + 0.5)) & (sub_nerf < (i + 1.5))
                if sel.sum() >= 1:
                    loss_range = loss_range + 1e-1 * hp.lambda_loss * self.loss(1e1 * pred_ranges[sel], 1e1 * gt_ranges[sel])
                    loss_range_fine = loss_range_fine + 1e-1 * hp.lambda_loss_fine * self.loss(
                        1e1 * pred_ranges_fine[sel], 1e1 * gt_ranges[sel])
        else:
            loss_range = 1e-1 * hp.lambda_loss * self.loss(1e1 * pred_ranges, 1e1 * gt_ranges)
            loss_range_fine = 1e-1 * hp.lambda_loss * self.loss(1e1 * pred_ranges_fine, 1e1 * gt_ranges)   # sic: lambda_loss
        loss = loss_range + loss_range_fine + \
            hp.lambda_child_free_loss * results['child_free_loss_fine'] + hp.lambda_child_free_loss * results['child_free_loss'] + \
            hp.lambda_child_depth_loss * results['child_depth_loss_fine'] + hp.lambda_child_depth_loss * results['child_depth_loss']
        self.last_terms = {"loss_range": loss_range, "loss_range_fine": loss_range_fine, **results}
        return loss
