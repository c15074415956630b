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
        self.fuse_range_loss = True        # take the scene-level SmoothL1 range terms from K4 (False: the nof_loss modules)

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
        elif self.fuse_range_loss and hp.loss_type == 'smoothl1' and results.get('range_sl1') is not None:
            # the SmoothL1 means were accumulated by K4 in the compositing pass (forward and backward): no loss kernels
            # here.  Uses rays[:, 14] as ground truth, which is what batch['ranges'] holds (ipb2dmapping.py:447-453).
            loss_range = 1e-1 * hp.lambda_loss * results['range_sl1']
            loss_range_fine = 1e-1 * hp.lambda_loss * results['range_sl1_fine']                            # sic: lambda_loss
        else:
            loss_range = 1e-1 * hp.lambda_loss * self.loss(1e1 * pred_ranges, 1e1 * gt_ranges)
            loss_range_fine = 1e-1 * hp.lambda_loss * self.loss(1e1 * pred_ranges_fine, 1e1 * gt_ranges)   # sic: lambda_loss
        loss = loss_range + loss_range_fine + \
            hp.lambda_child_free_loss * results['child_free_loss_fine'] + hp.lambda_child_free_loss * results['child_free_loss'] + \
            hp.lambda_child_depth_loss * results['child_depth_loss_fine'] + hp.lambda_child_depth_loss * results['child_depth_loss']
        self.last_terms = {"loss_range": loss_range, "loss_range_fine": loss_range_fine, **results}
        return loss


def microbatch_scales(n_micro, n_local, world=1):
    """Loss factors that make gradient ACCUMULATION over ray micro-batches (and averaging over `world` data-parallel ranks)
    reproduce one batch of world x n_local rays (BASELINE.json configs[3]: 262,144 rays x 384 samples do not fit one pass:
    the saved activations alone are 8 x 512 B per sample).  The range and child-free losses are batch means
    (train_kitti.py:145-146, nof/render.py:121), so a micro-batch of n_micro rays contributes with n_micro / n_local; the
    child depth loss carries an extra 1/N (nof/render.py:155): (n_micro / n_local)^2 / world.  BatchNorm batches are the
    `chunk`-row chunks of the flattened samples either way, provided n_micro x N_samples and n_micro x (N_samples +
    N_importance) are multiples of `chunk`.  Returns (mean-term factor, depth-term factor)."""
    f = float(n_micro) / float(n_local)
    return f, f * f / float(world)


def accumulate_microbatches(system, rays, gt_ranges, micro_rays, world=1, **rng):
    """Forward + backward of one optimizer step's batch in micro-batches of `micro_rays` rays; gradients accumulate in
    .grad (zero them before).  Returns the batch loss (detached; the value one pass over the whole batch would log)."""
    hp = system.hparams
    n = rays.shape[0]
    total = None
    for a in range(0, n, micro_rays):
        r, g = rays[a:a + micro_rays], gt_ranges[a:a + micro_rays]
        f_mean, f_depth = microbatch_scales(r.shape[0], n, world)
        res = system.forward(r, False, **rng)
        if system.fuse_range_loss and hp.loss_type == 'smoothl1' and res.get('range_sl1') is not None:
            lr_c, lr_f = res['range_sl1'], res['range_sl1_fine']
        else:
            lr_c, lr_f = system.loss(1e1 * res['depth'], 1e1 * g), system.loss2(1e1 * res['depth_fine'], 1e1 * g)
        loss = f_mean * (1e-1 * hp.lambda_loss * lr_c + 1e-1 * hp.lambda_loss * lr_f
                         + hp.lambda_child_free_loss * (res['child_free_loss_fine'] + res['child_free_loss'])) \
            + f_depth * hp.lambda_child_depth_loss * (res['child_depth_loss_fine'] + res['child_depth_loss'])
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    return total


def load_ckpt(model, ckpt_path, model_name='model', prefixes_to_ignore=()):
    """nof/nof_utils.py:176-199: load the `model_name.*` entries of a (Lightning-style) checkpoint into `model`."""
    ckpt = torch.load(ckpt_path, map_location=torch.device('cpu'))
    if 'state_dict' in ckpt:
        ckpt = ckpt['state_dict']
    sd = model.state_dict()
    for k, v in ckpt.items():
        if k.startswith(model_name) and not any(k[len(model_name) + 1:].startswith(p) for p in prefixes_to_ignore):
            sd[k[len(model_name) + 1:]] = v
    model.load_state_dict(sd)


def fit(hparams, max_steps=None, device="cuda", precision=None, ckpt_path=None, history_paths=None):
    """A Lightning-free `trainer.fit(NOFSystem(hparams))` for the hot path (train_kitti.py:52-86, :262-296): build the dataset
    named by hparams.datasettype (`nof.dataset.nof_dataset`, .npy cache or files), shuffle it into batches of
    hparams.batch_size rays, and run training_step -> backward -> optimizer -> MultiStepLR (stepped per epoch) for
    hparams.num_epochs epochs (or max_steps).  The seven loss curves go to a device-side LossHistory (no per-step host
    synchronisation); a checkpoint with the reference's key layout (`state_dict` -> `nof_coarse.*`, `nof_fine.*`) is written
    to ckpt_path.  Returns (system, history array (steps, 7))."""
    from .nof.dataset import nof_dataset
    from .optim import LossHistory
    common = dict(root_dir=hparams.root_dir, data_start=hparams.data_start, data_end=hparams.data_end,
                  cloud_size_val=hparams.cloud_size_val, range_delete_x=hparams.range_delete_x,
                  range_delete_y=hparams.range_delete_y, range_delete_z=hparams.range_delete_z,
                  sub_nerf_test_num=hparams.sub_nerf_test_num, pose_path=hparams.pose_path, subnerf_path=hparams.subnerf_path,
                  surface_expand=hparams.surface_expand, re_loaddata=hparams.re_loaddata, result_path=hparams.result_path)
    if hparams.datasettype == "kitti_dataload":                                   # train_kitti.py:52-64
        common.update(parentnerf_path=hparams.parentnerf_path, interest_x=hparams.interest_x, interest_y=hparams.interest_y,
                      over_height=hparams.over_height, over_low=hparams.over_low)
    else:                                                                         # :65-77
        common.update(nerf_length_min=hparams.nerf_length_min, nerf_length_max=hparams.nerf_length_max,
                      nerf_width_min=hparams.nerf_width_min, nerf_width_max=hparams.nerf_width_max,
                      nerf_height_min=hparams.nerf_height_min, nerf_height_max=hparams.nerf_height_max)
    train = nof_dataset[hparams.datasettype](split='train', **common)
    system = NOFSystem(hparams).to(device)
    if precision is not None:
        system.nof_coarse.precision = system.nof_fine.precision = precision
    (opt,), (sched,) = system.configure_optimizers()
    hist = LossHistory(device=device)
    rays_all, ranges_all = train.all_rays.to(device), train.all_ranges.to(device)
    gen = torch.Generator(device="cpu").manual_seed(int(getattr(hparams, "seed", 0) or 0))
    step = 0
    system.train()
    for _ in range(hparams.num_epochs):
        perm = torch.randperm(rays_all.shape[0], generator=gen).to(device)
        for b in range(0, perm.shape[0], hparams.batch_size):
            idx = perm[b:b + hparams.batch_size]
            if idx.shape[0] < 2:                                                  # a one-row BN batch raises (like torch)
                continue
            opt.zero_grad()
            loss = system.training_step({'rays': rays_all[idx], 'ranges': ranges_all[idx]}, step)
            loss.backward()
            opt.step()
            t = system.last_terms
            hist.append([loss, t["loss_range"], t["loss_range_fine"],
                         hparams.lambda_child_free_loss * t["child_free_loss"], hparams.lambda_child_free_loss * t["child_free_loss_fine"],
                         hparams.lambda_child_depth_loss * t["child_depth_loss"], hparams.lambda_child_depth_loss * t["child_depth_loss_fine"]])
            step += 1
            if max_steps is not None and step >= max_steps:
                break
        sched.step()
        if max_steps is not None and step >= max_steps:
            break
    if ckpt_path:
        torch.save({'state_dict': {k: v.detach().cpu() for k, v in system.state_dict().items()}}, ckpt_path)
    if history_paths:
        hist.save(history_paths)
    return system, hist.to_numpy()
