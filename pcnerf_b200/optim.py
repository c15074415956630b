"""Flat-buffer optimizer and device-side loss history for the training loop.

`FlatAdam` is torch.optim.Adam as the reference configures it (nof/nof_utils.py:162-173: lr, eps = 1e-8, weight_decay)
with every parameter of the two occupancy nets re-homed into ONE contiguous fp32 buffer (the tensors keep their identity,
shapes and names: only their storage moves) next to the flat gradient buffer of `parallel.GradBucket`; a step is one
element-wise kernel over 994,818 values with the 1/world gradient averaging folded in.  It subclasses
torch.optim.Optimizer with a single parameter group, so `MultiStepLR` (train_kitti.py:113) drives its learning rate
unchanged; the rate is mirrored into a device scalar, which keeps the step replayable from a CUDA graph.

`LossHistory` replaces the every-5-steps `.cpu()` + 7 x np.save of train_kitti.py:164-189 by appends into a device ring
(no host synchronisation in the step); `save()` writes the same seven .npy files on demand.
"""
import numpy as np
import torch

from . import parallel
from ._lib import check, lib
from .ops import _p, _stream, note_param_write


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FlatAdam: no parameters")
        dev = params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise RuntimeError("FlatAdam: parameters must be float32 CUDA tensors on one device (no CPU fallback)")
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        o = 0
        with torch.no_grad():
            for p in params:                                   # move the storage, keep the tensor objects
                self.flat[o:o + p.numel()].copy_(p.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        self.bucket = parallel.GradBucket(params)              # .grad of every parameter aliases bucket.flat
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._coef = torch.zeros(2, dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def zero_grad(self, set_to_none=False):
        self.bucket.zero()

    @torch.no_grad()
    def step(self, closure=None, allreduce=True):
        """All-reduce (sum) of the flat gradients when running data-parallel, then ONE Adam kernel (1/world folded in)."""
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        if self._lr_host is None or float(g["lr"]) != self._lr_host:                    # a scheduler moved the rate: mirror it (outside any graph)
            self._lr_host = float(g["lr"])
            self.lr_dev.fill_(self._lr_host)
        w = parallel.world()
        if not torch.cuda.is_current_stream_capturing():
            self.bucket.realias()                              # nn.Module.zero_grad(set_to_none=True) detaches .grad
        if allreduce and w > 1:
            torch.distributed.all_reduce(self.bucket.flat, op=torch.distributed.ReduceOp.SUM)
        b1, b2 = g["betas"]
        check(lib().pcnerf_adam_step(_p(self.flat), _p(self.bucket.flat), _p(self.exp_avg), _p(self.exp_avg_sq),
                                     self.flat.numel(), _p(self.step_dev), _p(self.lr_dev), _p(self._coef), float(b1),
                                     float(b2), float(g["eps"]), float(g["weight_decay"]), 1.0 / w if allreduce else 1.0,
                                     _stream()))
        note_param_write()                                     # the kernel wrote every parameter through a raw pointer
        return loss

    # ---- checkpoint / resume: the moments and the step count live outside Optimizer.state (flat buffers)
    def state_dict(self):
        sd = super().state_dict()
        sd["flat_adam"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                           "step": self.step_dev.detach().clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("flat_adam", None)
        super().load_state_dict(state_dict)
        if extra is not None:
            with torch.no_grad():
                self.exp_avg.copy_(extra["exp_avg"])
                self.exp_avg_sq.copy_(extra["exp_avg_sq"])
                self.step_dev.copy_(extra["step"])
        self._lr_host = None                                   # re-mirror the (possibly restored) learning rate


class LossHistory:
    """Device-side history of the seven curves the reference saves (train_kitti.py:168-189)."""
    NAMES = ("loss", "loss_range", "loss_range_fine", "loss_child_free", "loss_child_free_fine", "loss_child_depth",
             "loss_child_depth_fine")

    def __init__(self, capacity=65536, device="cuda"):
        self.buf = torch.zeros((capacity, len(self.NAMES)), dtype=torch.float32, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)

    def append(self, terms):
        """terms: the seven 0-d / 1-element tensors in NAMES order.  No host synchronisation; wraps around when full."""
        row = torch.stack([t.detach().reshape(()).to(torch.float32) for t in terms]).reshape(1, -1)
        self.buf.index_copy_(0, self.count % self.buf.shape[0], row)
        self.count += 1

    def to_numpy(self):
        n = min(int(self.count.item()), self.buf.shape[0])
        return self.buf[:n].cpu().numpy()

    def save(self, paths):
        """paths: seven file names in NAMES order (saveploty_path, saveploty_path_range, ...)."""
        h = self.to_numpy()
        for i, p in enumerate(paths):
            np.save(p, arr=h[:, i])
