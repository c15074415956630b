"""Data-parallel training of the two occupancy nets: rays are sharded across ranks (one process per GPU), every rank
runs the whole render path on its shard, and the 2 x 497,409 fp32 parameter gradients are summed with ONE NCCL
all-reduce of a flat 3.98 MB buffer over NVLink/NVSwitch.  The reference is single-GPU
(train_kitti.py:283-292); the semantics below are this package's definition:

  * BatchNorm statistics are per rank and per chunk (DDP semantics); every rank's running statistics evolve on its
    own shard, rank 0's are the ones a checkpoint would keep.
  * range / child-free losses are batch means, so the average of the per-rank gradients is the global-batch gradient;
    the child depth loss carries an extra 1/N (nof/render.py:155), so each rank's term is additionally divided by the
    world size (`depth_loss_scale`) to reproduce the loss of one batch of world x N rays.
  * depth inference shards physical rays (candidate groups are never split) and needs no communication.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def depth_loss_scale():
    """Factor for each rank's child depth loss so that the rank-averaged gradient equals the global-batch gradient."""
    return 1.0 / world()


class GradBucket:
    """One flat fp32 buffer that aliases the .grad of every parameter: the all-reduce needs no packing copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=p0.dtype, device=p0.device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            p._pcnerf_bucketed = True          # ops.MLPFunction.backward may accumulate straight into p.grad
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def realias(self):
        """Make every p.grad alias its slice of the flat buffer again.  Autograd accumulates in place when .grad exists,
        so normally nothing changed; `zero_grad(set_to_none=True)` / `p.grad = None` make it allocate fresh tensors, whose
        values are copied in (a parameter that received no gradient contributes zeros)."""
        o = 0
        for p in self.params:
            v = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
            o += p.numel()

    def allreduce_mean(self):
        """Sum over ranks then scale by 1/world (in place); returns the async work handle already waited on."""
        w = world()
        if w == 1:
            return
        self.realias()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / w)


def shard_rows(n, world_size, rank_):
    """Contiguous shard [start, stop) of n independent units (rays / candidate groups / parent blocks)."""
    base, rem = divmod(n, world_size)
    start = rank_ * base + min(rank_, rem)
    return start, start + base + (1 if rank_ < rem else 0)


def shard_groups(other, world_size, rank_):
    """Shard candidate rows of the depth-inference driver by PHYSICAL ray: `other` (N',) has head = n-1 >= 0 and
    followers = 0 with a head in front (eval_kitti_render.py:449-450).  Returns the row range [start, stop) of this
    rank such that no candidate group is split (the reference never splits one either, :989-999)."""
    tag = other.reshape(-1)
    n = tag.shape[0]
    if n == 0:
        return 0, 0
    # heads: rows that start a group.  A follower is a 0 that lies within `head value` rows after a head.
    heads = []
    i = 0
    tag_l = tag.tolist()
    while i < n:
        heads.append(i)
        i += int(tag_l[i]) + 1
    a, b = shard_rows(len(heads), world_size, rank_)
    start = heads[a] if a < len(heads) else n
    stop = heads[b] if b < len(heads) else n
    return start, stop
