"""Large scenes made of several parent NeRF blocks (BASELINE.json configs[4], SURVEY.md 8e "multi-parent scene").

The reference trains and renders ONE parent block per run (one `parentnerf_path`, one checkpoint, one set of child boxes:
train_kitti.py / eval_kitti_render.py); a large scene is a collection of such blocks (README.md:46).  This module is the
orchestration around the unchanged single-block path: each LiDAR return is routed to the parent block whose box contains
it, blocks are sharded over the ranks (every rank owns whole blocks: their networks, child boxes and rays), and a frame
is rendered block by block with K1 (candidate groups) + `eval_kitti_render.render_frame`.  No per-step communication:
a block's rays never leave its rank; gathering the rendered points is optional and happens once per frame.
"""
import numpy as np
import torch

from . import ops, parallel
from .eval_kitti_render import FramePlan, render_frame


class ParentBlock:
    """One parent NeRF: its box, its child boxes and its two occupancy networks (None until loaded on the owning rank)."""

    def __init__(self, parent_min, parent_max, child_bounds, child_bounds_larger=None, nof_coarse=None, nof_fine=None):
        self.parent_min = np.asarray(parent_min, dtype=np.float64).reshape(3)
        self.parent_max = np.asarray(parent_max, dtype=np.float64).reshape(3)
        self.child_bounds = np.asarray(child_bounds, dtype=np.float64).reshape(-1, 6)
        self.child_bounds_larger = self.child_bounds if child_bounds_larger is None else \
            np.asarray(child_bounds_larger, dtype=np.float64).reshape(-1, 6)
        self.nof_coarse, self.nof_fine = nof_coarse, nof_fine
        self._dev = {}

    def child_bounds_dev(self, device):
        """The child boxes as a float64 device tensor (uploaded once per block)."""
        k = ("b", str(device))
        if k not in self._dev:
            self._dev[k] = torch.as_tensor(self.child_bounds, dtype=torch.float64, device=device)
        return self._dev[k]

    def child_bounds_larger_dev(self, device):
        k = ("l", str(device))
        if k not in self._dev:
            self._dev[k] = torch.as_tensor(self.child_bounds_larger, dtype=torch.float64, device=device)
        return self._dev[k]


def route_points(points, parents):
    """Index of the first parent block whose closed box contains each point, -1 if none.  points (N,3); host numpy
    (one pass per frame, N x P comparisons; P is tens of blocks)."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.full(p.shape[0], -1, dtype=np.int64)
    for i in range(len(parents) - 1, -1, -1):                  # descending so that the FIRST containing block wins
        b = parents[i]
        inside = np.all((p >= b.parent_min) & (p <= b.parent_max), axis=1)
        out[inside] = i
    return out


def owned_blocks(n_parents, world_size=None, rank_=None, owners=None):
    """Parent indices of this rank.  owners (one rank per parent, e.g. `spread_owners`) or None: contiguous shards of whole
    blocks (parallel.shard_rows)."""
    w = parallel.world() if world_size is None else world_size
    r = parallel.rank() if rank_ is None else rank_
    if owners is not None:
        return [i for i in range(n_parents) if int(owners[i]) == r]
    a, b = parallel.shard_rows(n_parents, w, r)
    return list(range(a, b))


def spread_owners(parents, world_size):
    """Owner rank of every parent block such that spatially ADJACENT blocks land on different ranks: a LiDAR frame only sees
    the blocks around the sensor, so with contiguous shards one rank would render the whole frame.  Blocks are binned on
    the grid of their own typical size, block (ix, iy) goes to rank (ix + 3 iy) mod world (a 3 x 3 neighbourhood spreads over
    min(world, 8) ranks)."""
    c = np.stack([(b.parent_min + b.parent_max) / 2 for b in parents], 0)
    ext = np.median(np.stack([b.parent_max - b.parent_min for b in parents], 0), 0)
    ij = np.round((c[:, :2] - c[:, :2].min(0)) / np.maximum(ext[:2], 1e-9)).astype(np.int64)
    return ((ij[:, 0] + 3 * ij[:, 1]) % world_size).tolist()


@torch.no_grad()
def render_scene_frame(parents, origin, points, embedding_position, N_samples, N_importance, chunk,
                       depth_inference_method=2, batch_size_set=18432, grow_step=0.05, world_size=None, rank_=None):
    """Depth inference of one frame over the blocks THIS rank owns.  origin (3,), points (N,3): the frame's returns in
    the scene frame (they define the ray directions; eval_kitti_render.py:706-709).  Returns {parent index: (M,3) rendered
    points on the device} -- only owned blocks that received rays appear."""
    origin = np.asarray(origin, dtype=np.float64).reshape(3)
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    which = route_points(points, parents)
    out = {}
    for i in owned_blocks(len(parents), world_size, rank_):
        sel = np.nonzero(which == i)[0]
        if sel.size == 0:
            continue
        b = parents[i]
        if b.nof_coarse is None or b.nof_fine is None:
            raise RuntimeError("parent block %d is owned by this rank but its networks are not loaded" % i)
        vec = points[sel] - origin
        dist = np.linalg.norm(vec, axis=1)
        dirs = vec / dist[:, None]
        rays, _, other, _ = ops.aabb_build_groups(origin, dirs, dist, b.child_bounds, b.child_bounds_larger, b.parent_min,
                                                  b.parent_max, depth_inference_method, grow_step, 0.65)
        if rays.shape[0] == 0:
            continue
        out[i] = render_frame(b.nof_coarse, b.nof_fine, embedding_position, rays, other, N_samples, N_importance, chunk,
                              depth_inference_method=depth_inference_method, batch_size_set=batch_size_set)
    return out


def parent_boxes(parents):
    """(P,6) float64: min xyz, max xyz of every parent block (the layout ops.route_points takes)."""
    return np.stack([np.concatenate([b.parent_min, b.parent_max]) for b in parents], 0)


@torch.no_grad()
def render_scene_frames(parents, origins, points, frame_id, embedding_position, N_samples, N_importance, chunk,
                        depth_inference_method=2, batch_size_set=18432, grow_step=0.05, world_size=None, rank_=None,
                        boxes_dev=None, owners=None):
    """BATCHED depth inference of several frames over the parent blocks THIS rank owns (BASELINE.json configs[4]).

    origins (F,3): sensor position of every frame; points (M,3): the returns of all F frames in the scene frame;
    frame_id (M,): the frame of every return (device tensors, float64 / integer); owners: rank of every parent block
    (`spread_owners`; None = contiguous shards).  Every return is routed to the parent block
    that contains it (K0', one kernel); per owned block the rays of ALL frames of the batch are grouped (K1: candidate child
    boxes per ray) and rendered together (`render_frame`, once per physical ray) -- a block's networks are loaded once per
    batch instead of once per frame, and the MLP kernels see F times more rows per launch.  No communication: a block's rays
    never leave its rank.  Host synchronisations: one for the per-block ray counts, one per block for its group structure.
    Returns {parent index: (m,3) rendered points on the device}."""
    pts = ops._f64(points).reshape(-1, 3)
    dev = pts.device
    P = len(parents)
    if boxes_dev is None:
        boxes_dev = torch.as_tensor(parent_boxes(parents), dtype=torch.float64, device=dev)
    which, o, d, r = ops.route_points(pts, boxes_dev, origins, frame_id)
    order = torch.argsort(which, stable=True)                       # returns of block 0, block 1, ... (-1 first)
    counts = torch.bincount((which + 1).to(torch.int64), minlength=P + 1).tolist()
    starts = np.concatenate([[0], np.cumsum(counts)])
    out = {}
    for i in owned_blocks(P, world_size, rank_, owners):
        a, b_ = int(starts[i + 1]), int(starts[i + 2])
        if b_ <= a:
            continue
        blk = parents[i]
        if blk.nof_coarse is None or blk.nof_fine is None:
            raise RuntimeError("parent block %d is owned by this rank but its networks are not loaded" % i)
        idx = order[a:b_]
        rays, _, other, _ = ops.aabb_build_groups(o.index_select(0, idx), d.index_select(0, idx), r.index_select(0, idx),
                                                  blk.child_bounds_dev(dev), blk.child_bounds_larger_dev(dev),
                                                  blk.parent_min, blk.parent_max, depth_inference_method, grow_step, 0.65)
        if rays.shape[0] == 0:
            continue
        plan = FramePlan(rays, other, batch_size_set)
        out[i] = render_frame(blk.nof_coarse, blk.nof_fine, embedding_position, rays, other, N_samples, N_importance, chunk,
                              depth_inference_method=depth_inference_method, batch_size_set=batch_size_set, plan=plan)
    return out
