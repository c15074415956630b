"""Drop-in for the AABB leaf functions of nof/dataset/ipb2dmapping.py (same names / argument meaning / error
behaviour) plus the batched ray-packing rule that the reference inlines in its dataset loaders.

The scalar-style functions accept what the reference accepts (numpy arrays / python scalars for ONE ray) and return
python / numpy values; they launch the same fp64 kernels as the batched entry point, so they are meant for parity
checks and drop-in compatibility -- use `pack_train_rays` for real work.  File IO (PCD / pose parsing, filtering,
the .npy cache) is out of scope.
"""
import numpy as np

from ... import ops


def compute_far_bound(ray_o, ray_d, x_max, x_min, y_max, y_min, z_max, z_min):
    """ipb2dmapping.py:36-77.  Returns the distance, or None when no plane is hit in front of the ray."""
    t = float(ops.aabb_far_bound(np.asarray(ray_o, dtype=np.float64), np.asarray(ray_d, dtype=np.float64).reshape(1, 3),
                                 x_max, x_min, y_max, y_min, z_max, z_min)[0].item())
    return None if t == np.inf else t


def _one(variant, p, d, p_min, p_max):
    box = np.concatenate([np.asarray(p_min, dtype=np.float64).reshape(3), np.asarray(p_max, dtype=np.float64).reshape(3)])
    flag, near, far = ops.aabb_child_pairs(variant, np.asarray(p, dtype=np.float64),
                                           np.asarray(d, dtype=np.float64).reshape(1, 3), box.reshape(1, 6))
    return bool(flag[0, 0].item()), float(near[0, 0].item()), float(far[0, 0].item())


def compute_far_bound0406(p, d, p_min, p_max):
    """ipb2dmapping.py:82-114.  Raises IndexError like the reference when fewer than two faces are hit."""
    ok, near, far = _one(406, p, d, p_min, p_max)
    if not ok:
        raise IndexError("list index out of range")
    return near, far


def compute_far_bound0606(p, d, p_min, p_max):
    """ipb2dmapping.py:119-172 -> (intersect, near, far)."""
    return _one(606, p, d, p_min, p_max)


def find_aabb_box(points, aabb_list, query_point):
    """ipb2dmapping.py:174-197 -> (True, index) or (False, None).  Raises ValueError when fewer than 10 boxes exist
    (sklearn KDTree.query(k=10) does)."""
    idx = int(ops.aabb_find_box(points, aabb_list, np.asarray(query_point, dtype=np.float64).reshape(1, 3), 10)[0].item())
    return (False, None) if idx < 0 else (True, idx)


def pack_train_rays(origin, points, centres, child_bounds, child_bounds_bigger, parent_box, surface_expand,
                    variant="maicity", dir_vec=None, dist_vec=None):
    """Batched loop body of ipb2dmapping.py:367-397 (MaiCity) / :736-768 (KITTI) + the record of :447-452.
    parent_box = (x_min, x_max, y_min, y_max, z_min, z_max).  Returns ((N,15) fp32 CUDA tensor, keep mask)."""
    origin = np.asarray(origin, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    if dir_vec is None:
        vec = points - origin                                     # ipb2dmapping.py:341-343
        dist_vec = np.linalg.norm(vec, axis=1)
        dir_vec = vec / dist_vec[:, None]
    return ops.aabb_pack_train(406 if variant == "maicity" else 606, origin, dir_vec, dist_vec, points, centres,
                               child_bounds, child_bounds_bigger, parent_box, surface_expand, 10)
