"""Seeded synthetic scenes for tests and bench.  Pure numpy; no oracle, no CUDA.

A scene is one parent AABB, K child AABBs (1 m-pitch cells with random centres inside the parent,
half-extent U(0.25,0.5), padded by 0.025 m like nof/dataset/ipb2dmapping.py:265-271), one sensor origin near
z = 0 inside the parent, and LiDAR returns drawn uniformly inside uniformly chosen child boxes.
"""
import numpy as np

MAICITY_PARENT = (-12.0, 61.0, -12.0, 12.0, -2.0, 0.5)   # x_min,x_max,y_min,y_max,z_min,z_max (MaiCity00 shell script)
KITTI_PARENT = (-20.0, 20.0, -20.0, 20.0, -1.7, 0.5)      # 40 x 40 x 2.2 m KITTI-like block


class Scene:
    def __init__(self, parent, raw_bounds, origin):
        self.parent = tuple(float(v) for v in parent)
        self.parent_min = np.array([parent[0], parent[2], parent[4]], dtype=np.float64)
        self.parent_max = np.array([parent[1], parent[3], parent[5]], dtype=np.float64)
        self.raw_bounds = raw_bounds                                  # (K,6) min xyz, max xyz of the child point sets
        ext = 0.025
        self.child_bounds = raw_bounds + np.array([-ext] * 3 + [ext] * 3)            # sub_nerf_bound
        self.child_bounds_bigger = self.child_bounds.copy()                        # sub_nerf_bound_bigger (same 0.025)
        self.centres = (raw_bounds[:, :3] + raw_bounds[:, 3:]) / 2.0               # sub_nerf_center_point
        self.origin = origin

    @property
    def K(self):
        return self.raw_bounds.shape[0]


def make_scene(seed, K, parent=MAICITY_PARENT):
    rng = np.random.default_rng(seed)
    lo = np.array([parent[0], parent[2], parent[4]])
    hi = np.array([parent[1], parent[3], parent[5]])
    half = rng.uniform(0.25, 0.5, size=(K, 3))
    half[:, 2] = np.minimum(half[:, 2], 0.45 * (hi[2] - lo[2]))
    c = rng.uniform(lo + half + 0.05, hi - half - 0.05)
    raw = np.concatenate([c - half, c + half], axis=1)
    origin = np.array([rng.uniform(lo[0] + 2, min(lo[0] + 12, hi[0] - 2)), rng.uniform(-1, 1), rng.uniform(-0.3, 0.0)])
    origin = np.clip(origin, lo + 0.1, hi - 0.1)
    return Scene(parent, raw, origin)


def make_points(scene, seed, n):
    """LiDAR returns: uniform inside a uniformly chosen child box (never closer than 1 m to the sensor)."""
    rng = np.random.default_rng(seed + 7919)
    which = rng.integers(0, scene.K, size=n)
    b = scene.raw_bounds[which]
    pts = rng.uniform(b[:, :3], b[:, 3:])
    d = np.linalg.norm(pts - scene.origin, axis=1)
    bad = d < 1.0
    if bad.any():
        pts[bad] = pts[bad] + np.array([2.0, 0.0, 0.0])
        pts = np.clip(pts, scene.parent_min + 1e-3, scene.parent_max - 1e-3)
    return pts


def rays_from_points(origin, points):
    """dir/range per point, as nof/dataset/ipb2dmapping.py:341-343 (fp64)."""
    vec = points - origin
    dist = np.linalg.norm(vec, axis=1)
    return vec / dist[:, None], dist


def synth_train_rays(seed, n, K=8, parent=MAICITY_PARENT, surface_expand=0.05):
    """Cheap (N,15) fp32 training rays whose child interval brackets the return -- used where the AABB stage is
    not under test (bench MLP/compositing legs).  Columns per SURVEY.md section 3.4."""
    rng = np.random.default_rng(seed)
    scene = make_scene(seed, K, parent)
    pts = make_points(scene, seed, n)
    d, r = rays_from_points(scene.origin, pts)
    half = rng.uniform(0.3, 0.7, size=n)
    near_c = np.maximum(r - half, 0.05) - surface_expand
    far_c = r + half + surface_expand
    far_p = np.maximum(r + rng.uniform(1.0, 30.0, size=n), far_c)
    rays = np.zeros((n, 15), dtype=np.float64)
    rays[:, 0:3] = scene.origin
    rays[:, 3:6] = d
    rays[:, 7] = far_p
    rays[:, 8] = 3
    rays[:, 9] = rng.integers(1, K + 1, size=n)
    rays[:, 10] = near_c
    rays[:, 11] = far_c
    rays[:, 12] = r - surface_expand
    rays[:, 13] = far_c
    rays[:, 14] = r
    return rays.astype(np.float32)


GROUP_HIST = ((1, 0.215), (2, 0.359), (3, 0.183), (4, 0.090), (5, 0.054), (6, 0.04), (8, 0.03), (12, 0.02), (28, 0.009))


def synth_infer_rows(seed, n_phys, parent=MAICITY_PARENT):
    """(N',13) fp32 inference rows + `other` (N',) int64 with the shipped group-size histogram."""
    rng = np.random.default_rng(seed)
    sizes = np.array([g for g, _ in GROUP_HIST])
    probs = np.array([p for _, p in GROUP_HIST])
    probs = probs / probs.sum()
    n_per = rng.choice(sizes, size=n_phys, p=probs)
    scene = make_scene(seed, 16, parent)
    pts = make_points(scene, seed, n_phys)
    d, r = rays_from_points(scene.origin, pts)
    far_p = r + rng.uniform(2.0, 30.0, size=n_phys)
    rows, other = [], []
    for i in range(n_phys):
        n = int(n_per[i])
        nears = np.sort(rng.uniform(0.5, max(far_p[i] - 1.5, 1.0), size=n))
        if n > 0:
            nears[rng.integers(0, n)] = max(r[i] - 0.4, 0.1)
            nears = np.sort(nears)
        for k in range(n):
            rows.append([*scene.origin, *d[i], nears[k], nears[k] + rng.uniform(0.5, 1.2), 3.0, 0.0, far_p[i],
                         float(k + 1), float(n - 1) if k == 0 else -1.0])
            other.append(n - 1 if k == 0 else 0)
    return np.asarray(rows, dtype=np.float64).astype(np.float32), np.asarray(other, dtype=np.int64), r
