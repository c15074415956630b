#!/usr/bin/env python
"""GPU: K1 (AABB stage) at the child-box counts of the shipped scenes (K = 15,333 KITTI, 5,729 MaiCity)
and at the benchmark's K = 200: training ray packing (k_pack_train: k = 10 nearest centres, exact KD-tree order) and the
depth-inference candidate-group builder (k_groups_count / k_groups_fill: O(N K) fp64 scan) for one 131,072-ray frame.
Prints one JSON line (CUDA-event times, median of 5)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcnerf_b200 import ops, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    n = int(os.environ.get("RAYS", 131072))
    out = {"rays": n}
    dev = torch.device("cuda:0")
    f64 = dict(dtype=torch.float64, device=dev)
    for name, K, parent in (("bench_K200", 200, synth.KITTI_PARENT), ("maicity_K5729", 5729, synth.MAICITY_PARENT),
                            ("kitti_K15333", 15333, synth.KITTI_PARENT)):
        scene = synth.make_scene(4242, K, parent)
        pts = synth.make_points(scene, 6, n)
        dirs, dist = synth.rays_from_points(scene.origin, pts)
        o, d, r, p = (torch.tensor(x, **f64) for x in (scene.origin, dirs, dist, pts))
        c, b, bb = (torch.tensor(x, **f64) for x in (scene.centres, scene.child_bounds, scene.child_bounds_bigger))
        sbl = torch.tensor(scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3), **f64)
        t_pack = timed(lambda: ops.aabb_pack_train(606, o, d, r, p, c, b, bb, scene.parent, 0.05, 10, compact=False))
        m = n if K <= 5729 else n // 8          # the dense synthetic K = 15,333 scene yields > 100 candidate rows per ray
        rows = {}

        def groups(grid=None):
            rows["n"] = ops.aabb_build_groups(o, d[:m], r[:m], b, sbl, scene.parent_min, scene.parent_max, 2, 0.05, 0.65,
                                              grid=grid)[0].shape[0]

        t_groups = timed(groups, reps=3)                       # default: uniform grid over the box centres from K >= 1024
        t_scan = timed(lambda: groups(False), reps=3)          # the reference's scan of every box
        out[name] = {"K": K, "pack_train_ms": t_pack, "pack_train_ns_per_ray_box": 1e6 * t_pack / (n * K),
                     "groups_rays": m, "groups_ms": t_groups, "groups_full_scan_ms": t_scan, "groups_rows": rows["n"],
                     "groups_full_scan_ns_per_ray_box": 1e6 * t_scan / (m * K)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
