"""GPU: K1 AABB kernels through the C ABI, bit-exact against the golden vectors (reference) and the oracle."""
import numpy as np
import pytest
import torch

import pcnerf_oracle as orc
from conftest import golden

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def test_leaf_functions_bit_exact_vs_reference():
    from pcnerf_b200 import ops
    g = golden("aabb_leaf")
    o, dirs, bb = g["origin"], g["dirs"], g["child_bounds_bigger"]
    x_min, x_max, y_min, y_max, z_min, z_max = g["parent"]
    assert np.array_equal(_np(ops.aabb_far_bound(o, dirs, x_max, x_min, y_max, y_min, z_max, z_min)), g["far_parent"])
    idx = _np(ops.aabb_find_box(g["centres"], g["child_bounds"], g["points"], 10))
    assert np.array_equal(idx, g["idx"]) and np.array_equal(idx >= 0, g["inside"])
    for v, fk, nk, rk in ((429, "f0429", "n0429", "r0429"), (606, "f0606", "n0606", "r0606")):
        f, n, r = ops.aabb_child_pairs(v, o, dirs, bb)
        assert np.array_equal(_np(f), g[fk]) and np.array_equal(_np(n), g[nk]) and np.array_equal(_np(r), g[rk])
    f, n, r = ops.aabb_child_pairs(406, o, dirs, bb)
    assert np.array_equal(_np(n), g["n0406"], equal_nan=True) and np.array_equal(_np(r), g["r0406"], equal_nan=True)
    pmin, pmax = np.array([x_min, y_min, z_min]), np.array([x_max, y_max, z_max])
    assert np.array_equal(_np(ops.aabb_slab(o, dirs, pmin, pmax)), g["slab"])
    centre = (g["child_bounds"][:, :3] + g["child_bounds"][:, 3:]) / 2
    assert np.array_equal(_np(ops.aabb_dist_to_ray(o, dirs[:64], centre)), g["dist_to_ray"], equal_nan=True)


def test_scalar_mirrors_and_errors():
    from pcnerf_b200.nof.dataset import ipb2dmapping as ipb
    from pcnerf_b200 import eval_kitti_render as ev
    g = golden("aabb_leaf")
    o, d, bb = g["origin"], g["dirs"][3], g["child_bounds_bigger"][5]
    x_min, x_max, y_min, y_max, z_min, z_max = g["parent"]
    assert ipb.compute_far_bound(o, d, x_max, x_min, y_max, y_min, z_max, z_min) == g["far_parent"][3]
    assert ipb.compute_far_bound(o, np.zeros(3), x_max, x_min, y_max, y_min, z_max, z_min) is None
    assert ipb.compute_far_bound0606(o, d, bb[:3], bb[3:]) == (bool(g["f0606"][3, 5]), g["n0606"][3, 5], g["r0606"][3, 5])
    assert ev.compute_far_bound0429(o, d, bb[:3], bb[3:]) == (bool(g["f0429"][3, 5]), g["n0429"][3, 5], g["r0429"][3, 5])
    k = int(np.argmax(g["f0429"][3]))
    bk = g["child_bounds_bigger"][k]
    assert ipb.compute_far_bound0406(o, d, bk[:3], bk[3:]) == (g["n0406"][3, k], g["r0406"][3, k])
    miss = int(np.argmin(g["f0606"][3]))
    with pytest.raises(IndexError):
        ipb.compute_far_bound0406(o, d, g["child_bounds_bigger"][miss][:3], g["child_bounds_bigger"][miss][3:])
    ok, i = ipb.find_aabb_box(g["centres"], g["child_bounds"], g["points"][1])
    assert (ok, -1 if i is None else i) == (bool(g["inside"][1]), int(g["idx"][1]))
    with pytest.raises(ValueError):      # sklearn KDTree.query(k=10) with 8 training points
        ipb.find_aabb_box(g["centres"][:8], g["child_bounds"][:8], g["points"][1])


def test_pack_train_bit_exact_vs_reference():
    from pcnerf_b200.nof.dataset import ipb2dmapping as ipb
    g = golden("aabb_leaf")
    for variant in ("maicity", "kitti"):
        gp = golden("aabb_pack_" + variant)
        rays, keep = ipb.pack_train_rays(g["origin"], g["points"], g["centres"], g["child_bounds"],
                                         g["child_bounds_bigger"], tuple(g["parent"]), float(gp["surface_expand"]),
                                         variant, dir_vec=g["dirs"], dist_vec=g["dist"])
        assert np.array_equal(_np(rays), gp["rays"], equal_nan=True)


def test_groups_bit_exact_vs_reference():
    from pcnerf_b200 import eval_kitti_render as ev
    g = golden("aabb_leaf")
    x_min, x_max, y_min, y_max, z_min, z_max = g["parent"]
    pmin, pmax = np.array([x_min, y_min, z_min]), np.array([x_max, y_max, z_max])
    sbl = g["child_bounds"] + np.array([-0.025] * 3 + [0.025] * 3)
    for method in (2, 1):
        for grow, ds in ((0.005, "maicity"), (0.05, "kitti")):
            gg = golden("aabb_groups_m%d_g%s" % (method, str(grow).replace(".", "p")))
            n = int(gg["nray"])
            rays, ranges, other = ev.build_test_rays(g["origin"], g["dirs"][:n], g["dist"][:n], g["child_bounds"], sbl,
                                                     pmin, pmax, method, ds)
            assert np.array_equal(_np(rays), gg["rays"])
            assert np.array_equal(_np(ranges), gg["ranges"]) and np.array_equal(_np(other), gg["other"])


@pytest.mark.parametrize("K,n", [(200, 4000), (10, 257), (600, 1500)])
def test_larger_scene_bit_exact_vs_oracle(K, n):
    from pcnerf_b200 import ops, synth
    scene = synth.make_scene(100 + K, K, synth.KITTI_PARENT if K == 200 else synth.MAICITY_PARENT)
    pts = synth.make_points(scene, 5, n)
    rng = np.random.default_rng(K)
    pts[::7] = rng.uniform(scene.parent_min, scene.parent_max, size=pts[::7].shape)
    dirs, dist = synth.rays_from_points(scene.origin, pts)
    for variant, code in (("maicity", 406), ("kitti", 606)):
        ref, _ = orc.pack_train_rays_from_dirs(scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                               scene.child_bounds_bigger, scene.parent, 0.05, variant)
        got, _ = ops.aabb_pack_train(code, scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                     scene.child_bounds_bigger, scene.parent, 0.05, 10)
        assert np.array_equal(_np(got), ref, equal_nan=True)
    m = min(n, 600)
    sbl = scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3)
    for method in (2, 1):
        r_ref, rg_ref, o_ref, _ = orc.build_candidate_groups(scene.origin, dirs[:m], dist[:m], scene.child_bounds, sbl,
                                                             scene.parent_min, scene.parent_max, method, 0.005)
        r, rg, o, _ = ops.aabb_build_groups(scene.origin, dirs[:m], dist[:m], scene.child_bounds, sbl, scene.parent_min,
                                            scene.parent_max, method, 0.005, 0.65)
        assert np.array_equal(_np(r), r_ref) and np.array_equal(_np(rg), rg_ref) and np.array_equal(_np(o), o_ref)


def test_empty_inputs():
    from pcnerf_b200 import ops
    scene_boxes = np.zeros((12, 6))
    z3 = np.zeros((0, 3))
    assert ops.aabb_find_box(np.zeros((12, 3)), scene_boxes, z3, 10).shape[0] == 0
    r, rg, o, _ = ops.aabb_build_groups(np.zeros(3), z3, np.zeros(0), scene_boxes, scene_boxes, np.zeros(3), np.ones(3))
    assert r.shape == (0, 13) and o.shape == (0, 1)


@pytest.mark.parametrize("K,parent", [(3000, "kitti"), (15333, "kitti"), (5729, "maicity")])
def test_real_scale_box_count_bit_exact_vs_oracle(K, parent):
    """The child-box counts of the shipped scenes (15,333 boxes KITTI, 5,729 MaiCity; 3,000 = the first
    size beyond the shared-memory staging): the kernels read the boxes through L2 instead -- results must not change."""
    from pcnerf_b200 import ops, synth
    n = 300
    scene = synth.make_scene(4242, K, synth.KITTI_PARENT if parent == "kitti" else synth.MAICITY_PARENT)
    pts = synth.make_points(scene, 6, n)
    dirs, dist = synth.rays_from_points(scene.origin, pts)
    ref, _ = orc.pack_train_rays_from_dirs(scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                           scene.child_bounds_bigger, scene.parent, 0.05, "kitti")
    got, _ = ops.aabb_pack_train(606, scene.origin, dirs, dist, pts, scene.centres, scene.child_bounds,
                                 scene.child_bounds_bigger, scene.parent, 0.05, 10)
    assert np.array_equal(_np(got), ref, equal_nan=True)
    sbl = scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3)
    m = 120
    r_ref, rg_ref, o_ref, _ = orc.build_candidate_groups(scene.origin, dirs[:m], dist[:m], scene.child_bounds, sbl,
                                                         scene.parent_min, scene.parent_max, 2, 0.05)
    r, rg, o, _ = ops.aabb_build_groups(scene.origin, dirs[:m], dist[:m], scene.child_bounds, sbl, scene.parent_min,
                                        scene.parent_max, 2, 0.05, 0.65)
    assert np.array_equal(_np(r), r_ref) and np.array_equal(_np(rg), rg_ref) and np.array_equal(_np(o), o_ref)


@pytest.mark.parametrize("method", [2, 1])
@pytest.mark.parametrize("K,parent", [(1500, "kitti"), (5729, "maicity")])
def test_group_builder_grid_equals_full_scan(K, parent, method):
    """The uniform grid over the box centres (ops.BoxGrid) only restricts WHICH boxes a ray looks at (a superset of those the
    0.65 m prefilter keeps): rows, order, ranges and the side array must be identical to scanning every box -- also for vertical
    rays, rays whose origin lies outside the grid, rays that miss everything and NaN directions."""
    from pcnerf_b200 import ops, synth
    scene = synth.make_scene(99, K, synth.KITTI_PARENT if parent == "kitti" else synth.MAICITY_PARENT)
    n = 4000
    pts = synth.make_points(scene, 8, n)
    dirs, dist = synth.rays_from_points(scene.origin, pts)
    rng = np.random.default_rng(3)
    origins = np.tile(scene.origin, (n, 1))
    origins[:200] += rng.uniform(-30, 30, size=(200, 3)) * np.array([1, 1, 0.02])          # outside / elsewhere in the box
    dirs[200:230] = np.array([0.0, 0.0, -1.0])                                               # vertical
    dirs[230:260] = np.array([0.0, 1e-12, 1.0]) / np.linalg.norm([0.0, 1e-12, 1.0])
    d = rng.normal(size=(300, 3))
    dirs[260:560] = d / np.linalg.norm(d, axis=1, keepdims=True)                             # random: many miss everything
    dirs[560] = np.nan
    sbl = scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3)
    grow = 0.05 if parent == "kitti" else 0.005
    a = ops.aabb_build_groups(origins, dirs, dist, scene.child_bounds, sbl, scene.parent_min, scene.parent_max, method, grow,
                              0.65, grid=False)
    b = ops.aabb_build_groups(origins, dirs, dist, scene.child_bounds, sbl, scene.parent_min, scene.parent_max, method, grow,
                              0.65)
    assert a[0].shape[0] > n // 2
    for x, y in zip(a, b):
        assert torch.equal(x, y)
