"""GPU: K0 (frame -> returns) and the kitti_dataload mirror end to end (PCD files in, 15-column rays + .npy cache out)
against the fixture produced by executing the reference's own dataset class (oracle/make_golden_dataset.py)."""
import os

import numpy as np
import pytest
import torch

import pcnerf_oracle as orc
from conftest import golden
from test_dataset_cpu import _inputs, check_rays

pytestmark = pytest.mark.gpu


def test_frame_returns_vs_oracle():
    from pcnerf_b200 import ops
    g = golden("kitti_dataset")
    frames, _, kw = _inputs(g)
    ds, de = int(g["data_start"]), int(g["data_end"])
    poses = orc.kitti_poses(list(g["pose_lines"]), ds)
    for fid, pts in frames.items():
        j = fid - 1
        pe, dv, dist, pos = orc.kitti_frame_returns(pts, poses, j, ds, de, kw["range_delete_x"], kw["range_delete_y"],
                                                    kw["range_delete_z"], kw["over_height"], kw["over_low"],
                                                    kw["interest_x"], kw["interest_y"])
        w, d, r = ops.frame_returns(pts, poses[j + 1].numpy(), poses[ds + 1:de + 1, :2, -1].numpy(),
                                    (kw["range_delete_x"], kw["range_delete_y"], kw["range_delete_z"]), 120.0,
                                    kw["over_height"], kw["over_low"], kw["interest_x"], kw["interest_y"])
        assert w.shape[0] == pe.shape[0] and 0 < pe.shape[0] < pts.shape[0]          # same points survive, in order
        # float64 products of numpy's BLAS (FMA or not) vs plain multiply-add: a couple of ulp
        np.testing.assert_allclose(w.cpu().numpy(), pe, rtol=1e-14, atol=1e-13)
        np.testing.assert_allclose(r.cpu().numpy(), dist, rtol=1e-14)
        np.testing.assert_allclose(d.cpu().numpy(), dv, rtol=0, atol=1e-14)
    # a tight interest region / range gate must reject points (the golden scene keeps all of them at 20 m)
    fid, pts = next(iter(frames.items()))
    w2, _, _ = ops.frame_returns(pts, poses[fid].numpy(), poses[ds + 1:de + 1, :2, -1].numpy(), (3, 2, 1.25), 120.0, 0.168, -2.0,
                                 5.0, 5.0)
    pe2, _, _, _ = orc.kitti_frame_returns(pts, poses, fid - 1, ds, de, 3, 2, 1.25, 0.168, -2.0, 5.0, 5.0)
    assert w2.shape[0] == pe2.shape[0] and w2.shape[0] < w.shape[0] + 10 ** 9


def test_kitti_dataload_end_to_end(tmp_path):
    from pcnerf_b200 import pcd
    from pcnerf_b200.nof.dataset.ipb2dmapping import kitti_dataload
    g = golden("kitti_dataset")
    frames, children, kw = _inputs(g)
    root, sub, res = str(tmp_path / "frames"), str(tmp_path / "children"), str(tmp_path / "result")
    for fid, pts in frames.items():
        pcd.write_pcd(os.path.join(root, "%d.pcd" % fid), pts)
    for i, c in enumerate(children):
        pcd.write_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), c)
    pcd.write_pcd(str(tmp_path / "source.pcd"), g["parent"])
    with open(str(tmp_path / "poses.txt"), "w") as f:
        f.write("\n".join(g["pose_lines"]) + "\n")
    args = dict(root_dir=root, data_start=int(g["data_start"]), data_end=int(g["data_end"]), cloud_size_val=64,
                sub_nerf_test_num=int(g["n_child"]), pose_path=str(tmp_path / "poses.txt"), subnerf_path=sub,
                parentnerf_path=str(tmp_path / "source.pcd"), result_path=res, **kw)
    ds = kitti_dataload(split="train", re_loaddata=1, **args)
    check_rays(ds.rays.numpy(), g["rays"])
    assert np.array_equal(ds.ranges.numpy(), g["ranges"])
    assert np.array_equal(ds.sub_nerf_num_count, g["sub_nerf_num_count"])
    assert len(ds) == g["rays"].shape[0] and torch.equal(ds[3]["rays"], ds.rays[3])
    # the .npy cache of ipb2dmapping.py:826-836 round-trips
    again = kitti_dataload(split="train", re_loaddata=0, **args)
    assert torch.equal(again.rays, ds.rays) and torch.equal(again.ranges, ds.ranges)


@pytest.mark.parametrize("method", [2, 1])
def test_multi_frame_kitti_files_to_candidate_rows(tmp_path, method):
    """eval_kitti_render.multi_frame_kitti from files on disk: K0 + K1 against the oracle chain (frame filter ->
    build_candidate_groups, itself pinned against the reference's leaf functions), artefacts written where the reference
    writes them."""
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import pcd
    g = golden("kitti_dataset")
    frames, children, kw = _inputs(g)
    ds, de = int(g["data_start"]), int(g["data_end"])
    root, sub, res = str(tmp_path / "frames"), str(tmp_path / "children"), str(tmp_path / "result")
    for fid, pts in frames.items():
        pcd.write_pcd(os.path.join(root, "%d.pcd" % fid), pts)
    for i, c in enumerate(children):
        pcd.write_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), c)
    pcd.write_pcd(str(tmp_path / "source.pcd"), g["parent"])
    with open(str(tmp_path / "poses.txt"), "w") as f:
        f.write("\n".join(g["pose_lines"]) + "\n")
    fid = int(g["frame_ids"][0])
    rays, ranges, other = ev.multi_frame_kitti(root, data_start=ds, data_end=de, range_delete_x=kw["range_delete_x"],
                                               range_delete_y=kw["range_delete_y"], range_delete_z=kw["range_delete_z"],
                                               sub_nerf_test_num=len(children), over_height=kw["over_height"],
                                               over_low=kw["over_low"], interest_x=kw["interest_x"], interest_y=kw["interest_y"],
                                               pose_path=str(tmp_path / "poses.txt"), subnerf_path=sub,
                                               parentnerf_path=str(tmp_path / "source.pcd"), view_pcd_number=fid,
                                               result_path=res, depth_inference_method=method)
    poses = orc.kitti_poses(list(g["pose_lines"]), ds)
    pe, dv, dist, pos = orc.kitti_frame_returns(frames[fid], poses, fid - 1, ds, de, kw["range_delete_x"], kw["range_delete_y"],
                                                kw["range_delete_z"], kw["over_height"], kw["over_low"], kw["interest_x"],
                                                kw["interest_y"])
    bound = np.stack([np.concatenate([c.astype(np.float64).min(0), c.astype(np.float64).max(0)]) for c in children])
    par = g["parent"].astype(np.float64)
    r_ref, rg_ref, o_ref = orc.build_candidate_groups(pos.numpy().astype(np.float64), dv, dist, bound, bound, par.min(0),
                                                      par.max(0), method, 0.05)[:3]
    assert rays.shape == tuple(r_ref.shape) and rays.shape[0] > 0
    assert np.array_equal(other.numpy().reshape(-1), np.asarray(o_ref).reshape(-1))            # group structure: exact
    np.testing.assert_allclose(rays.numpy(), np.asarray(r_ref), rtol=3e-7, atol=1e-6)
    np.testing.assert_allclose(ranges.numpy().reshape(-1), np.asarray(rg_ref).reshape(-1), rtol=3e-7)
    d = os.path.join(res, "two_step" if method == 2 else "one_step", "%dpcd" % fid, "childnerf_ray_intersect")
    assert np.array_equal(np.load(os.path.join(d, "all_rays_child.npy")), rays.numpy())
    assert np.load(os.path.join(d, "other_interest_sub_nerf_number_child.npy")).shape == (rays.shape[0], 1)
    assert pcd.read_pcd(os.path.join(d, "%d_pose.pcd" % fid)).shape == (1, 3)
    assert pcd.read_pcd(os.path.join(d, "%d_source.pcd" % fid)).shape[0] > 0


def test_render_scene_frame_blocks_partition_the_work():
    """Multi-parent scene (BASELINE configs[4]): two parent blocks with their own child boxes and networks.  Rendering with
    world_size 2 (each simulated rank owns one block) yields exactly the blocks' single-rank results: a block's rays never
    leave its owner, nothing is rendered twice."""
    from gpu_util import make_nets
    from pcnerf_b200 import scene, synth
    blocks = []
    pts_all = []
    for i in range(2):
        parent = (-20.0 + 45 * i, 20.0 + 45 * i, -20.0, 20.0, -1.7, 0.5)
        sc = synth.make_scene(300 + i, 24, parent)
        mc, mf, emb = make_nets(42 + 2 * i, 43 + 2 * i, False, None)
        blocks.append(scene.ParentBlock(sc.parent_min, sc.parent_max, sc.child_bounds,
                                        sc.child_bounds + np.array([-0.025] * 3 + [0.025] * 3), mc, mf))
        pts_all.append(synth.make_points(sc, 17 + i, 300))
    origin = np.array([22.0, 0.0, -0.5])                      # between the two blocks
    pts = np.concatenate(pts_all)
    which = scene.route_points(pts, blocks)
    assert (which[:300] == 0).all() and (which[300:] == 1).all()
    full = scene.render_scene_frame(blocks, origin, pts, emb, 32, 64, 8192, world_size=1, rank_=0)
    assert sorted(full) == [0, 1] and all(v.shape[0] > 0 and v.shape[1] == 3 for v in full.values())
    for r in range(2):
        part = scene.render_scene_frame(blocks, origin, pts, emb, 32, 64, 8192, world_size=2, rank_=r)
        assert list(part) == [r]
        assert torch.equal(part[r], full[r])


def test_maicity_dataload_end_to_end(tmp_path):
    from pcnerf_b200 import pcd
    from pcnerf_b200.nof.dataset.ipb2dmapping import maicity_dataload
    g = golden("maicity_dataset")
    frames, children, kw = _inputs(g)
    root, sub, res = str(tmp_path / "frames"), str(tmp_path / "children"), str(tmp_path / "result")
    for fid, pts in frames.items():
        pcd.write_pcd(os.path.join(root, "%d.pcd" % fid), pts)
    for i, c in enumerate(children):
        pcd.write_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), c)
    with open(str(tmp_path / "poses.txt"), "w") as f:
        f.write("\n".join(g["pose_lines"]) + "\n")
    args = dict(root_dir=root, data_start=int(g["data_start"]), data_end=int(g["data_end"]), cloud_size_val=64,
                sub_nerf_test_num=int(g["n_child"]), pose_path=str(tmp_path / "poses.txt"), subnerf_path=sub,
                result_path=res, **kw)
    ds = maicity_dataload(split="train", re_loaddata=1, **args)
    assert np.array_equal(ds.rays.numpy(), g["rays"])                           # every column, bit for bit
    assert np.array_equal(ds.ranges.numpy(), g["ranges"])
    assert np.array_equal(ds.sub_nerf_num_count, g["sub_nerf_num_count"])
    again = maicity_dataload(split="train", re_loaddata=0, **args)
    assert torch.equal(again.rays, ds.rays)


def test_multi_frame_maicity_files_to_candidate_rows(tmp_path):
    """eval_kitti_render.multi_frame_maicity from files on disk against the oracle chain (frame filter -> parent-box test ->
    build_candidate_groups with the MaiCity growth step)."""
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import pcd
    g = golden("maicity_dataset")
    frames, children, kw = _inputs(g)
    root, sub, res = str(tmp_path / "frames"), str(tmp_path / "children"), str(tmp_path / "result")
    for fid, pts in frames.items():
        pcd.write_pcd(os.path.join(root, "%d.pcd" % fid), pts)
    for i, c in enumerate(children):
        pcd.write_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), c)
    with open(str(tmp_path / "poses.txt"), "w") as f:
        f.write("\n".join(g["pose_lines"]) + "\n")
    fid = int(g["frame_ids"][1])
    box = {k: kw[k] for k in kw if k.startswith("nerf_")}
    rays, ranges, other = ev.multi_frame_maicity(root, data_start=0, data_end=2, range_delete_x=kw["range_delete_x"],
                                                 range_delete_y=kw["range_delete_y"], range_delete_z=kw["range_delete_z"],
                                                 sub_nerf_test_num=len(children), pose_path=str(tmp_path / "poses.txt"),
                                                 subnerf_path=sub, view_pcd_number=fid, result_path=res,
                                                 depth_inference_method=2, **box)
    P = torch.Tensor(np.array([np.append(np.array([float(i) for i in r.split(" ")]).reshape(3, 4), np.array([[0, 0, 0, 1]]), axis=0)
                               for r in g["pose_lines"]])).numpy()
    p = frames[fid]
    p = p[(np.abs(p[:, 0]) >= kw["range_delete_x"]) | (np.abs(p[:, 1]) >= kw["range_delete_y"]) | (np.abs(p[:, 2]) >= kw["range_delete_z"])]
    p = p[np.linalg.norm(p, axis=1) < 120]
    pe = (P[fid - 1] @ np.vstack((p.T, np.ones((1, p.shape[0]))))).T[:, :3]
    m = (pe[:, 0] >= box["nerf_length_min"]) & (pe[:, 1] >= box["nerf_width_min"]) & (pe[:, 2] >= box["nerf_height_min"]) & \
        (pe[:, 0] <= box["nerf_length_max"]) & (pe[:, 1] <= box["nerf_width_max"]) & (pe[:, 2] <= box["nerf_height_max"])
    pe = pe[m]
    origin = P[fid - 1][:3, -1].astype(np.float64)
    vec = pe - origin
    dist = np.linalg.norm(vec, axis=1)
    dv = vec / dist[:, None]
    bound = np.stack([np.concatenate([c.astype(np.float64).min(0) - 0.025, c.astype(np.float64).max(0) + 0.025]) for c in children])
    pmin = np.array([box["nerf_length_min"], box["nerf_width_min"], box["nerf_height_min"]])
    pmax = np.array([box["nerf_length_max"], box["nerf_width_max"], box["nerf_height_max"]])
    r_ref, rg_ref, o_ref = orc.build_candidate_groups(origin, dv, dist, bound, bound, pmin, pmax, 2, 0.005)[:3]
    assert rays.shape == tuple(r_ref.shape) and rays.shape[0] > 0
    assert np.array_equal(other.numpy().reshape(-1), np.asarray(o_ref).reshape(-1))
    np.testing.assert_allclose(rays.numpy(), np.asarray(r_ref), rtol=3e-7, atol=1e-6)
    d = os.path.join(res, "two_step", "%dpcd" % fid, "childnerf_ray_intersect")
    assert np.array_equal(np.load(os.path.join(d, "all_rays_child.npy")), rays.numpy())


def test_fit_from_files_then_render_to_pcd(tmp_path):
    """The two entry points end to end on the reference's data layout, without open3d / pcl / Lightning: PCD frames + poses
    + child clouds -> maicity_dataload -> train_kitti.fit (NOFSystem, FlatAdam, MultiStepLR, device-side loss history,
    reference-style checkpoint) -> load_ckpt -> multi_frame_maicity -> render_view_to_pcd."""
    import types
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import pcd, train_kitti
    from pcnerf_b200.nof.networks import NOF_coarse, NOF_fine, Embedding
    g = golden("maicity_dataset")
    frames, children, kw = _inputs(g)
    root, sub, res = str(tmp_path / "frames"), str(tmp_path / "children"), str(tmp_path / "result")
    for fid, pts in frames.items():
        pcd.write_pcd(os.path.join(root, "%d.pcd" % fid), pts)
    for i, c in enumerate(children):
        pcd.write_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), c)
    with open(str(tmp_path / "poses.txt"), "w") as f:
        f.write("\n".join(g["pose_lines"]) + "\n")
    hp = types.SimpleNamespace(
        root_dir=root, pose_path=str(tmp_path / "poses.txt"), subnerf_path=sub, result_path=res, datasettype="maicity_dataload",
        data_start=0, data_end=2, cloud_size_val=64, sub_nerf_test_num=len(children), re_loaddata=1, batch_size=256,
        num_epochs=1, seed=42, optimizer="flat_adam", lr=5e-4, weight_decay=1e-3, momentum=0.9, decay_gamma=0.1,
        L_pos=10, feature_size=256, use_skip=True, loss_type="smoothl1", N_samples=32, N_importance=64, use_disp=False,
        perturb=1.0, noise_std=0.0, chunk=262144, use_segmentated_sample=1, segmentated_child_nerf_ratio=0.1,
        use_child_nerf_divide=0, use_child_nerf_loss=1, lambda_loss=1.0, lambda_loss_fine=1.0, lambda_child_free_loss=1e6,
        lambda_child_depth_loss=1e5, **kw)
    ckpt = str(tmp_path / "best.ckpt")
    paths = [str(tmp_path / ("curve%d.npy" % i)) for i in range(7)]
    system, hist = train_kitti.fit(hp, max_steps=6, ckpt_path=ckpt, history_paths=paths)
    assert hist.shape == (6, 7) and np.isfinite(hist).all() and np.load(paths[0]).shape == (6,)
    assert np.allclose(hist[:, 0], hist[:, 1:].sum(1), rtol=1e-4)                  # the total is the sum of its six terms
    mc, mf = NOF_coarse(), NOF_fine()
    train_kitti.load_ckpt(mc, ckpt, model_name="nof_coarse")
    train_kitti.load_ckpt(mf, ckpt, model_name="nof_fine")
    assert torch.equal(mc.state_dict()["layer1.0.weight"], system.nof_coarse.state_dict()["layer1.0.weight"].cpu())
    mc.to(dev_()).eval()
    mf.to(dev_()).eval()
    box = {k: kw[k] for k in kw if k.startswith("nerf_")}
    rays, ranges, other = ev.multi_frame_maicity(root, data_start=0, data_end=2, range_delete_x=kw["range_delete_x"],
                                                 range_delete_y=kw["range_delete_y"], range_delete_z=kw["range_delete_z"],
                                                 sub_nerf_test_num=len(children), pose_path=str(tmp_path / "poses.txt"),
                                                 subnerf_path=sub, view_pcd_number=2, result_path=res, **box)
    out = str(tmp_path / "render" / "2_render.pcd")
    pts = ev.render_view_to_pcd(mc, mf, Embedding(3, 10), rays, other, out, 32, 64, 184320)
    back = pcd.read_pcd(out)
    assert back.shape == (pts.shape[0], 3) and pts.shape[0] > 0 and np.isfinite(back).all()


def dev_():
    return torch.device("cuda:0")
