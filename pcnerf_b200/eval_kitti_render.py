"""Hot-path part of eval_kitti_render.py: the AABB leaf functions (:170-244), the candidate-group builder
(:353-461 / :681-803), the group-aligned batch driver (:979-1030 / :1111-1161) and the file-level frame builder
multi_frame_kitti (:538-881) on pcnerf_b200.pcd (no open3d / pcl).
"""
import numpy as np
import torch

from . import ops
from .nof.render import render_rays_view_0525_2_2


def compute_far_bound0429(p, d, p_min, p_max):
    """eval_kitti_render.py:170-211 -> (intersectFlag, near, far) for one ray / one box."""
    box = np.concatenate([np.asarray(p_min, dtype=np.float64).reshape(3), np.asarray(p_max, dtype=np.float64).reshape(3)])
    flag, near, far = ops.aabb_child_pairs(429, np.asarray(p, dtype=np.float64),
                                           np.asarray(d, dtype=np.float64).reshape(1, 3), box.reshape(1, 6))
    return bool(flag[0, 0].item()), float(near[0, 0].item()), float(far[0, 0].item())


def ray_aabb_distances(ray_origin, ray_dirs, aabb_min, aabb_max):
    """eval_kitti_render.py:213-235 -> (N,) numpy float64."""
    return ops.aabb_slab(ray_origin, ray_dirs, aabb_min, aabb_max).cpu().numpy()


def distance_to_ray(ray_origin, ray_dir, points):
    """eval_kitti_render.py:237-244 -> (K,) numpy float64 for one ray."""
    o = ray_origin.numpy() if isinstance(ray_origin, torch.Tensor) else np.asarray(ray_origin)
    return ops.aabb_dist_to_ray(o, np.asarray(ray_dir, dtype=np.float64).reshape(1, 3), points)[0].cpu().numpy()


def build_test_rays(origin, dir_vec, dist_vec, sub_nerf_bound, sub_nerf_bound_larger, parent_min, parent_max,
                    depth_inference_method=2, dataset="maicity"):
    """Batched per-ray loop of multi_frame_maicity (:353-461, grow step 0.005) / multi_frame_kitti (:681-803, 0.05).
    Returns (all_rays (N',13) f32, all_ranges (N',1) f32, other_interest_sub_nerf_number (N',1) i64) on the device."""
    grow = 0.005 if dataset == "maicity" else 0.05
    rays, ranges, other, _ = ops.aabb_build_groups(origin, dir_vec, dist_vec, sub_nerf_bound, sub_nerf_bound_larger,
                                                   parent_min, parent_max, depth_inference_method, grow, 0.65)
    return rays, ranges, other


def eval_batches(dataset_rays_last_col, batch_size_set):
    """Group-aligned batching of eval_kitti_render.py:979-1005 / :1111-1136 (never split a candidate group: follower
    rows carry -1 in the last column).  Takes the last column as a host array; returns [(start, stop)]."""
    tag = np.asarray(dataset_rays_last_col).reshape(-1)
    n = tag.shape[0]
    out, i = [], 0
    while i < n:
        if i == n - 1:
            break
        if i + batch_size_set < n - 0.5 * batch_size_set:
            extra = 0
            while tag[i + batch_size_set + extra] < -0.5:
                extra += 1
                if i + batch_size_set + extra == n:
                    break
            out.append((i, i + batch_size_set + extra))
            i = i + batch_size_set + extra
        else:
            out.append((i, n))
            i = n
    return out


class FramePlan(ops.GroupPlan):
    """ops.GroupPlan of one frame's candidate rows + the number of rows the reference's batch loop renders (device scalar).
    Built once per frame by whoever produces the rows (the group builder K1 / a loaded .npy cache); costs one host sync."""

    def __init__(self, dataset_rays, dataset_other, batch_size_set):
        super().__init__(dataset_rays, dataset_other)
        self.batch_size_set = int(batch_size_set)
        self.n_rendered = ops.eval_rows_rendered(dataset_rays, batch_size_set)


FRAME_ENC_BYTES = 6 << 30       # encodings of one ray batch of the grouped frame path (fp32: 256 B, fp16: 128 B per sample)


@torch.no_grad()
def render_frame(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other, N_samples,
                 N_importance, chunk, depth_inference_method=2, batch_size_set=18432, use_disp=False, perturb=0,
                 noise_std=0, plan=None):
    """The per-frame loop of eval_kitti_render.py:979-1030: render the candidate rows, keep the rows flagged by the fine
    pass.  Returns the rendered point cloud (M,3) on the device.

    The reference walks the rows in group-aligned batches of `batch_size_set` (an out-of-memory guard: with eval-mode
    BatchNorm no value depends on the batch boundaries) and evaluates every candidate row.  Here the frame is rendered
    ONCE PER PHYSICAL RAY (nof.render._view_grouped's argument: the rows of a group share their samples): coarse and fine
    MLP passes over the G physical rays in a few large ray batches, then one K5 pass over all N' rows (child masks, peak
    test, in-child depth, group winner).  No host synchronisation besides the final compaction; the one quirk of the
    reference loop that does depend on the batch walk (a lone trailing row is never rendered) is reproduced on the device
    (FramePlan.n_rendered).  With perturb != 0 (per-row random numbers) the per-row batch loop below is used."""
    from .nof import render as R
    if perturb != 0 or not R.GROUP_RAYS or dataset_rays.shape[0] == 0:
        return _render_frame_rows(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other,
                                  N_samples, N_importance, chunk, depth_inference_method, batch_size_set, use_disp, perturb,
                                  noise_std)
    if plan is None:
        plan = FramePlan(dataset_rays, dataset_other, batch_size_set)
    if not plan.uniform or plan.batch_size_set != int(batch_size_set):
        return _render_frame_rows(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other,
                                  N_samples, N_importance, chunk, depth_inference_method, batch_size_set, use_disp, perturb,
                                  noise_std)
    rays = dataset_rays.contiguous()
    dev = rays.device
    G, S, F = plan.G, int(N_samples), int(N_samples) + int(N_importance)
    tc_c, tc_f = nof_coarse_model.mlp_precision() == 1, nof_fine_model.mlp_precision() == 1
    hr = rays.index_select(0, plan.head_rows)                        # (G,13): one row per physical ray
    lazy_c, lazy_f = R._lazy(nof_coarse_model), R._lazy(nof_fine_model)      # closed-form engine: no encoding tensor at all
    per_ray = F * (8 if lazy_f else 128 if tc_f else 256)
    B = max(1, min(G, FRAME_ENC_BYTES // per_ray))
    zf = torch.empty((G, F), dtype=torch.float32, device=dev)
    pf = torch.empty((G, F), dtype=torch.float32, device=dev)
    for g0 in range(0, G, B):
        h = hr[g0:g0 + B]
        z, enc = ops.sample_encode_coarse(h, S, 0, 9, 10, 10, 11, False, 0.0, None, not lazy_c, tc_c)
        p = nof_coarse_model.forward_encoded(ops.LazyEnc(h, z) if lazy_c else enc, chunk).view(-1, S)
        del enc
        _, w, _, _, _ = ops.search_rows(p, z, h, 6, 7, 1e-10, depth_inference_method)     # the coarse weights (K5 arithmetic)
        zf_b, encf = ops.sample_encode_fine(h, z, w, int(N_importance), None, True, not lazy_f, tc_f)
        pf_b = nof_fine_model.forward_encoded(ops.LazyEnc(h, zf_b) if lazy_f else encf, chunk).view(-1, F)
        del encf
        if B >= G:
            zf, pf = zf_b, pf_b
        else:
            zf[g0:g0 + B].copy_(zf_b)
            pf[g0:g0 + B].copy_(pf_b)
    depth, _, _, peak, wsum = ops.search_rows(pf, zf, rays, 6, 7, 1e-10, depth_inference_method, row_ray=plan.row_ray,
                                              want_w=False)
    keep = ops.search_select(plan.other, peak, wsum, n_rendered=plan.n_rendered).reshape(-1)
    return ops.points(rays, depth)[keep]


@torch.no_grad()
def _render_frame_rows(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other, N_samples,
                       N_importance, chunk, depth_inference_method=2, batch_size_set=18432, use_disp=False, perturb=0,
                       noise_std=0):
    """The reference's loop as it stands (eval_kitti_render.py:979-1030): group-aligned batches, every candidate row."""
    tags = dataset_rays[:, -1].cpu().numpy()
    pts = []
    for a, b in eval_batches(tags, batch_size_set):
        rays = dataset_rays[a:b].contiguous()
        other = dataset_other[a:b].reshape(-1)
        res = render_rays_view_0525_2_2(nof_coarse_model, nof_fine_model, embedding_position, rays, other,
                                        N_samples=N_samples, N_importance=N_importance, use_disp=use_disp,
                                        perturb=perturb, noise_std=noise_std, chunk=chunk,
                                        depth_inference_method=depth_inference_method)
        keep = res['rays_effective_flag_fine'].reshape(-1).bool()
        pts.append(res['points_inference_fine'][keep])
    return torch.cat(pts, 0) if pts else torch.zeros((0, 3), device=dataset_rays.device)


def multi_frame_kitti(root_dir, split='test', data_start=1439, data_end=1510, range_delete_x=2, range_delete_y=1,
                      range_delete_z=0.5, sub_nerf_test_num=4, over_height=0.168, over_low=-2, interest_x=12, interest_y=10,
                      pose_path=None, subnerf_path=None, parentnerf_path=None, view_pcd_number=0, result_path=None,
                      depth_inference_method=2):
    """eval_kitti_render.py:538-881 with the reference's signature: candidate rows of frame `view_pcd_number` from the files
    on disk -- PCD / pose IO by pcnerf_b200.pcd instead of open3d / pcl, the point filters and the pose transform by K0
    (:641-693; the range gate is `< 120` here, `<= 120` in the training loader), the per-ray group construction by K1
    (:695-803).  Writes the reference's artefacts under result_path/{one,two}_step/<frame>pcd/childnerf_ray_intersect/
    (all_rays_child.npy, all_ranges_child.npy, other_interest_sub_nerf_number_child.npy, <frame>_source.pcd,
    <frame>_pose.pcd) and returns (all_rays (N',13) f32, all_ranges (N',1) f32, other (N',1) i64) on the CPU like it."""
    import os
    from . import pcd
    lo, hi = pcd.axis_aligned_bounds(pcd.read_pcd(parentnerf_path))
    poses = pcd.read_kitti_poses(pose_path, data_start)
    bound = np.zeros((sub_nerf_test_num, 6))
    for i in range(sub_nerf_test_num):                                             # :590-604 (extend_tmp = 0)
        blo, bhi = pcd.axis_aligned_bounds(pcd.read_pcd(os.path.join(subnerf_path, "%d.pcd" % (i + 1))))
        bound[i, :3], bound[i, 3:] = blo, bhi
    j = view_pcd_number - 1
    if not (data_start <= j < data_end):
        raise ValueError("view_pcd_number %d is outside [data_start+1, data_end]" % view_pcd_number)
    pts = pcd.read_pcd(os.path.join(root_dir, "%d.pcd" % view_pcd_number))
    strict_120 = float(np.nextafter(np.float32(120.0), np.float32(0.0)))           # r < 120  <=>  r <= prev(120) in float32
    world, dirs, dist = ops.frame_returns(pts, poses[j + 1], poses[data_start + 1:data_end + 1, :2, -1],
                                          (range_delete_x, range_delete_y, range_delete_z), strict_120, over_height,
                                          over_low, interest_x, interest_y)
    origin = poses[j + 1][:3, -1].astype(np.float64)
    rays, ranges, other, kept = ops.aabb_build_groups(origin, dirs, dist, bound, bound, lo, hi, depth_inference_method,
                                                      0.05, 0.65)
    rays, ranges, other = rays.cpu(), ranges.cpu(), other.cpu()
    if result_path:
        d = os.path.join(result_path, "two_step" if depth_inference_method == 2 else "one_step",
                         "%dpcd" % view_pcd_number, "childnerf_ray_intersect")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "all_ranges_child.npy"), ranges.numpy())
        np.save(os.path.join(d, "all_rays_child.npy"), rays.numpy())
        np.save(os.path.join(d, "other_interest_sub_nerf_number_child.npy"), other.numpy())
        pcd.write_pcd(os.path.join(d, "%d_source.pcd" % view_pcd_number), world[kept.bool()].cpu().numpy())
        pcd.write_pcd(os.path.join(d, "%d_pose.pcd" % view_pcd_number), origin.reshape(1, 3))
    return rays, ranges, other


def render_view_to_pcd(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other, out_path,
                       N_samples, N_importance, chunk, depth_inference_method=2, batch_size_set=18432):
    """The frame loop of the reference's __main__ (:979-1044): render the candidate rows in group-aligned batches, keep the
    rows flagged by the fine pass, write the rendered cloud as a binary PCD.  Returns the (M,3) points (device)."""
    from . import pcd
    dev = next(nof_coarse_model.parameters()).device
    pts = render_frame(nof_coarse_model, nof_fine_model, embedding_position, torch.as_tensor(dataset_rays).to(dev),
                       torch.as_tensor(dataset_other).to(dev), N_samples, N_importance, chunk,
                       depth_inference_method=depth_inference_method, batch_size_set=batch_size_set)
    pcd.write_pcd(out_path, pts.cpu().numpy())
    return pts


def multi_frame_maicity(root_dir, split='test', data_start=1, data_end=2, range_delete_x=2, range_delete_y=1,
                        range_delete_z=0.5, sub_nerf_test_num=4, nerf_length_min=-4.5, nerf_length_max=25.5,
                        nerf_width_min=-12, nerf_width_max=12, nerf_height_min=-2, nerf_height_max=0.5, pose_path=None,
                        subnerf_path=None, view_pcd_number=0, result_path=None, depth_inference_method=2):
    """eval_kitti_render.py:246-535 with the reference's signature: the MaiCity counterpart of multi_frame_kitti -- raw
    poses (frame file j+1 uses pose j), child boxes +-0.025 (:277-292), `< 120 m` gate, closed parent-box test on the
    transformed points (:343-345), float32 sensor position (:265), group growth step 0.005 (:353-461).  Same artefacts."""
    import os
    from . import pcd
    with open(pose_path, "r", encoding="utf-8") as f:
        rows = [r.strip() for r in f.readlines() if r.strip()]
    poses = torch.Tensor(np.array([np.append(np.array([float(i) for i in r.split(' ')]).reshape(3, 4),
                                             np.array([[0, 0, 0, 1]]), axis=0) for r in rows])).numpy()
    bound = np.zeros((sub_nerf_test_num, 6))
    for i in range(sub_nerf_test_num):
        blo, bhi = pcd.axis_aligned_bounds(pcd.read_pcd(os.path.join(subnerf_path, "%d.pcd" % (i + 1))))
        bound[i, :3], bound[i, 3:] = blo - 0.025, bhi + 0.025
    j = view_pcd_number - 1
    if not (data_start <= j < data_end):
        raise ValueError("view_pcd_number %d is outside [data_start+1, data_end]" % view_pcd_number)
    pts = pcd.read_pcd(os.path.join(root_dir, "%d.pcd" % view_pcd_number))
    box = (nerf_length_min, nerf_length_max, nerf_width_min, nerf_width_max, nerf_height_min, nerf_height_max)
    strict_120 = float(np.nextafter(np.float32(120.0), np.float32(0.0)))
    world, dirs, dist = ops.frame_returns(pts, poses[j], None, (range_delete_x, range_delete_y, range_delete_z), strict_120,
                                          float("inf"), -float("inf"), 0.0, 0.0, parent_box=box)
    origin = poses[j][:3, -1].astype(np.float64)
    pmin = np.array([nerf_length_min, nerf_width_min, nerf_height_min], dtype=np.float64)
    pmax = np.array([nerf_length_max, nerf_width_max, nerf_height_max], dtype=np.float64)
    rays, ranges, other, kept = ops.aabb_build_groups(origin, dirs, dist, bound, bound, pmin, pmax, depth_inference_method,
                                                      0.005, 0.65)
    rays, ranges, other = rays.cpu(), ranges.cpu(), other.cpu()
    if result_path:
        d = os.path.join(result_path, "two_step" if depth_inference_method == 2 else "one_step",
                         "%dpcd" % view_pcd_number, "childnerf_ray_intersect")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "all_ranges_child.npy"), ranges.numpy())
        np.save(os.path.join(d, "all_rays_child.npy"), rays.numpy())
        np.save(os.path.join(d, "other_interest_sub_nerf_number_child.npy"), other.numpy())
        pcd.write_pcd(os.path.join(d, "%d_source.pcd" % view_pcd_number), world[kept.bool()].cpu().numpy())
        pcd.write_pcd(os.path.join(d, "%d_pose.pcd" % view_pcd_number), origin.reshape(1, 3))
    return rays, ranges, other
