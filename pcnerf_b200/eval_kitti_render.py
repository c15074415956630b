"""Hot-path part of eval_kitti_render.py: the AABB leaf functions (:170-244), the candidate-group builder
(:353-461 / :681-803), the group-aligned batch driver (:979-1030 / :1111-1161).  PCD / pose file IO is out of scope.
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


@torch.no_grad()
def render_frame(nof_coarse_model, nof_fine_model, embedding_position, dataset_rays, dataset_other, N_samples,
                 N_importance, chunk, depth_inference_method=2, batch_size_set=18432, use_disp=False, perturb=0,
                 noise_std=0):
    """The per-frame loop of eval_kitti_render.py:979-1030: batch, render, keep the rows flagged by the fine pass.
    Returns the rendered point cloud (M,3) on the device."""
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
