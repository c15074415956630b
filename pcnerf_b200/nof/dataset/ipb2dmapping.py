"""Drop-in for the AABB leaf functions of nof/dataset/ipb2dmapping.py (same names / argument meaning / error
behaviour) plus the batched ray-packing rule that the reference inlines in its dataset loaders.

The scalar-style functions accept what the reference accepts (numpy arrays / python scalars for ONE ray) and return
python / numpy values; they launch the same fp64 kernels as the batched entry point, so they are meant for parity
checks and drop-in compatibility -- use `pack_train_rays` for real work.

`kitti_dataload` is the reference's KITTI dataset class with the per-point python loops
replaced by two kernels per frame (K0 frame -> returns, K1 returns -> 15-column rays) and open3d / pcl replaced by
pcnerf_b200.pcd: same constructor arguments, same `rays` / `ranges`, same `.npy` cache files.
"""
import os

import numpy as np
import torch

from ... import ops, pcd


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



class kitti_dataload(torch.utils.data.Dataset):
    """nof/dataset/ipb2dmapping.py:516-836 (constructor signature, frame selection, outputs and cache paths as there).

    re_loaddata = 1: read the frames `root_dir/{j+1}.pcd`, the child clouds `subnerf_path/{i+1}.pcd` and the parent cloud,
    build `self.rays` (N,15) / `self.ranges` on the GPU and save them under `result_path/save_npy/split_child_nerf2_3/`;
    re_loaddata = 0: load that cache (:826-836).  Against the reference executed on the same files the child index, origins,
    directions and ranges are identical; the four bound columns (7, 10, 11, 13) agree to one float32 ulp -- the reference
    evaluates compute_far_bound0606 / compute_far_bound with a float32 tensor origin, i.e. partly in float32, this path in
    float64 (tests/test_gpu_dataset.py, fixture produced by oracle/make_golden_dataset.py running the reference class)."""

    def __init__(self, root_dir, split='train', data_start=1439, data_end=1510, cloud_size_val=2048,
                 range_delete_x=2, range_delete_y=1, range_delete_z=0.5, sub_nerf_test_num=3, surface_expand=0.1,
                 over_height=0.168, over_low=-2, interest_x=12, interest_y=12, pose_path=None, subnerf_path=None,
                 parentnerf_path=None, re_loaddata=0, result_path=None):
        super(kitti_dataload, self).__init__()
        self.root_dir, self.split, self.cloud_size_val = root_dir, split, cloud_size_val
        self.parentnerf_path, self.result_path, self.re_loaddata = parentnerf_path, result_path, re_loaddata
        if re_loaddata:
            self.data_start, self.data_end = data_start, data_end
            self.range_delete = (range_delete_x, range_delete_y, range_delete_z)
            self.sub_nerf_test_num, self.subnerf_path, self.surface_expand = sub_nerf_test_num, subnerf_path, surface_expand
            self.over_height, self.over_low, self.interest_x, self.interest_y = over_height, over_low, interest_x, interest_y
            lo, hi = pcd.axis_aligned_bounds(pcd.read_pcd(parentnerf_path))
            self.parent_box = (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
            self.poses = pcd.read_kitti_poses(pose_path, data_start)            # (n,4,4) float32
            self.positions = self.poses[:, :3, -1]
        self.load_data()

    def _cache(self, what):
        return os.path.join(self.result_path, "save_npy", "split_child_nerf2_3", "self_%s_%s.npy" % (what, self.split))

    def frames(self):
        """File numbers j+1 of this split (frame sparsity 20 %, :643-658)."""
        out = []
        for j in range(self.data_start, self.data_end):
            if self.split == 'train' and (j + 1 - 3 - self.data_start) % 5 != 0:
                out.append(j)
            elif self.split == 'val' and (j + 1 - 3) % 5 == 0:
                out.append(j)
        return out

    def load_data(self):
        self._build_or_load()
        if self.split == 'train':                                                 # :854-856
            self.all_rays, self.all_ranges = self.rays, self.ranges
        else:                                                                     # :865-870 (evenly spaced validation rays)
            n = self.rays.shape[0]
            sel = torch.linspace(1, n - 2, steps=self.cloud_size_val, dtype=torch.float32).floor().long()
            self._val_rays, self._val_ranges = self.rays[sel].float(), self.ranges[sel].float()

    def _build_or_load(self):
        if not self.re_loaddata:
            self.rays = torch.from_numpy(np.load(self._cache("rays")))
            self.ranges = torch.from_numpy(np.load(self._cache("ranges")))
            return
        K = self.sub_nerf_test_num
        bound, centre = np.zeros((K, 6)), np.zeros((K, 3))
        for i in range(K):                                                        # :600-626
            lo, hi = pcd.axis_aligned_bounds(pcd.read_pcd(os.path.join(self.subnerf_path, "%d.pcd" % (i + 1))))
            bound[i, :3], bound[i, 3:] = lo - 0.025, hi + 0.025
            centre[i] = (lo + hi) / 2.0
        pose_xy = self.poses[self.data_start + 1:self.data_end + 1, :2, -1]
        rays, counts = [], torch.zeros(K, dtype=torch.int64)
        for j in self.frames():
            pts = pcd.read_pcd(os.path.join(self.root_dir, "%d.pcd" % (j + 1)))
            world, dirs, dist = ops.frame_returns(pts, self.poses[j + 1], pose_xy, self.range_delete, 120.0,
                                                  self.over_height, self.over_low, self.interest_x, self.interest_y)
            origin = self.positions[j + 1].astype(np.float64)
            r, _ = ops.aabb_pack_train(606, origin, dirs, dist, world, centre, bound, bound, self.parent_box,
                                       self.surface_expand, 10)
            rays.append(r)
            if r.shape[0]:
                counts += torch.bincount(r[:, 9].long().cpu() - 1, minlength=K)
        self.rays = (torch.cat(rays, 0) if rays else torch.zeros((0, 15), device="cuda")).cpu()
        self.ranges = self.rays[:, 14].clone()
        self.sub_nerf_num_count = counts.numpy().astype(np.float64)
        if self.result_path:
            os.makedirs(os.path.dirname(self._cache("rays")), exist_ok=True)
            np.save(self._cache("rays"), self.rays.numpy())
            np.save(self._cache("ranges"), self.ranges.numpy())

    def __getitem__(self, index):
        if self.split == 'train':
            return {'rays': self.all_rays[index], 'ranges': self.all_ranges[index]}
        return {'rays': self._val_rays[index], 'ranges': self._val_ranges[index]}

    def __len__(self):
        return len(self.rays) if self.split == 'train' else self.cloud_size_val


class maicity_dataload(kitti_dataload):
    """nof/dataset/ipb2dmapping.py:200-463 (constructor signature, frame selection, outputs and cache paths as there): raw
    poses (frame file j+1 uses pose j), near-sensor box and `< 120 m` gates, closed parent-box test on the transformed
    points (the box is given by the nerf_* arguments, there is no parent cloud), float64 sensor positions and
    compute_far_bound0406.  Against the reference executed on the same files every column is identical
    (tests/test_gpu_dataset.py::test_maicity_dataload_end_to_end)."""

    def __init__(self, root_dir, split='train', data_start=0, data_end=36, cloud_size_val=2048, range_delete_x=2,
                 range_delete_y=1, range_delete_z=0.5, sub_nerf_test_num=3, surface_expand=0.1, nerf_length_min=-4.5,
                 nerf_length_max=25.5, nerf_width_min=-12, nerf_width_max=12, nerf_height_min=-2, nerf_height_max=0.5,
                 pose_path=None, subnerf_path=None, re_loaddata=0, result_path=None):
        torch.utils.data.Dataset.__init__(self)
        self.root_dir, self.split, self.cloud_size_val = root_dir, split, cloud_size_val
        self.result_path, self.re_loaddata = result_path, re_loaddata
        if re_loaddata:
            self.data_start, self.data_end = data_start, data_end
            self.range_delete = (range_delete_x, range_delete_y, range_delete_z)
            self.sub_nerf_test_num, self.subnerf_path, self.surface_expand = sub_nerf_test_num, subnerf_path, surface_expand
            self.parent_box = (nerf_length_min, nerf_length_max, nerf_width_min, nerf_width_max, nerf_height_min,
                               nerf_height_max)
            with open(pose_path, "r", encoding="utf-8") as f:
                rows = [r.strip() for r in f.readlines() if r.strip()]
            poses = np.array([np.append(np.array([float(i) for i in r.split(' ')]).reshape(3, 4), np.array([[0, 0, 0, 1]]),
                                        axis=0) for r in rows])
            self.positions = poses[:, :3, -1]                                     # float64 (:245-246)
            self.poses = torch.Tensor(poses).numpy()                              # float32 (:247)
        self.load_data()

    def frames(self):
        out = []
        for j in range(self.data_start, self.data_end):
            if self.split == 'train' and (j + 1 - 3 - self.data_start) % 5 != 0:
                out.append(j)
            elif self.split == 'val' and (j + 1 - 3 - self.data_start) % 5 == 0:
                out.append(j)
        return out

    def _build_or_load(self):
        if not self.re_loaddata:
            self.rays = torch.from_numpy(np.load(self._cache("rays")))
            self.ranges = torch.from_numpy(np.load(self._cache("ranges")))
            return
        K = self.sub_nerf_test_num
        bound, centre = np.zeros((K, 6)), np.zeros((K, 3))
        for i in range(K):                                                        # :256-280
            lo, hi = pcd.axis_aligned_bounds(pcd.read_pcd(os.path.join(self.subnerf_path, "%d.pcd" % (i + 1))))
            bound[i, :3], bound[i, 3:] = lo - 0.025, hi + 0.025
            centre[i] = (lo + hi) / 2.0
        strict_120 = float(np.nextafter(np.float32(120.0), np.float32(0.0)))      # dist < 120 (:327)
        rays, counts = [], torch.zeros(K, dtype=torch.int64)
        for j in self.frames():
            pts = pcd.read_pcd(os.path.join(self.root_dir, "%d.pcd" % (j + 1)))
            # transform with the float32 pose, directions / ranges from the float64 position (:331, :340)
            world, dirs, dist = ops.frame_returns(pts, self.poses[j], None, self.range_delete, strict_120, float("inf"),
                                                  -float("inf"), 0.0, 0.0, parent_box=self.parent_box,
                                                  position=self.positions[j])
            r, keep = ops.aabb_pack_train(406, self.positions[j], dirs, dist, world, centre, bound, bound, self.parent_box,
                                          self.surface_expand, 10)
            r[:, 0:3] = torch.from_numpy(self.poses[j][:3, -1]).to(r.device)      # rays_o = float32 pose translation (:410)
            rays.append(r)
            if r.shape[0]:
                counts += torch.bincount(r[:, 9].long().cpu() - 1, minlength=K)
        self.rays = (torch.cat(rays, 0) if rays else torch.zeros((0, 15), device="cuda")).cpu()
        self.ranges = self.rays[:, 14].clone()
        self.sub_nerf_num_count = counts.numpy().astype(np.float64)
        if self.result_path:
            os.makedirs(os.path.dirname(self._cache("rays")), exist_ok=True)
            np.save(self._cache("rays"), self.rays.numpy())
            np.save(self._cache("ranges"), self.ranges.numpy())
