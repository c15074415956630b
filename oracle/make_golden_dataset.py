"""TEST INFRASTRUCTURE ONLY -- golden fixture for the KITTI dataset-build path.

Executes the UNMODIFIED reference class `nof.dataset.ipb2dmapping.kitti_dataload` (ipb2dmapping.py:516-836) on a real
shipped frames (data/kitti/00/pcd_remove_dynamic/1151.pcd and 1152.pcd, every 40th point) with the shipped poses.txt and synthetic child
point clouds cut out of that frame, through functional stand-ins for the two libraries that are absent here
(open3d.io.read_point_cloud -> axis-aligned bounds; pcl.load -> (N,3) float32 array: both only read binary `x y z` PCD
files, implemented below with numpy.frombuffer, independently of the product's pcnerf_b200/pcd.py).

Writes tests/golden/kitti_dataset.npz: the input files' contents (raw frame points, child clouds, parent cloud, the
pose lines that are used) and the reference's outputs (self.rays (N,15), self.ranges, sub_nerf_num_count).

    python oracle/make_golden_dataset.py
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "kitti_dataset.npz")
DATA_START, DATA_END, FRAMES = 1150, 1152, (1151, 1152)      # both are 'train' frames of the 20 % sparsity rule (:645)
N_CHILD = 40
ARGS = dict(range_delete_x=3, range_delete_y=2, range_delete_z=1.25, surface_expand=0.05, over_height=0.168,
            over_low=-2.0, interest_x=20, interest_y=20)


def read_xyz_pcd(path):
    raw = open(path, "rb").read()
    head_end = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    head = raw[:head_end].decode("ascii").splitlines()
    fields = [l for l in head if l.startswith("FIELDS")][0].split()[1:]
    assert fields[:3] == ["x", "y", "z"] and len(fields) == 3, fields
    n = int([l for l in head if l.startswith("POINTS")][0].split()[1])
    return np.frombuffer(raw, dtype="<f4", count=3 * n, offset=head_end).reshape(n, 3).copy()


def write_xyz_pcd(path, xyz):
    xyz = np.ascontiguousarray(xyz, dtype="<f4")
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
            "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz)))
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(xyz.tobytes())


class _Box:
    def __init__(self, pts):
        self.lo, self.hi = pts.min(0), pts.max(0)

    def get_min_bound(self):
        return self.lo

    def get_max_bound(self):
        return self.hi


class _O3dCloud:
    def __init__(self, path):
        self.points = read_xyz_pcd(path).astype(np.float64)       # open3d stores points as float64

    def get_axis_aligned_bounding_box(self):
        return _Box(self.points)


class _PclCloud:
    def __init__(self, path=None):
        self.arr = read_xyz_pcd(path) if path else np.zeros((0, 3), np.float32)
        self.size = self.arr.shape[0]

    def to_array(self):
        return self.arr


def install_functional_stubs():
    ref_shim.install_stubs()
    o3d = sys.modules["open3d"]
    o3d.io = types.SimpleNamespace(read_point_cloud=lambda p: _O3dCloud(p))
    pcl = sys.modules["pcl"]
    pcl.PointCloud = _PclCloud
    pcl.load = lambda p: _PclCloud(p)


def main():
    install_functional_stubs()
    ref = ref_shim.import_reference()
    src_dir = os.path.join(ref_shim.REF_ROOT, "data", "kitti", "00")
    frames = {f: read_xyz_pcd(os.path.join(src_dir, "pcd_remove_dynamic", "%d.pcd" % f))[::40] for f in FRAMES}
    pose_lines = open(os.path.join(src_dir, "poses.txt")).read().splitlines()[:DATA_END + 2]
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as tmp:
        root = os.path.join(tmp, "frames")
        sub = os.path.join(tmp, "children")
        res = os.path.join(tmp, "result")
        os.makedirs(root)
        os.makedirs(sub)
        os.makedirs(os.path.join(res, "save_npy", "split_child_nerf2_3"))
        for f, pts in frames.items():
            write_xyz_pcd(os.path.join(root, "%d.pcd" % f), pts)
        pose_path = os.path.join(tmp, "poses.txt")
        open(pose_path, "w").write("\n".join(pose_lines) + "\n")
        # first pass with a dummy child set to learn where the reference places the returns of this frame
        # (world frame of pose DATA_START+1): children are 1 m cells cut out of those points
        T = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                      [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                      [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01], [0, 0, 0, 1]])
        P = [np.vstack([np.array([float(v) for v in l.split(" ")]).reshape(3, 4), [[0, 0, 0, 1]]]) @ T for l in pose_lines]
        world = []
        for f, frame in frames.items():
            rel = np.linalg.inv(P[DATA_START + 1]) @ P[f]
            keep = ((np.abs(frame[:, 0]) >= 3) | (np.abs(frame[:, 1]) >= 2) | (np.abs(frame[:, 2]) >= 1.25)) & \
                   (frame[:, 2] <= 0.168) & (frame[:, 2] >= -2.0)
            world.append((rel @ np.vstack([frame[keep].T.astype(np.float64), np.ones((1, int(keep.sum())))])).T[:, :3])
        world = np.concatenate(world)
        children = []
        for c in world[rng.choice(len(world), N_CHILD, replace=False)]:
            cell = world[np.all(np.abs(world - c) <= 0.5, axis=1)]
            children.append(cell.astype(np.float32))
        for i, ch in enumerate(children):
            write_xyz_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), ch)
        parent = np.concatenate([world.min(0, keepdims=True) - 1.0, world.max(0, keepdims=True) + 1.0]).astype(np.float32)
        parent_path = os.path.join(tmp, "source.pcd")
        write_xyz_pcd(parent_path, parent)
        ds = ref.ipb.kitti_dataload(root, split="train", data_start=DATA_START, data_end=DATA_END, cloud_size_val=64,
                                    sub_nerf_test_num=N_CHILD, pose_path=pose_path, subnerf_path=sub,
                                    parentnerf_path=parent_path, re_loaddata=1, result_path=res, **ARGS)
        rays = ds.rays.numpy()
        ranges = ds.ranges.numpy()
        cached = np.load(os.path.join(res, "save_npy", "split_child_nerf2_3", "self_rays_train.npy"))
        assert np.array_equal(cached, rays)
        out = {"frame_ids": np.array(FRAMES), "parent": parent, "pose_lines": np.array(pose_lines), "rays": rays, "ranges": ranges,
               "sub_nerf_num_count": ds.sub_nerf_num_count, "n_child": np.int64(N_CHILD),
               "data_start": np.int64(DATA_START), "data_end": np.int64(DATA_END)}
        for f, pts in frames.items():
            out["frame_%d" % f] = pts
        for i, ch in enumerate(children):
            out["child_%d" % (i + 1)] = ch
        for k, v in ARGS.items():
            out["arg_" + k] = np.float64(v)
        np.savez_compressed(OUT, **out)
        print("wrote", OUT, "rays", rays.shape, "frame points", [v.shape for v in frames.values()], "bytes", os.path.getsize(OUT))


OUT_M = os.path.join(os.path.dirname(HERE), "tests", "golden", "maicity_dataset.npz")
M_START, M_END, M_FRAMES = 0, 2, (1, 2)                                   # both 'train' frames of the sparsity rule (:306)
M_ARGS = dict(range_delete_x=2, range_delete_y=1, range_delete_z=0.5, surface_expand=0.05, nerf_length_min=-12,
              nerf_length_max=61, nerf_width_min=-12, nerf_width_max=12, nerf_height_min=-2, nerf_height_max=0.5)


def main_maicity():
    """Same for `maicity_dataload` (ipb2dmapping.py:200-463): shipped frames data/maicity/00/pcd/{1,2}.pcd (every 25th
    point), shipped poses, 40 synthetic child clouds; compute_far_bound0406 raises IndexError when a ray misses its box,
    so the children are 1 m cells around returns (every ray crosses the box that contains its return)."""
    install_functional_stubs()
    ref = ref_shim.import_reference()
    src_dir = os.path.join(ref_shim.REF_ROOT, "data", "maicity", "00")
    frames = {f: read_xyz_pcd(os.path.join(src_dir, "pcd", "%d.pcd" % f))[::25] for f in M_FRAMES}
    pose_lines = open(os.path.join(src_dir, "poses.txt")).read().splitlines()[:M_END + 2]
    rng = np.random.default_rng(11)
    with tempfile.TemporaryDirectory() as tmp:
        root, sub, res = os.path.join(tmp, "frames"), os.path.join(tmp, "children"), os.path.join(tmp, "result")
        os.makedirs(root)
        os.makedirs(sub)
        os.makedirs(os.path.join(res, "save_npy", "split_child_nerf2_3"))
        for f, pts in frames.items():
            write_xyz_pcd(os.path.join(root, "%d.pcd" % f), pts)
        pose_path = os.path.join(tmp, "poses.txt")
        open(pose_path, "w").write("\n".join(pose_lines) + "\n")
        P = [np.vstack([np.array([float(v) for v in l.split(" ")]).reshape(3, 4), [[0, 0, 0, 1]]]) for l in pose_lines]
        world = []
        for f, frame in frames.items():
            keep = ((np.abs(frame[:, 0]) >= 2) | (np.abs(frame[:, 1]) >= 1) | (np.abs(frame[:, 2]) >= 0.5))
            w = (P[f - 1] @ np.vstack([frame[keep].T.astype(np.float64), np.ones((1, int(keep.sum())))])).T[:, :3]
            world.append(w[(w[:, 0] >= -12) & (w[:, 0] <= 61) & (np.abs(w[:, 1]) <= 12) & (w[:, 2] >= -2) & (w[:, 2] <= 0.5)])
        world = np.concatenate(world)
        children = []
        for c in world[rng.choice(len(world), N_CHILD, replace=False)]:
            children.append(world[np.all(np.abs(world - c) <= 0.5, axis=1)].astype(np.float32))
        for i, ch in enumerate(children):
            write_xyz_pcd(os.path.join(sub, "%d.pcd" % (i + 1)), ch)
        ds = ref.ipb.maicity_dataload(root, split="train", data_start=M_START, data_end=M_END, cloud_size_val=64,
                                      sub_nerf_test_num=N_CHILD, pose_path=pose_path, subnerf_path=sub, re_loaddata=1,
                                      result_path=res, **M_ARGS)
        rays, ranges = ds.rays.numpy(), ds.ranges.numpy()
        out = {"frame_ids": np.array(M_FRAMES), "pose_lines": np.array(pose_lines), "rays": rays, "ranges": ranges,
               "sub_nerf_num_count": ds.sub_nerf_num_count, "n_child": np.int64(N_CHILD), "data_start": np.int64(M_START),
               "data_end": np.int64(M_END)}
        for f, pts in frames.items():
            out["frame_%d" % f] = pts
        for i, ch in enumerate(children):
            out["child_%d" % (i + 1)] = ch
        for k, v in M_ARGS.items():
            out["arg_" + k] = np.float64(v)
        np.savez_compressed(OUT_M, **out)
        print("wrote", OUT_M, "rays", rays.shape, "frame points", [v.shape for v in frames.values()], "bytes", os.path.getsize(OUT_M))


if __name__ == "__main__":
    if "--maicity" in sys.argv:
        main_maicity()
    else:
        main()
