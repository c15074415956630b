"""PCD v0.7 point-cloud files and KITTI pose files, without open3d / pcl.

The reference reads its LiDAR frames with `pcl.load(path).to_array()` (ipb2dmapping.py:662-665), its child / parent
clouds with `o3d.io.read_point_cloud(path).get_axis_aligned_bounding_box()` (ipb2dmapping.py:553-556, :603-607;
eval_kitti_render.py:319-323) and writes rendered clouds with `o3d.io.write_point_cloud` (eval_kitti_render.py:1034-1044).
Every file it ships or produces is `FIELDS x y z`, `SIZE 4 4 4`, `TYPE F F F`, `DATA binary` (header of
data/kitti/00/pcd_remove_dynamic/1151.pcd).  This module covers that format plus the obvious neighbours (extra fields,
ascii data, float64 coordinates) so that the entry points run on the reference's data layout unchanged.  Host-side IO only:
nothing here is on the hot path.
"""
import os

import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8",
       ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


def _header(raw):
    """Parse the text header; returns (dict, byte offset of the data)."""
    hdr, pos = {}, 0
    while True:
        end = raw.find(b"\n", pos)
        if end < 0:
            raise ValueError("PCD: header without DATA line")
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        hdr[key.upper()] = val.split()
        if key.upper() == "DATA":
            return hdr, pos


def read_pcd(path, dtype=np.float32):
    """(N,3) array of the x, y, z fields of a PCD file (ascii or binary; any extra fields are skipped)."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr, off = _header(raw)
    fields = hdr["FIELDS"]
    for need in ("x", "y", "z"):
        if need not in fields:
            raise ValueError("PCD %s: no %r field (FIELDS %s)" % (path, need, " ".join(fields)))
    sizes = [int(v) for v in hdr.get("SIZE", ["4"] * len(fields))]
    types = hdr.get("TYPE", ["F"] * len(fields))
    counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr.get("HEIGHT", ["1"])[0])
    kind = hdr["DATA"][0].lower()
    if kind == "binary":
        rec = np.dtype({"names": fields, "formats": [(_NP[(t, s)], c) if c > 1 else _NP[(t, s)]
                                                     for t, s, c in zip(types, sizes, counts)]})
        if len(raw) - off < n * rec.itemsize:
            raise ValueError("PCD %s: %d points declared, data holds %d" % (path, n, (len(raw) - off) // rec.itemsize))
        data = np.frombuffer(raw, dtype=rec, count=n, offset=off)
        return np.stack([data["x"], data["y"], data["z"]], axis=1).astype(dtype, copy=False)
    if kind == "ascii":
        cols = np.cumsum([0] + counts)
        ix = [int(cols[fields.index(a)]) for a in ("x", "y", "z")]
        txt = np.loadtxt(raw[off:].decode("ascii").splitlines(), ndmin=2, dtype=np.float64) if n else np.zeros((0, cols[-1]))
        return txt[:n][:, ix].astype(dtype)
    raise NotImplementedError("PCD %s: DATA %s is not supported (the reference's files are binary)" % (path, kind))


def write_pcd(path, xyz):
    """Write (N,3) points the way o3d.io.write_point_cloud does by default: binary float32 x y z, PCD v0.7."""
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype="<f4").reshape(-1, 3))
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
            "COUNT 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(xyz), len(xyz)))
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(xyz.tobytes())


def axis_aligned_bounds(xyz):
    """(min (3,), max (3,)) in float64 -- `get_axis_aligned_bounding_box().get_min_bound() / get_max_bound()`
    (open3d keeps points as float64)."""
    p = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    if p.shape[0] == 0:
        raise ValueError("axis_aligned_bounds: empty cloud")
    return p.min(0), p.max(0)


T_VELO2CAM = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                       [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                       [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01],
                       [0, 0, 0, 1]])


def read_kitti_poses(pose_path_or_lines, data_start):
    """ipb2dmapping.py:559-591: one `3x4` row-major pose per line, times T_velo2cam, re-expressed in the frame of pose
    data_start+1 with the product taken in float32 (the reference goes through torch.Tensor).  Returns (n,4,4) float32."""
    import torch
    if isinstance(pose_path_or_lines, (str, os.PathLike)):
        with open(pose_path_or_lines, "r", encoding="utf-8") as f:
            lines = [r.strip() for r in f.readlines()]
    else:
        lines = [str(r).strip() for r in pose_path_or_lines]
    poses = []
    for row in lines:
        if not row:
            continue
        P = np.append(np.array([float(i) for i in row.split(" ")]).reshape(3, 4), np.array([[0, 0, 0, 1]]), axis=0)
        poses.append(np.matmul(P, T_VELO2CAM))
    poses = np.array(poses)
    t_inv = torch.from_numpy(np.linalg.inv(poses[data_start + 1])).float()
    return (t_inv @ torch.Tensor(poses)).numpy()
