"""CPU: PCD / pose IO (pcnerf_b200.pcd) and the oracle's restatement of the KITTI dataset build against the
fixture produced by executing the reference class itself (oracle/make_golden_dataset.py -> tests/golden/kitti_dataset.npz:
two shipped KITTI frames, shipped poses, 40 synthetic child clouds)."""
import os

import numpy as np

import pcnerf_oracle as orc
from conftest import golden

BOUND_COLS = [7, 10, 11, 13]          # evaluated partly in float32 by the reference (float32 tensor origin): 1 ulp


def _inputs(g):
    frames = {int(f): g["frame_%d" % f] for f in g["frame_ids"]}
    children = [g["child_%d" % (i + 1)] for i in range(int(g["n_child"]))]
    kw = {k[4:]: float(g[k]) for k in g.files if k.startswith("arg_")}
    return frames, children, kw


def check_rays(rays, ref):
    assert rays.shape == ref.shape
    exact = [c for c in range(15) if c not in BOUND_COLS]
    assert np.array_equal(rays[:, exact], ref[:, exact])                       # origins, directions, child index, ranges
    np.testing.assert_allclose(rays[:, BOUND_COLS], ref[:, BOUND_COLS], rtol=2.4e-7, atol=0)


def test_oracle_dataset_build_matches_reference_run():
    g = golden("kitti_dataset")
    frames, children, kw = _inputs(g)
    rays, ranges = orc.kitti_build_rays(frames, list(g["pose_lines"]), children, g["parent"], int(g["data_start"]),
                                        int(g["data_end"]), **kw)
    check_rays(rays, g["rays"])
    assert np.array_equal(ranges, g["ranges"])
    counts = np.bincount(rays[:, 9].astype(np.int64) - 1, minlength=int(g["n_child"]))
    assert np.array_equal(counts, g["sub_nerf_num_count"].astype(np.int64))


def test_pcd_roundtrip_and_variants(tmp_path):
    from pcnerf_b200 import pcd
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(1000, 3)).astype(np.float32) * 30
    p = str(tmp_path / "a" / "cloud.pcd")
    pcd.write_pcd(p, xyz)
    raw = open(p, "rb").read()
    assert raw.startswith(b"# .PCD v0.7") and b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n" in raw and b"DATA binary\n" in raw
    assert np.array_equal(pcd.read_pcd(p), xyz)
    lo, hi = pcd.axis_aligned_bounds(xyz)
    assert lo.dtype == np.float64 and np.array_equal(lo, xyz.min(0).astype(np.float64)) and np.array_equal(hi, xyz.max(0))
    # extra fields and float64 coordinates, binary
    rec = np.zeros(5, dtype=[("x", "<f8"), ("intensity", "<f4"), ("y", "<f8"), ("z", "<f8"), ("ring", "<u2")])
    rec["x"], rec["y"], rec["z"] = [1, 2, 3, 4, 5], [6, 7, 8, 9, 10], [-1, -2, -3, -4, -5]
    q = str(tmp_path / "b.pcd")
    with open(q, "wb") as f:
        f.write(b"VERSION .7\nFIELDS x intensity y z ring\nSIZE 8 4 8 8 2\nTYPE F F F F U\nCOUNT 1 1 1 1 1\nWIDTH 5\nHEIGHT 1\n"
                b"POINTS 5\nDATA binary\n" + rec.tobytes())
    got = pcd.read_pcd(q, dtype=np.float64)
    assert np.array_equal(got, np.stack([rec["x"], rec["y"], rec["z"]], 1))
    # ascii
    a = str(tmp_path / "c.pcd")
    with open(a, "w") as f:
        f.write("# comment\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n"
                "1.5 2.5 3.5\n-4 5 6e-1\n")
    assert np.allclose(pcd.read_pcd(a), [[1.5, 2.5, 3.5], [-4, 5, 0.6]])
    # empty cloud, truncated data
    pcd.write_pcd(str(tmp_path / "e.pcd"), np.zeros((0, 3)))
    assert pcd.read_pcd(str(tmp_path / "e.pcd")).shape == (0, 3)
    with open(str(tmp_path / "t.pcd"), "wb") as f:
        f.write(raw[:len(raw) - 40])
    try:
        pcd.read_pcd(str(tmp_path / "t.pcd"))
        raise AssertionError("truncated file accepted")
    except ValueError:
        pass


def test_kitti_poses_match_oracle():
    from pcnerf_b200 import pcd
    g = golden("kitti_dataset")
    ref = orc.kitti_poses(list(g["pose_lines"]), int(g["data_start"])).numpy()
    got = pcd.read_kitti_poses(list(g["pose_lines"]), int(g["data_start"]))
    assert got.dtype == np.float32 and np.array_equal(got, ref)
    assert np.allclose(got[int(g["data_start"]) + 1], np.eye(4), atol=1e-4)       # the run's first pose is the origin (float32 product)


def test_oracle_maicity_build_matches_reference_run():
    """maicity_dataload restated (float64 sensor positions there: every column is exact)."""
    g = golden("maicity_dataset")
    frames, children, kw = _inputs(g)
    rays, ranges = orc.maicity_build_rays(frames, list(g["pose_lines"]), children, int(g["data_start"]), int(g["data_end"]), **kw)
    assert np.array_equal(rays, g["rays"]) and np.array_equal(ranges, g["ranges"])
