"""pcnerf_b200 -- B200-native (sm_100a) implementation of PC-NeRF's LiDAR ray-rendering hot path behind the
reference's own Python API (`pcnerf_b200.nof.render`, `.nof.networks`, `.nof.criteria`, `.nof.dataset.ipb2dmapping`,
`.train_kitti`, `.eval_kitti_render`).  Kernels live in libpcnerf_b200.so (C ABI: include/pcnerf_b200.h)."""
__version__ = "0.1.0"


def install_as_nof():
    """Make `import nof` resolve to this package's mirror (for unmodified reference entry points)."""
    import sys
    from . import nof as _nof
    sys.modules.setdefault("nof", _nof)
    for sub in ("render", "networks", "criteria", "criteria.metrics", "dataset", "dataset.ipb2dmapping"):
        mod = __import__("pcnerf_b200.nof." + sub, fromlist=["x"])
        sys.modules.setdefault("nof." + sub, mod)
    return _nof
