"""TEST INFRASTRUCTURE ONLY -- import the *reference itself* (biter0088/pc-nerf) on a CPU box.

Only usable where /root/reference exists (this container; NOT the GPU box).  It is used by
`oracle/make_golden.py` to generate the committed fixtures under `tests/golden/` and by the optional
`tests/test_oracle_vs_reference.py` cross-check.  Nothing in the product package imports this file.

Shims:
  1. stub modules for open3d / pcl / matplotlib / pytorch_lightning / tqdm so that
     `eval_kitti_render`, `nof.dataset.ipb2dmapping` and `train_kitti` import;
  2. the hard-coded `u.to("cuda:0")` at nof/render.py:397 is neutralised by wrapping
     `torch.Tensor.to` while a reference call is in flight (CPU-only host).
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("PCNERF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nof"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "open3d" not in sys.modules:
        _stub("open3d")
    if "pcl" not in sys.modules:
        _stub("pcl")
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot", ion=lambda: None, ioff=lambda: None, show=lambda: None)
        mpl.pyplot = plt
    if "pytorch_lightning" not in sys.modules:
        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, hparams):
                self.hparams = hparams

            def log(self, *a, **k):
                pass

        pl = _stub("pytorch_lightning", LightningModule=LightningModule, Trainer=object,
                   seed_everything=lambda *a, **k: None)
        cb = _stub("pytorch_lightning.callbacks", ModelCheckpoint=object)
        lg = _stub("pytorch_lightning.loggers", TensorBoardLogger=object)
        pl.callbacks, pl.loggers = cb, lg
    try:
        import tqdm  # noqa: F401
    except Exception:  # pragma: no cover
        _stub("tqdm", tqdm=lambda x, **k: x)


def import_reference():
    """Returns a namespace with the reference's modules.  The reference package is called `nof`;
    the product mirror lives at `pcnerf_b200.nof`, so there is no name clash."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import nof.render as render
    import nof.networks as networks
    import nof.criteria as criteria
    import nof.dataset.ipb2dmapping as ipb
    import eval_kitti_render as evalmod
    import train_kitti as trainmod
    return types.SimpleNamespace(render=render, networks=networks, criteria=criteria, ipb=ipb,
                                 evalmod=evalmod, trainmod=trainmod)


@contextlib.contextmanager
def cuda0_to_cpu():
    """nof/render.py:397 does `u=u.to("cuda:0")`; on a CPU-only host map that device to cpu."""
    orig = torch.Tensor.to

    def patched(self, *args, **kwargs):
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
        return orig(self, *args, **kwargs)

    torch.Tensor.to = patched
    try:
        yield
    finally:
        torch.Tensor.to = orig
