"""CPU: the C-ABI library builds, loads and exports every symbol declared in include/pcnerf_b200.h; the ctypes binding
covers the same set; compute entry points refuse to run without a CUDA device (no CPU fallback)."""
import ctypes

import pytest
import torch

from pcnerf_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    build.build()
    l = ctypes.CDLL(_lib.LIB_PATH)
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(l, s), "missing export: " + s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert _lib.lib().pcnerf_version() >= 100


def test_error_channel_and_no_cpu_fallback():
    l = _lib.lib()
    # argument validation happens before any CUDA call: null pointers -> PCNERF_ERR_ARG + message
    rc = l.pcnerf_embed(None, 4, None, 63, None)
    assert rc == _lib.ERR_ARG and l.pcnerf_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    from pcnerf_b200 import ops
    with pytest.raises(RuntimeError):
        ops.embed(torch.zeros(3, 3), 63)           # CPU tensor: rejected, never computed on the host
    from pcnerf_b200.nof.networks import NOF_coarse
    with pytest.raises(RuntimeError):
        NOF_coarse()(torch.zeros(8, 63))


def test_product_never_imports_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "pcnerf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|pcnerf_oracle|ref_shim)", txt, re.M), f


def test_torch_ops_registered_for_cuda_only():
    """the stages are dispatcher ops `torch.ops.pcnerf.*` with a CUDA kernel only -- a CPU tensor finds no
    kernel (the dispatcher raises), nothing is computed on the host."""
    import pcnerf_b200.torch_ops as t
    for name in t.OPS:
        assert hasattr(torch.ops.pcnerf, name), name
    with pytest.raises(NotImplementedError):
        torch.ops.pcnerf.points(torch.zeros(4, 13), torch.zeros(4))
    with pytest.raises(NotImplementedError):
        torch.ops.pcnerf.search_select(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8), torch.zeros(4))
