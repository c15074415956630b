"""CPU: host-side logic of the boundary that needs no device -- the group-aligned batch driver of the depth-inference
entry point (eval_kitti_render.py:979-1005) against the oracle's restatement, and the mirror's public surface."""
import inspect

import numpy as np

import pcnerf_oracle as orc
from pcnerf_b200 import eval_kitti_render as ev
from pcnerf_b200 import synth


def test_eval_batches_never_split_a_group():
    rows, other, _ = synth.synth_infer_rows(5, 400)
    for bs in (64, 100, 257, 5000):
        got = ev.eval_batches(rows[:, -1], bs)
        assert got == orc.eval_batches(rows, bs)
        for a, b in got:
            assert rows[a, -1] >= 0                       # every batch starts at a group head
        assert got[0][0] == 0 and all(got[i][1] == got[i + 1][0] for i in range(len(got) - 1))


def test_mirror_signatures_match_reference_names():
    """Parameter names / defaults of the reference's public entry points (nof/render.py:13-15, :38-40, :166-167, :229-231,
    :371, :416-418, :485-486, :538-539, :614-616) as recorded in SURVEY.md section 8b."""
    from pcnerf_b200.nof import render
    want = {
        "render_rays_train": ["model", "model_fine", "embedding_xy", "rays", "sub_nerf_test_num", "N_samples", "N_importance",
                              "use_disp", "perturb", "noise_std", "chunk", "isval", "issegmentated", "childnerf_ratio",
                              "use_child_nerf_divide", "use_child_nerf_loss"],
        "render_rays_val": ["model", "model_fine", "embedding_xy", "rays", "sub_nerf_test_num", "N_samples", "N_importance",
                            "use_disp", "perturb", "noise_std", "chunk", "isval"],
        "render_rays": ["model", "model_fine", "embedding_xy", "rays", "N_samples", "N_importance", "use_disp", "perturb",
                        "noise_std", "chunk", "isval"],
        "render_rays_view_0525_2_2": ["model", "model_fine", "embedding_xy", "rays", "other_interest_sub_nerf_number",
                                      "N_samples", "N_importance", "use_disp", "perturb", "noise_std", "chunk", "isval",
                                      "depth_inference_method"],
        "sample_pdf": ["bins", "weights", "N_samples", "det", "pytest"],
    }
    for name, params in want.items():
        sig = inspect.signature(getattr(render, name))
        positional = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert positional == params, name
    d = {k: v.default for k, v in inspect.signature(render.render_rays_train).parameters.items()}
    assert (d["N_samples"], d["N_importance"], d["chunk"], d["noise_std"], d["childnerf_ratio"]) == (64, 128, 1024 * 3, 1, 0.5)
    assert render.__all__ == ["render_rays"]


def test_scene_routing_and_block_sharding():
    """Multi-parent scenes (BASELINE configs[4]): returns go to the first parent box that contains them, whole blocks are
    sharded contiguously over the ranks, every block is owned exactly once."""
    import numpy as np
    from pcnerf_b200 import scene
    blocks = [scene.ParentBlock([10 * i, 0, -2], [10 * i + 10, 10, 1], np.zeros((0, 6))) for i in range(5)]
    pts = np.array([[5.0, 5, 0], [10.0, 5, 0], [49.9, 9.9, 0.9], [50.1, 5, 0], [25, -1, 0], [25, 5, 0]])
    assert scene.route_points(pts, blocks).tolist() == [0, 0, 4, -1, -1, 2]        # a shared face goes to the first block
    owned = [scene.owned_blocks(5, 3, r) for r in range(3)]
    assert owned == [[0, 1], [2, 3], [4]]
    assert sorted(sum([scene.owned_blocks(64, 8, r) for r in range(8)], [])) == list(range(64))
    assert all(len(scene.owned_blocks(64, 8, r)) == 8 for r in range(8))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) needs no GPU: one JSON line with the
    contract's keys; the UNMODIFIED reference (imported from /root/reference) where its tree exists, the oracle port
    elsewhere (the GPU box), timed on a bounded sample of the bench workload."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-rays", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    want = "reference" if os.path.isdir(os.path.join(os.environ.get("PCNERF_REFERENCE_ROOT", "/root/reference"), "nof")) else "port"
    assert line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert k in line


def test_hoisted_reciprocal_division_matches_ieee_division():
    """K4's `div_rc` (csrc/composite.cu): q = a * rc, two residual corrections q += fma(-d, q, a) * rc with
    rc = RN(1 / d) hoisted out of the per-sample loop.  Exact emulation (rational arithmetic, one rounding per fp32
    operation / FMA): the result equals the correctly rounded quotient RN(a / d) for operands in the range the weights
    and normalisers of the compositing stage live in."""
    from fractions import Fraction
    import numpy as np

    def rn(x):                                             # Fraction -> nearest float32, ties to even
        c = np.float32(float(x))
        best, bd = c, abs(Fraction(float(c)) - x)
        for n in (np.nextafter(c, np.float32(-np.inf)), np.nextafter(c, np.float32(np.inf))):
            dlt = abs(Fraction(float(n)) - x)
            if dlt < bd or (dlt == bd and (n.view(np.uint32) & 1) == 0 and (best.view(np.uint32) & 1) == 1):
                best, bd = n, dlt
        return best

    rng = np.random.default_rng(7)
    a = (rng.random(4000) * 10.0 ** rng.uniform(-12, 0, 4000)).astype(np.float32)      # weights: 1e-12 .. 1
    d = (rng.random(4000) * 0.999 + 0.001).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 2, 4000).astype(np.float32)
    bad = 0
    for ai, di in zip(a, d):
        if ai == 0 or di == 0:
            continue
        A, D = Fraction(float(ai)), Fraction(float(di))
        rc = Fraction(float(rn(1 / D)))
        q = Fraction(float(rn(A * rc)))
        for _ in range(2):
            r = Fraction(float(rn(A - D * q)))             # fma(-d, q, a): exact product and sum, one rounding
            q = Fraction(float(rn(q + r * rc)))            # fma(r, rc, q)
        bad += (np.float32(float(q)) != rn(A / D))
    assert bad == 0


def test_closed_form_rows_are_lazy_only_where_the_engine_can_take_them():
    """nof/render.py::_lazy decides when K2 / K2' write depths only and the closed-form engine re-derives the encodings
    (csrc/affine_rays.cu): training passes of a precision-2 model, eval passes without autograd -- never for the layered
    engines, and never for an eval pass that autograd records (those get the encoding tensor).  Host logic only: no kernel
    runs; a CPU tensor must be refused by the row container like by every other op (no CPU fallback)."""
    import pytest
    import torch
    from pcnerf_b200 import ops
    from pcnerf_b200.nof import render
    from pcnerf_b200.nof.networks import NOF_coarse
    m = NOF_coarse()
    for prec, train, grad, want in (("affine", True, True, True), ("affine", True, False, True), ("affine", False, False, True),
                                    ("affine", False, True, False), ("tc", True, True, False), ("tc", False, False, False),
                                    ("fp32", True, True, False), ("fp32", False, False, False)):
        m.precision = prec
        m.train(train)
        with torch.set_grad_enabled(grad):
            assert render._lazy(m) is want, (prec, train, grad)
    with pytest.raises((ValueError, TypeError, RuntimeError)):
        ops.LazyEnc(torch.zeros(4, 15), torch.zeros(4, 8))
