"""GPU: depth inference evaluated once per PHYSICAL ray (nof.render._view_grouped, eval_kitti_render.render_frame) against
the reference's once-per-candidate-row evaluation (nof/render.py:614-699, eval_kitti_render.py:979-1030).  All rows of a
candidate group share origin, direction and the parent segment, so samples, occupancies and weights are identical within a
group: the outputs must be BIT-identical, for every MLP engine."""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import dev, make_nets

pytestmark = pytest.mark.gpu
KEYS = ("depth", "depth_fine", "weights", "z_vals", "opacity", "opacity_fine", "points_inference", "points_inference_fine",
        "rays_effective_flag", "rays_effective_flag_fine")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _both(fn):
    from pcnerf_b200.nof import render
    out = {}
    old = render.GROUP_RAYS
    try:
        for flag in (False, True):
            render.GROUP_RAYS = flag
            out[flag] = fn()
    finally:
        render.GROUP_RAYS = old
    return out[False], out[True]


@pytest.mark.parametrize("precision", ["fp32", "tc", "affine"])
@pytest.mark.parametrize("method", [2, 1])
def test_view_grouped_is_bit_identical_to_per_row(precision, method):
    from pcnerf_b200.nof import render
    g = golden("view_m%d" % method)
    rays, other = _t(g["rays"]), _t(g["other"])
    mc, mf, emb = make_nets(42, 43, False, precision)

    def run():
        with torch.no_grad():
            return render.render_rays_view_0525_2_2(mc, mf, emb, rays, other, N_samples=int(g["S"]), N_importance=int(g["Ni"]),
                                                    perturb=0, noise_std=0, chunk=int(g["chunk"]),
                                                    depth_inference_method=method)

    a, b = _both(run)
    for k in KEYS:
        if k.startswith("opacity"):             # a sum over all rows: fp64 atomics, identical after the fp32 rounding
            np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), rtol=1e-6)
        else:
            assert torch.equal(a[k], b[k]), k


def test_grouped_plan_structure_and_fallback():
    from pcnerf_b200 import ops, synth
    from pcnerf_b200.nof import render
    rows, other, _ = synth.synth_infer_rows(5, 300)
    plan = ops.GroupPlan(_t(rows), _t(other))
    heads = np.nonzero(rows[:, 12] >= 0)[0]
    assert plan.G == 300 and plan.uniform and np.array_equal(plan.head_rows.cpu().numpy(), heads)
    assert np.array_equal(plan.row_ray.cpu().numpy(), np.cumsum(rows[:, 12] >= 0) - 1)
    # a follower that does NOT carry its head's ray: the rows are evaluated one by one, like the reference would
    bad = rows.copy()
    f = int(np.nonzero(rows[:, 12] < 0)[0][3])
    bad[f, 3:6] = bad[f, [4, 5, 3]]
    plan_b = ops.GroupPlan(_t(bad), _t(other))
    assert not plan_b.uniform
    mc, mf, emb = make_nets(42, 43, False, "fp32")

    def run():
        with torch.no_grad():
            return render.render_rays_view_0525_2_2(mc, mf, emb, _t(bad), _t(other), N_samples=32, N_importance=32, perturb=0,
                                                    noise_std=0, chunk=8192, depth_inference_method=2)

    a, b = _both(run)
    for k in ("depth_fine", "rays_effective_flag_fine", "weights"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("n_phys,batch,cut", [(700, 256, 0), (700, 8, 1), (4096, 18432, 0)])
def test_render_frame_grouped_is_bit_identical_to_the_batch_loop(precision, n_phys, batch, cut):
    """Whole frame at once vs the reference's group-aligned batch loop; `cut` trims the rows so that a batch ends exactly one
    row before the end -- the reference loop then never renders that last row (eval_kitti_render.py:984-985; needs a
    candidate group longer than half a batch, so only reachable with tiny batches)."""
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import synth
    rows, other, _ = synth.synth_infer_rows(91, n_phys)
    if cut:
        heads = np.nonzero(rows[:, 12] >= 0)[0]
        single = int(np.nonzero((rows[:, 12] == 0))[0][0])
        for h in heads[5:400]:
            r2 = np.concatenate([rows[:h], rows[single:single + 1]], 0)
            if ev.eval_batches(r2[:, 12], batch)[-1][1] == r2.shape[0] - 1:
                rows, other = r2, np.concatenate([other[:h], other[single:single + 1]], 0)
                break
        else:
            raise AssertionError("no trailing-row configuration found")
    rays, oth = _t(rows), _t(other)
    mc, mf, emb = make_nets(42, 43, False, precision)
    from pcnerf_b200.nof import render
    render.GROUP_RAYS = False                                  # the reference's evaluation: every candidate row, batch by batch
    try:
        ref = ev._render_frame_rows(mc, mf, emb, rays, oth, 32, 64, 8192, depth_inference_method=2, batch_size_set=batch)
    finally:
        render.GROUP_RAYS = True
    old = ev.FRAME_ENC_BYTES
    try:
        for budget in (old, 96 * 256 * 100):                   # one ray batch / many ray batches
            ev.FRAME_ENC_BYTES = budget
            pts = ev.render_frame(mc, mf, emb, rays, oth, 32, 64, 8192, depth_inference_method=2, batch_size_set=batch)
            assert torch.equal(pts, ref), (budget, pts.shape, ref.shape)
    finally:
        ev.FRAME_ENC_BYTES = old
    n_groups = int((rows[:, 12] >= 0).sum())
    assert ref.shape[0] == n_groups - (1 if cut else 0)
