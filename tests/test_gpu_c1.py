"""GPU: BASELINE.json configs[0] ("C1": MaiCity-00-shaped block, 1 parent + 8 child AABBs, 4,096 rays x 64 coarse + 128
importance samples, chunk 32,768, shipped training flags) against tests/golden/c1_train.npz -- the reference's own
render_rays_train + six-term loss + backward, executed in float32 (as shipped) and in float64 (same code, same inputs,
same weights: oracle/make_golden.py golden_c1).

The float64 run arbitrates the tolerances.  The reference's float32 output is itself
    coarse depth 1.1e-5,   fine depth 2.0e-4,   total loss 1.6e-5,   fine free loss 3.7e-5
(max relative) away from it: the hierarchical resampling (nof/render.py:371-412) divides by cdf differences as small as
1e-5, so 1-ulp differences of the pdf move resampled depths, and train-mode BatchNorm couples every sample of a chunk.
No float32 implementation can be "within 1e-5" of another on the fine pass; what can be asked is that it sits no further
from the float64 truth than the reference's float32 run does, up to a small factor.  Gates below:
  fp32 / affine engines: coarse depth 3e-5 of the reference float32 run and of the truth; fine depth within
      2.5 x the reference's own distance to the truth; losses 1e-4 (coarse) / 2.5 x the reference's own distance (fine).
  tensor-core engine: every depth (coarse and fine, all 4,096 rays) and every loss within 1e-3 of the reference.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import BIG, STRIDE, dev, make_nets

pytestmark = pytest.mark.gpu
LOSS_KEYS = ("child_free_loss", "child_depth_loss", "child_free_loss_fine", "child_depth_loss_fine")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.abs(b)


def _run(precision):
    from pcnerf_b200.nof import render
    g = golden("c1_train")
    rays = torch.from_numpy(g["rays"]).to(dev())
    mc, mf, emb = make_nets(42, 43, True, precision)
    res = render.render_rays_train(mc, mf, emb, rays, N_samples=int(g["S"]), N_importance=int(g["Ni"]), perturb=0,
                                   noise_std=0, chunk=int(g["chunk"]), issegmentated=1, childnerf_ratio=0.1,
                                   use_child_nerf_divide=0, use_child_nerf_loss=1)
    lam = [float(x) for x in g["lam"]]
    gt = rays[:, 14]
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")
    loss = 1e-1 * lam[0] * sl1(1e1 * res["depth"], 1e1 * gt) + 1e-1 * lam[0] * sl1(1e1 * res["depth_fine"], 1e1 * gt) \
        + lam[1] * (res["child_free_loss_fine"] + res["child_free_loss"]) \
        + lam[2] * (res["child_depth_loss_fine"] + res["child_depth_loss"])
    loss.backward()
    out = {k: v.detach().double().cpu().numpy() for k, v in res.items()}
    out["loss"] = float(loss)
    grads = {}
    for tag, m in (("c", mc), ("f", mf)):
        for k, p in m.named_parameters():
            gr = p.grad.detach().cpu().numpy()
            grads["%s_%s" % (tag, k)] = gr.reshape(-1)[::STRIDE] if k in BIG else gr
    return g, out, grads


@pytest.mark.parametrize("precision", ["fp32", "affine"])
def test_c1_fp32_engines_vs_reference_and_float64_truth(precision):
    g, out, grads = _run(precision)
    floor_f = float(_rel(g["f32_depth_fine"], g["f64_depth_fine"]).max())          # 2.0e-4 (asserted on the CPU side)
    assert float(_rel(out["depth"], g["f32_depth"]).max()) < 3e-5
    assert float(_rel(out["depth"], g["f64_depth"]).max()) < 3e-5
    assert float(_rel(out["depth_fine"], g["f64_depth_fine"]).max()) < 2.5 * floor_f
    assert float(_rel(out["depth_fine"], g["f32_depth_fine"]).max()) < 2.5 * floor_f
    for k in ("child_free_loss", "child_depth_loss"):
        assert float(_rel(out[k], g["f32_" + k])) < 1e-4 and float(_rel(out[k], g["f64_" + k])) < 1e-4, k
    for k in ("child_free_loss_fine", "child_depth_loss_fine", "loss"):
        own = max(float(_rel(g["f32_" + k], g["f64_" + k])), 1e-5)
        assert float(_rel(out[k], g["f64_" + k])) < 2.5 * own + 1e-5, (k, float(_rel(out[k], g["f64_" + k])), own)
    # parameter gradients of the coarse net against the reference's float32 autograd (fine net: inherits the resampling noise)
    for k, gr in grads.items():
        ref = g["f32_grad_" + k]
        tol = (2e-3 if k.startswith("c_") else 2e-2) * np.abs(ref).max() + 1e-12
        if k.endswith(".bias") and k.split("_", 1)[1].split(".")[0] in ("layer1", "layer2") and not k.endswith("layer2.7.bias"):
            continue       # Linear biases in front of a train-mode BN: exactly zero in exact arithmetic, noise in float32
        assert np.abs(gr - ref).max() <= tol, (k, float(np.abs(gr - ref).max()), float(np.abs(ref).max()))


def test_c1_tensor_core_engine_within_1e3_on_every_ray():
    """north_star: rendered depth and losses within 1e-3 relative for the tensor-core MLP path -- on ALL rays, both passes."""
    g, out, grads = _run("tc")
    for ref in ("f32", "f64"):
        assert float(_rel(out["depth"], g[ref + "_depth"]).max()) < 1e-3
        assert float(_rel(out["depth_fine"], g[ref + "_depth_fine"]).max()) < 1e-3
        for k in LOSS_KEYS + ("loss",):
            assert float(_rel(out[k], g["%s_%s" % (ref, k)])) < 1e-3, (ref, k)
    for k, gr in grads.items():
        ref = g["f32_grad_" + k]
        if k.endswith(".bias") and k.split("_", 1)[1].split(".")[0] in ("layer1", "layer2") and not k.endswith("layer2.7.bias"):
            continue
        assert np.abs(gr - ref).max() <= 3e-2 * np.abs(ref).max() + 1e-12, k
