"""GPU: BASELINE.json's full sizes (configs[1]: 32,768 rays x 64 + 128 samples; configs[2]: one 131,072-ray frame), where
the CPU oracle is too slow to be the checker: size-independent properties of the path plus cross-engine agreement."""
import numpy as np
import pytest
import torch

from gpu_util import dev, make_nets

pytestmark = pytest.mark.gpu
N, S, NI, CHUNK = 32768, 64, 128, 262144


def _rays():
    from pcnerf_b200 import synth
    return torch.from_numpy(synth.synth_train_rays(2024, N, K=200, parent=synth.KITTI_PARENT)).to(dev())


def test_sampling_and_compositing_invariants_c2():
    from pcnerf_b200 import ops
    rays = _rays()
    U = torch.rand((N, S), device=dev(), generator=torch.Generator(device=dev()).manual_seed(1))
    z, enc = ops.sample_encode_coarse(rays, 57, 7, 6, 7, 10, 11, False, 1.0, U, True, False)
    assert bool((z[:, 1:] >= z[:, :-1]).all())                                   # sorted (render.py:442)
    assert bool((z[:, 0] >= rays[:, 6] - 1e-6).all()) and bool((z[:, -1] <= rays[:, 7] + 1e-6).all())
    e = enc.view(N, S, 64)
    assert bool((e[..., 63] == 0).all())
    x = rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]
    assert torch.equal(e[..., :3], x)                                            # o + d*z, no FMA (render.py:458)
    sc = e[..., 3:63].reshape(N, S, 10, 2, 3)
    assert float((sc[..., 0, :] ** 2 + sc[..., 1, :] ** 2 - 1).abs().max()) < 1e-5   # sin^2 + cos^2
    p = torch.sigmoid(torch.randn((N, S), device=dev(), generator=torch.Generator(device=dev()).manual_seed(2)) * 2 - 1)
    w, depth, fl, dl, *_ = ops.composite(p, z, rays, (10, 11, 14), None, 0.0, 1e-10, ops.COMP_CHILD_LOSS)
    assert bool((w >= 0).all()) and float((w.sum(1) - 1).abs().max()) < 1e-5      # normalised weights (render.py:59)
    assert bool((depth >= z[:, 0] - 1e-4).all()) and bool((depth <= z[:, -1] + 1e-4).all())
    assert float(fl) >= 0 and float(dl) >= 0
    # resampling: the fine depths are the sorted union of the coarse depths and N_importance new ones (render.py:467)
    u = torch.rand((N, NI), device=dev(), generator=torch.Generator(device=dev()).manual_seed(3))
    zf, _ = ops.sample_encode_fine(rays, z, w, NI, u, False, want_enc=False)
    assert zf.shape == (N, S + NI) and bool((zf[:, 1:] >= zf[:, :-1]).all())
    pos = torch.searchsorted(zf.contiguous(), z.contiguous())
    assert torch.equal(torch.gather(zf, 1, pos.clamp_max(S + NI - 1)), z)          # every coarse depth survives the merge
    mid_lo, mid_hi = 0.5 * (z[:, 0] + z[:, 1]), 0.5 * (z[:, -2] + z[:, -1])
    assert bool((zf[:, 0] >= z[:, 0] - 1e-6).all()) and bool((zf[:, -1] <= z[:, -1] + 1e-6).all())
    inside = ((zf >= mid_lo[:, None] - 1e-5) & (zf <= mid_hi[:, None] + 1e-5)).sum(1)
    assert bool((inside >= NI).all())                                            # new samples lie within the bins


def test_engines_agree_on_a_full_training_step_c2():
    """One C2 step on the three MLP engines from identical weights and random numbers: fp32 CUDA cores vs the closed form
    at the fp32 gate, tcgen05 at its 1e-3 gate (coarse pass; the fine pass inherits the resampling conditioning)."""
    from pcnerf_b200.nof import render
    rays = _rays()
    U = torch.rand((N, S), device=dev(), generator=torch.Generator(device=dev()).manual_seed(4))
    u = torch.rand((N, NI), device=dev(), generator=torch.Generator(device=dev()).manual_seed(5))
    out = {}
    for prec in ("fp32", "affine", "tc"):
        mc, mf, emb = make_nets(42, 43, True, prec)
        res = render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=NI, perturb=1.0, noise_std=0,
                                       chunk=CHUNK, issegmentated=1, childnerf_ratio=0.1, use_child_nerf_divide=0,
                                       use_child_nerf_loss=1, U=U, u=u)
        loss = res["depth"].mean() + res["depth_fine"].mean() + 1e6 * (res["child_free_loss"] + res["child_free_loss_fine"])
        loss.backward()
        out[prec] = ({k: v.detach().float().cpu().numpy() for k, v in res.items()},
                     mc.occ_out[0].weight.grad.cpu().numpy(), mc.layer1[3].weight.grad.cpu().numpy())
        del mc, mf, res, loss
        torch.cuda.empty_cache()
    ref = out["fp32"]
    for k in ("depth", "child_free_loss", "child_depth_loss"):
        np.testing.assert_allclose(out["affine"][0][k], ref[0][k], rtol=3e-5, atol=1e-6, err_msg=k)
    # tensor-core engine (north_star: depth and losses within 1e-3 relative): EVERY one of the 32,768 coarse and fine depths
    # and all four losses.  Measured with the linear weight-rounding correction of k_tc_fold: worst coarse depth 8.1e-4,
    # worst fine depth 7.1e-4 (1.17e-3 / 9.2e-4 without it), losses 1e-6 .. 5e-6 (profiles/r02_tc_error_c2.json).
    for k in ("child_free_loss", "child_depth_loss", "child_free_loss_fine", "child_depth_loss_fine"):
        np.testing.assert_allclose(out["tc"][0][k], ref[0][k], rtol=1e-3, err_msg=k)
    rel = np.abs(out["tc"][0]["depth"] - ref[0]["depth"]) / np.abs(ref[0]["depth"])
    assert rel.max() < 1e-3, (np.quantile(rel, 0.999), rel.max())
    relf = np.abs(out["tc"][0]["depth_fine"] - ref[0]["depth_fine"]) / np.abs(ref[0]["depth_fine"])
    assert relf.max() < 1e-3, (np.quantile(relf, 0.999), relf.max())
    np.testing.assert_allclose(out["affine"][0]["depth_fine"], ref[0]["depth_fine"], rtol=2e-3, atol=1e-5)
    for i in (1, 2):
        scale = np.abs(ref[i]).max()
        assert np.abs(out["affine"][i] - ref[i]).max() <= 1e-3 * scale
        assert np.abs(out["tc"][i] - ref[i]).max() <= 3e-2 * scale


def test_one_winner_per_physical_ray_full_frame_c3():
    from pcnerf_b200 import eval_kitti_render as ev
    from pcnerf_b200 import synth
    rows, other, _ = synth.synth_infer_rows(77, 2048)
    reps = 64                                                        # 131,072 physical rays
    rays = torch.from_numpy(np.tile(rows, (reps, 1))).to(dev())
    oth = torch.from_numpy(np.tile(other, reps)).to(dev())
    mc, mf, emb = make_nets(42, 43, False, "affine")
    pts = ev.render_frame(mc, mf, emb, rays, oth, S, NI, 184320, depth_inference_method=2, batch_size_set=18432)
    assert pts.shape == (2048 * reps, 3) and bool(torch.isfinite(pts).all())
    # idempotence of the tiling: the same physical ray renders to the same point in every repetition
    p = pts.view(reps, 2048, 3)
    assert float((p - p[:1]).abs().max()) < 1e-4
    # every rendered point lies on its ray between the parent bounds
    heads = np.nonzero(rows[:, 12] >= 0)[0]
    o = torch.from_numpy(rows[heads, :3]).to(dev())
    d = torch.from_numpy(rows[heads, 3:6]).to(dev())
    t = ((p[0] - o) * d).sum(1)
    assert float(((p[0] - o) - t[:, None] * d).abs().max()) < 1e-3
    far = torch.from_numpy(rows[heads, 10]).to(dev())
    assert bool((t >= -1e-3).all()) and bool((t <= far + 1e-3).all())
