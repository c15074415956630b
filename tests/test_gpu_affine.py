"""GPU: the closed-form ("affine") MLP mode -- csrc/affine.cu + the float64 parameter-sized algebra in
nof/networks/models.py.  It is exact algebra for the network as the reference builds it (identity activations), so it
is held to the fp32 gate: p / depth / losses 1e-5-class tolerances against the oracle and the reference fixtures,
parameter gradients 1e-4 of each tensor's scale."""
import numpy as np
import pytest
import torch

import pcnerf_oracle as orc
from conftest import golden
from gpu_util import assert_grads_match, dev, make_nets

pytestmark = pytest.mark.gpu


def _enc(rows, seed, spread=60.0):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.rand(rows, 3, generator=gen) - 0.5) * spread
    return orc.embedding(x)


def test_moments_kernel_vs_float64():
    from pcnerf_b200 import ops
    enc = torch.nn.functional.pad(_enc(10000, 1), (0, 1))
    m, C, cnt = ops.affine_moments(enc.to(dev()), 4096)
    e = enc.double()
    for k, i in enumerate(range(0, 10000, 4096)):
        blk = e[i:i + 4096]
        mr = blk.mean(0)
        Cr = blk.t() @ blk / blk.shape[0] - mr[:, None] * mr[None, :]
        assert float(cnt[k]) == blk.shape[0]
        # (x - s) is formed in fp32 (|x| up to 30 m): 1e-7-class relative errors, then fp64 accumulation
        np.testing.assert_allclose(m[k].cpu().numpy(), mr.numpy(), rtol=0, atol=3e-8 * 30.0)
        np.testing.assert_allclose(C[k].cpu().numpy(), Cr.numpy(), rtol=0, atol=5e-7 * float(Cr.abs().max()))   # fp32 products of (x - s), fp64 sums


@pytest.mark.parametrize("rows,chunk", [(4096, 4096), (5000, 2048), (130, 130), (70000, 16384)])
def test_forward_train_and_running_stats(rows, chunk):
    enc = _enc(rows, rows)
    sd = orc.init_state_dict(42)
    p_ref = torch.cat([orc.nof_forward(sd, enc[i:i + chunk], True) for i in range(0, rows, chunk)]).reshape(-1)
    mc, _, _ = make_nets(42, 43, True, "affine")
    p = mc.forward_encoded(torch.nn.functional.pad(enc, (0, 1)).to(dev()), chunk)
    np.testing.assert_allclose(p.detach().cpu().numpy(), p_ref.numpy(), rtol=2e-5, atol=1e-7)
    got = mc.state_dict()
    for k in ("layer1.1.running_mean", "layer1.1.running_var", "layer2.1.running_mean", "layer2.7.running_mean",
              "layer2.7.running_var"):
        np.testing.assert_allclose(got[k].cpu().numpy(), sd[k].numpy(), rtol=2e-5, atol=1e-6, err_msg=k)
    assert int(got["layer2.7.num_batches_tracked"]) == int(sd["layer2.7.num_batches_tracked"])


def test_forward_eval_and_errors():
    enc = _enc(3000, 7)
    sd = orc.init_state_dict(42)
    p_ref = orc.nof_forward(sd, enc, False).reshape(-1)
    mc, _, _ = make_nets(42, 43, False, "affine")
    with torch.no_grad():
        p = mc(enc.to(dev())).reshape(-1)
    np.testing.assert_allclose(p.cpu().numpy(), p_ref.numpy(), rtol=2e-5, atol=1e-7)
    mc.train()
    with pytest.raises(ValueError):
        mc(torch.zeros(1, 63, device=dev()))


@pytest.mark.parametrize("rows,chunk", [(4096, 4096), (3000, 1024)])
def test_backward_param_grads(rows, chunk):
    enc = _enc(rows, rows + 1)
    gen = torch.Generator().manual_seed(rows)
    gp = torch.randn(rows, generator=gen)
    sd = orc.init_state_dict(42)
    for k in orc.param_names():
        sd[k].requires_grad_(True)
    p_ref = torch.cat([orc.nof_forward(sd, enc[i:i + chunk], True) for i in range(0, rows, chunk)]).reshape(-1)
    (p_ref * gp).sum().backward()
    mc, _, _ = make_nets(42, 43, True, "affine")
    p = mc.forward_encoded(torch.nn.functional.pad(enc, (0, 1)).to(dev()), chunk)
    (p * gp.to(dev())).sum().backward()
    scale = max(float(sd[k].grad.abs().max()) for k in orc.param_names())
    for k, prm in mc.named_parameters():
        ref = sd[k].grad.numpy()
        atol = 1e-4 * np.abs(ref).max() + 1e-12
        if k.endswith(".bias") and k.split(".")[0] in ("layer1", "layer2") and k != "layer2.7.bias":
            atol = 1e-5 * scale          # exactly zero in exact arithmetic: the reference's value is rounding noise
        np.testing.assert_allclose(prm.grad.cpu().numpy(), ref, rtol=1e-4, atol=atol, err_msg=k)


@pytest.mark.parametrize("name", ["train_seg", "train_perturb", "train_plain"])
def test_render_rays_train_vs_reference(name):
    """Same gates as tests/test_gpu_render.py::_train for the fp32 engine (coarse 3e-5, fine end-to-end 2e-3)."""
    from pcnerf_b200.nof import render
    g = golden(name)
    rays = torch.from_numpy(g["rays"]).to(dev())
    perturb = float(g["perturb"])
    mc, mf, emb = make_nets(42, 43, True, "affine")
    kw = {}
    if perturb > 0:
        kw = dict(U=torch.from_numpy(g["U"]).to(dev()), u=torch.from_numpy(g["u"]).to(dev()))
    res = render.render_rays_train(mc, mf, emb, rays, N_samples=int(g["S"]), N_importance=int(g["Ni"]), perturb=perturb,
                                   noise_std=0, chunk=int(g["chunk"]), issegmentated=int(g["issegmentated"]),
                                   childnerf_ratio=float(g["ratio"]), use_child_nerf_divide=0,
                                   use_child_nerf_loss=int(g["use_child"]), **kw)
    for k in ("depth", "child_free_loss", "child_depth_loss"):
        np.testing.assert_allclose(res[k].detach().cpu().numpy(), g["out_" + k], rtol=3e-5, atol=1e-6, err_msg=k)
    for k in ("depth_fine", "child_free_loss_fine", "child_depth_loss_fine"):
        np.testing.assert_allclose(res[k].detach().cpu().numpy(), g["out_" + k], rtol=2e-3, atol=1e-6, err_msg=k)
    gt = rays[:, 14]
    lam = g["lam"]
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")
    loss = 0.1 * lam[0] * sl1(10 * res["depth"], 10 * gt) + 0.1 * lam[0] * sl1(10 * res["depth_fine"], 10 * gt)
    for k, l in (("child_free_loss_fine", lam[1]), ("child_free_loss", lam[1]), ("child_depth_loss_fine", lam[2]),
                 ("child_depth_loss", lam[2])):
        loss = loss + float(l) * res[k].to(loss.device)
    loss.backward()
    assert_grads_match("c", mc, g, 2e-3)
    assert_grads_match("f", mf, g, 1e-2)


def test_view_two_step_flags_bit_exact():
    from pcnerf_b200.nof import render
    for m in (2, 1):
        g = golden("view_m%d" % m)
        rays, other = torch.from_numpy(g["rays"]).to(dev()), torch.from_numpy(g["other"]).to(dev())
        mc, mf, emb = make_nets(42, 43, False, "affine")
        with torch.no_grad():
            r = render.render_rays_view_0525_2_2(mc, mf, emb, rays, other, N_samples=int(g["S"]),
                                                 N_importance=int(g["Ni"]), perturb=0, noise_std=0,
                                                 chunk=int(g["chunk"]), depth_inference_method=m)
        assert np.array_equal(r["rays_effective_flag"].cpu().numpy(), g["out_rays_effective_flag"])
        assert np.array_equal(r["rays_effective_flag_fine"].cpu().numpy(), g["out_rays_effective_flag_fine"])
        for k in ("depth", "depth_fine", "points_inference", "points_inference_fine"):
            np.testing.assert_allclose(r[k].cpu().numpy(), g["out_" + k], rtol=5e-5, atol=1e-6, err_msg=k)


# ---------------------------------------------------------------- the engine on rays (csrc/affine_rays.cu: no encoding tensor)


def _ray_rows(n, S, seed):
    gen = torch.Generator().manual_seed(seed)
    o = (torch.rand(1, 3, generator=gen) - 0.5) * 4.0
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=1)
    rays = torch.cat([o.expand(n, 3), d, torch.rand(n, 9, generator=gen)], 1).contiguous()
    z = torch.sort(torch.rand(n, S, generator=gen) * 40.0 + 0.5, dim=1).values.contiguous()
    return rays, z


@pytest.mark.parametrize("n,S,chunk", [(64, 64, 4096), (100, 48, 1000), (37, 192, 2048), (1, 130, 130), (700, 64, 16384)])
def test_rays_engine_vs_oracle_and_encoded_engine(n, S, chunk):
    """p, BN running statistics and every parameter gradient of the fused (rays, z) path against (a) the float32 oracle
    network on the materialised encodings and (b) the round-1 formulation of the same algebra (torch, on encodings)."""
    from pcnerf_b200 import ops
    rays, z = _ray_rows(n, S, 17 * n + S)
    rows = n * S
    gen = torch.Generator().manual_seed(rows)
    gp = torch.randn(rows, generator=gen)
    # (a) oracle
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
    enc = orc.embedding(pts)
    sd = orc.init_state_dict(42)
    for k in orc.param_names():
        sd[k].requires_grad_(True)
    p_ref = torch.cat([orc.nof_forward(sd, enc[i:i + chunk], True) for i in range(0, rows, chunk)]).reshape(-1)
    (p_ref * gp).sum().backward()
    # fused path
    mc, _, _ = make_nets(42, 43, True, "affine")
    lazy = ops.LazyEnc(rays.to(dev()), z.to(dev()))
    p = mc.forward_encoded(lazy, chunk)
    (p * gp.to(dev())).sum().backward()
    # (b) encoded path of the same engine
    mb, _, _ = make_nets(42, 43, True, "affine")
    pb = mb.forward_encoded(lazy.materialise(), chunk)
    (pb * gp.to(dev())).sum().backward()
    np.testing.assert_allclose(p.detach().cpu().numpy(), p_ref.detach().numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(p.detach().cpu().numpy(), pb.detach().cpu().numpy(), rtol=5e-6, atol=1e-7)
    got, gotb = mc.state_dict(), mb.state_dict()
    for k in got:
        if "running" in k or "num_batches" in k:
            np.testing.assert_allclose(got[k].cpu().numpy(), sd[k].detach().numpy(), rtol=2e-5, atol=1e-6, err_msg=k)
            np.testing.assert_allclose(got[k].cpu().numpy(), gotb[k].cpu().numpy(), rtol=2e-6, atol=1e-7, err_msg=k)
    scale = max(float(sd[k].grad.abs().max()) for k in orc.param_names())
    pb_grads = dict(mb.named_parameters())
    for k, prm in mc.named_parameters():
        ref = sd[k].grad.numpy()
        atol = 1e-4 * np.abs(ref).max() + 1e-12
        if k.endswith(".bias") and k.split(".")[0] in ("layer1", "layer2") and k != "layer2.7.bias":
            atol = 1e-5 * scale          # exactly zero in exact arithmetic: the reference's value is rounding noise
        np.testing.assert_allclose(prm.grad.cpu().numpy(), ref, rtol=1e-4, atol=atol, err_msg=k)
        b = pb_grads[k].grad.cpu().numpy()
        np.testing.assert_allclose(prm.grad.cpu().numpy(), b, rtol=2e-5, atol=2e-6 * scale + 1e-12, err_msg="vs encoded " + k)


def test_rays_engine_accumulates_into_existing_grads_and_rejects_eval():
    from pcnerf_b200 import ops
    rays, z = _ray_rows(50, 64, 5)
    mc, _, _ = make_nets(42, 43, True, "affine")
    lazy = ops.LazyEnc(rays.to(dev()), z.to(dev()))
    mc.forward_encoded(lazy, 1024).sum().backward()
    g1 = {k: p.grad.clone() for k, p in mc.named_parameters()}
    mc.forward_encoded(lazy, 1024).sum().backward()
    # second pass: same inputs, new running statistics but the same batch statistics -> the gradients double
    for k, p in mc.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), 2 * g1[k].cpu().numpy(), rtol=1e-5,
                                   atol=1e-6 * float(g1[k].abs().max()) + 1e-12, err_msg=k)
    with pytest.raises(ValueError):
        mc.forward_encoded(ops.LazyEnc(rays[:1].to(dev()), z[:1, :1].to(dev())), 64)
    # eval mode without autograd: the rays form of the running-statistics path (test_rays_engine_eval_mode); with autograd
    # enabled the lazy rows are materialised and take the encoded path
    mc.eval()
    pe = mc.forward_encoded(lazy, 1024)
    pm = mc.forward_encoded(lazy.materialise(), 1024)
    assert torch.equal(pe, pm)


@pytest.mark.parametrize("n,S", [(1, 1), (33, 64), (500, 192), (3000, 7)])
def test_rays_engine_eval_mode(n, S):
    """Eval mode (running statistics) on (ray, depth) rows -- pcnerf_affine_eval_alpha + pcnerf_affine_apply_rays -- against
    the float32 oracle network on the materialised encodings and against the encoded form of the same engine; the cached
    alpha follows torch-side and library-side parameter writes."""
    from pcnerf_b200 import ops
    from pcnerf_b200.optim import FlatAdam
    rays, z = _ray_rows(n, S, 31 * n + S)
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
    enc = orc.embedding(pts)
    mc, _, _ = make_nets(42, 43, True, "affine")
    lazy = ops.LazyEnc(rays.to(dev()), z.to(dev()))
    if n * S > 1:
        mc.forward_encoded(lazy, 4096)                  # a training pass first: non-trivial running statistics
    mc.eval()

    def ref():
        sd = {k: v.detach().cpu().clone() for k, v in mc.state_dict().items()}
        return orc.nof_forward(sd, enc, False).reshape(-1).numpy()

    with torch.no_grad():
        p = mc.forward_encoded(lazy, 4096)
        pm = mc.forward_encoded(lazy.materialise(), 4096)
        assert p.shape == (n * S,)
        np.testing.assert_allclose(p.cpu().numpy(), ref(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(p.cpu().numpy(), pm.cpu().numpy(), rtol=5e-6, atol=1e-7)
        assert torch.equal(p, mc.forward_encoded(lazy, 77))          # cached alpha; `chunk` changes no eval-mode value
        mc.layer2[0].weight.mul_(1.05)                               # torch-side in-place write
        mc.layer1[1].running_var.mul_(1.3)
        p2 = mc.forward_encoded(lazy, 4096)
        assert not torch.equal(p2, p)
        np.testing.assert_allclose(p2.cpu().numpy(), ref(), rtol=2e-5, atol=1e-7)
        opt = FlatAdam(list(mc.parameters()), lr=1e-3, eps=1e-8, weight_decay=1e-3)
        opt.bucket.flat.normal_(generator=torch.Generator(device=dev()).manual_seed(3))
        opt.step()                                                   # library-side write (raw pointers)
        p3 = mc.forward_encoded(lazy, 4096)
        assert not torch.equal(p3, p2)
        np.testing.assert_allclose(p3.cpu().numpy(), ref(), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("var,val", [("PCNERF_AFF_FUSED", "0"), ("PCNERF_AFF_MOMENTS", "tc")])
def test_rays_engine_alternative_kernels_in_a_subprocess(var, val):
    """Switches read once per process: PCNERF_AFF_FUSED=0 selects the GEMM + per-layer kernel pairs the fused per-layer kernels
    replaced, PCNERF_AFF_MOMENTS=tc the second-moment kernel on mma.sync with 3 x TF32 operands.  The same parity tests must
    pass on those paths too."""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    env[var] = val
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(here, "test_gpu_affine.py"), "-k",
                        "test_rays_engine_vs_oracle_and_encoded_engine or test_rays_engine_eval_mode"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
