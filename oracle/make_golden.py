"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):   python oracle/make_golden.py
The fixtures pin `oracle/pcnerf_oracle.py` (tests/test_oracle_golden.py) and are the golden vectors the
CUDA path is compared with on the GPU box (tests/test_gpu_*.py), where the reference tree is absent.

Every fixture stores the inputs it was produced from, the seeds, and the reference's outputs.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
import pcnerf_oracle as orc  # noqa: E402
from pcnerf_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BIG = ("layer1.3.weight", "layer1.6.weight", "layer1.9.weight", "layer2.0.weight", "layer2.2.weight",
       "layer2.4.weight", "layer2.6.weight")
STRIDE = 17


def compress_grads(named):
    out = {}
    for k, g in named.items():
        g = g.detach().numpy()
        out[k] = g.reshape(-1)[::STRIDE].copy() if k in BIG else g.copy()
    return out


def load_models(ref, seed_c, seed_f, train):
    mc, mf = ref.networks.NOF_coarse(), ref.networks.NOF_fine()
    mc.load_state_dict(orc.init_state_dict(seed_c))
    mf.load_state_dict(orc.init_state_dict(seed_f))
    mc.train(train)
    mf.train(train)
    return mc, mf, ref.networks.Embedding(3, 10)


# ------------------------------------------------------------------------------------------------ AABB


def golden_aabb(ref):
    rng = np.random.default_rng(11)
    scene = synth.make_scene(3, 40)
    pts = synth.make_points(scene, 3, 300)
    # a few points outside every box
    pts[::25] = rng.uniform(scene.parent_min, scene.parent_max, size=pts[::25].shape)
    dirs, dist = synth.rays_from_points(scene.origin, pts)
    o = scene.origin
    x_min, x_max, y_min, y_max, z_min, z_max = scene.parent
    far_parent = np.array([ref.ipb.compute_far_bound(o, d, x_max, x_min, y_max, y_min, z_max, z_min) for d in dirs],
                          dtype=np.float64)
    inside, idx = [], []
    for p in pts:
        f, i = ref.ipb.find_aabb_box(scene.centres, scene.child_bounds, p)
        inside.append(f)
        idx.append(-1 if i is None else i)
    inside, idx = np.array(inside), np.array(idx, dtype=np.int64)
    # ray x box grids for the three child-intersection variants
    nb = scene.K
    f0429 = np.zeros((pts.shape[0], nb), dtype=bool)
    n0429 = np.zeros((pts.shape[0], nb))
    r0429 = np.zeros((pts.shape[0], nb))
    f0606 = np.zeros((pts.shape[0], nb), dtype=bool)
    n0606 = np.zeros((pts.shape[0], nb))
    r0606 = np.zeros((pts.shape[0], nb))
    n0406 = np.full((pts.shape[0], nb), np.nan)
    r0406 = np.full((pts.shape[0], nb), np.nan)
    for i, d in enumerate(dirs):
        for k in range(nb):
            bmin, bmax = scene.child_bounds_bigger[k][:3], scene.child_bounds_bigger[k][3:6]
            f0429[i, k], n0429[i, k], r0429[i, k] = ref.evalmod.compute_far_bound0429(o, d, bmin, bmax)
            f0606[i, k], n0606[i, k], r0606[i, k] = ref.ipb.compute_far_bound0606(o, d, bmin, bmax)
            try:
                n0406[i, k], r0406[i, k] = ref.ipb.compute_far_bound0406(o, d, bmin, bmax)
            except IndexError:
                pass
    slab = ref.evalmod.ray_aabb_distances(o, dirs, scene.parent_min, scene.parent_max)
    centre = (scene.child_bounds[:, :3] + scene.child_bounds[:, 3:]) / 2
    dtr = np.stack([ref.evalmod.distance_to_ray(torch.tensor(o), d, centre) for d in dirs[:64]])
    np.savez_compressed(os.path.join(OUT, "aabb_leaf.npz"), origin=o, points=pts, dirs=dirs, dist=dist,
                        parent=np.array(scene.parent), centres=scene.centres, child_bounds=scene.child_bounds,
                        child_bounds_bigger=scene.child_bounds_bigger, far_parent=far_parent, inside=inside, idx=idx,
                        f0429=f0429, n0429=n0429, r0429=r0429, f0606=f0606, n0606=n0606, r0606=r0606,
                        n0406=n0406, r0406=r0406, slab=slab, dist_to_ray=dtr)

    # ---- train packing loop body (ipb2dmapping.py:367-397 MaiCity / :736-768 KITTI) around the imported leaf fns
    for variant in ("maicity", "kitti"):
        se = 0.05
        rows = []
        for i in range(pts.shape[0]):
            ok, a = ref.ipb.find_aabb_box(scene.centres, scene.child_bounds, pts[i])
            if not ok:
                continue
            bb = scene.child_bounds_bigger[a]
            if variant == "maicity":
                try:
                    nbd, fbd = ref.ipb.compute_far_bound0406(o, dirs[i], bb[:3], bb[3:6])
                except IndexError:
                    nbd, fbd = np.nan, np.nan
            else:
                inter, nbd, fbd = ref.ipb.compute_far_bound0606(o, dirs[i], bb[:3], bb[3:6])
                if not inter:
                    continue
            nbd, fbd = nbd - se, fbd + se
            fp = ref.ipb.compute_far_bound(o, dirs[i], x_max, x_min, y_max, y_min, z_max, z_min)
            if fp < fbd:
                fp = fbd
            rows.append([*o, *dirs[i], 0.0, fp, 3, a + 1, nbd, fbd, dist[i] - se, fbd, dist[i]])
        rays = torch.Tensor(np.asarray(rows, dtype=float)).numpy()
        np.savez_compressed(os.path.join(OUT, "aabb_pack_%s.npz" % variant), rays=rays, surface_expand=se)

    # ---- candidate-group loop body (eval_kitti_render.py:359-461 MaiCity 0.005 / :681-803 KITTI 0.05)
    sbl = scene.child_bounds + np.array([-0.025] * 3 + [0.025] * 3)
    nray = 160
    pf = ref.evalmod.ray_aabb_distances(o, dirs[:nray], scene.parent_min, scene.parent_max)
    for method in (2, 1):
        for grow in (0.005, 0.05):
            all_rows, all_other = [], []
            for i in range(nray):
                center = (scene.child_bounds[:, :3] + scene.child_bounds[:, 3:]) / 2
                dd = ref.evalmod.distance_to_ray(torch.tensor(o), dirs[i], center)
                filt = sbl[dd <= 0.65]
                cands = []
                hit = False

                def scan():
                    res = []
                    for k in range(filt.shape[0]):
                        fl, nb_, fb_ = ref.evalmod.compute_far_bound0429(o, dirs[i], filt[k][:3], filt[k][3:6])
                        if fl:
                            if method == 1:
                                res.append((0.0, pf[i]))
                                break
                            res.append((nb_, fb_))
                    return res

                cands = scan()
                hit = len(cands) > 0
                extend_iter = 0
                drop = False
                while not hit:
                    if extend_iter > 0.5:
                        drop = True
                        break
                    extend_iter = extend_iter + grow
                    filt[:, :3] = filt[:, :3] - extend_iter
                    filt[:, 3:6] = filt[:, 3:6] + extend_iter
                    cands = scan()
                    hit = len(cands) > 0
                if drop:
                    continue
                arr = np.zeros((len(cands), 12))
                for r, (a, b) in enumerate(cands):
                    arr[r] = [*o, *dirs[i], a, b, 3, dist[i], 0.0, pf[i]]
                order = np.argsort(arr[:, 6])
                arr = arr[order]
                arr = np.hstack((arr, np.arange(arr.shape[0]).reshape(-1, 1) + 1))
                arr = np.concatenate((arr, -1 * np.ones((arr.shape[0], 1))), axis=1)
                arr[0][-1] = len(cands) - 1
                oth = np.zeros((arr.shape[0], 1), dtype=int)
                oth[0] = len(cands) - 1
                all_rows.append(arr)
                all_other.append(oth)
            rows = torch.Tensor(np.concatenate(all_rows, 0)).float()
            all_rays = torch.cat([rows[:, :9], rows[:, 10:]], dim=1).numpy()
            np.savez_compressed(os.path.join(OUT, "aabb_groups_m%d_g%s.npz" % (method, str(grow).replace(".", "p"))),
                                rays=all_rays, ranges=rows[:, 9:10].numpy(), other=np.concatenate(all_other, 0),
                                nray=nray, grow=grow, method=method)


# ------------------------------------------------------------------------------------------------ training


def golden_train(ref, name, seed, N, S, Ni, chunk, perturb, issegmentated, ratio, use_child, K=8):
    rays = torch.from_numpy(synth.synth_train_rays(seed, N, K=K))
    mc, mf, emb = load_models(ref, 42, 43, train=True)
    F_ = S + Ni
    torch.manual_seed(seed)
    U = torch.rand(N, S) if perturb > 0 else torch.zeros(0)
    torch.randn(N, S)
    u = torch.rand(N, Ni) if perturb > 0 else torch.zeros(0)
    torch.randn(N, F_)
    torch.manual_seed(seed)
    with ref_shim.cuda0_to_cpu():
        res = ref.render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=Ni, perturb=perturb,
                                           noise_std=0, chunk=chunk, issegmentated=issegmentated,
                                           childnerf_ratio=ratio, use_child_nerf_divide=0,
                                           use_child_nerf_loss=use_child)
    gt = rays[:, 14]
    lam = (1.0, 1e6, 1e5)
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")
    loss = 1e-1 * lam[0] * sl1(1e1 * res["depth"], 1e1 * gt) + 1e-1 * lam[0] * sl1(1e1 * res["depth_fine"], 1e1 * gt) \
        + lam[1] * res["child_free_loss_fine"] + lam[1] * res["child_free_loss"] \
        + lam[2] * res["child_depth_loss_fine"] + lam[2] * res["child_depth_loss"]
    loss.backward()
    save = dict(rays=rays.numpy(), U=U.numpy(), u=u.numpy(), seed=seed, S=S, Ni=Ni, chunk=chunk, perturb=perturb,
                issegmentated=issegmentated, ratio=ratio, use_child=use_child, lam=np.array(lam), loss=loss.item())
    for k, v in res.items():
        save["out_" + k] = v.detach().numpy()
    for tag, m in (("c", mc), ("f", mf)):
        for k, g in compress_grads({n: p.grad for n, p in m.named_parameters()}).items():
            save["grad_%s_%s" % (tag, k)] = g
        sd = m.state_dict()
        for k in ("layer1.1.running_mean", "layer1.1.running_var", "layer2.7.running_mean", "layer2.7.running_var"):
            save["bn_%s_%s" % (tag, k)] = sd[k].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "loss", loss.item(), {k: float(v.detach().sum()) for k, v in res.items()})


def golden_system(ref):
    """train_kitti.py:117-156 through NOFSystem.training_step itself."""
    import argparse as ap
    hp = ap.Namespace(L_pos=10, feature_size=256, use_skip=True, ckpt_path=None, loss_type="smoothl1", N_samples=32,
                      N_importance=64, use_disp=False, perturb=0, noise_std=0, chunk=4096, sub_nerf_test_num=8,
                      use_segmentated_sample=1, segmentated_child_nerf_ratio=0.1, use_child_nerf_divide=0,
                      use_child_nerf_loss=1, lambda_loss=1.0, lambda_loss_fine=1.0, lambda_child_free_loss=1e6,
                      lambda_child_depth_loss=1e5, optimizer="adam", lr=5e-4, momentum=0.9, weight_decay=1e-3,
                      decay_gamma=0.1, current_epoch=0, visualize=0)
    sys_ = ref.trainmod.NOFSystem(hp)
    sys_.nof_coarse.load_state_dict(orc.init_state_dict(42))
    sys_.nof_fine.load_state_dict(orc.init_state_dict(43))
    sys_.train()
    sys_.configure_optimizers()
    rays = torch.from_numpy(synth.synth_train_rays(5, 128, K=8))
    with ref_shim.cuda0_to_cpu():
        loss = sys_.training_step({"rays": rays, "ranges": rays[:, 14]}, 1)
    loss.backward()
    sys_.optimizer.step()
    save = dict(rays=rays.numpy(), loss=loss.item())
    for tag, m in (("c", sys_.nof_coarse), ("f", sys_.nof_fine)):
        sd = m.state_dict()
        for k in ("layer1.0.bias", "layer1.1.weight", "layer2.7.bias", "occ_out.0.weight", "occ_out.0.bias"):
            save["after_%s_%s" % (tag, k)] = sd[k].numpy()
    np.savez_compressed(os.path.join(OUT, "system_step.npz"), **save)
    print("system_step loss", loss.item())


# ------------------------------------------------------------------------------------------------ val / legacy / view


def golden_val(ref):
    N, S, Ni = 96, 64, 128
    rays = torch.from_numpy(synth.synth_train_rays(9, N, K=8))
    mc, mf, emb = load_models(ref, 42, 43, train=False)
    with torch.no_grad(), ref_shim.cuda0_to_cpu():
        res = ref.render.render_rays_val(mc, mf, emb, rays, N_samples=S, N_importance=Ni, perturb=0, noise_std=0,
                                         chunk=4096)
        leg = ref.render.render_rays(mc, mf, emb, rays, N_samples=S, N_importance=Ni, perturb=0, noise_std=0,
                                     chunk=4096, isval=False)
        leg_d = ref.render.render_rays(mc, mf, emb, rays, N_samples=S, N_importance=Ni, use_disp=True, perturb=0,
                                       noise_std=0, chunk=4096, isval=True)
    raysn = rays.numpy().copy()
    save = dict(rays=raysn, S=S, Ni=Ni, chunk=4096)
    for k, v in res.items():
        save["val_" + k] = v.numpy()
    for k, v in leg.items():
        save["leg_" + k] = v.numpy()
    for k, v in leg_d.items():
        save["legdisp_" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "val_legacy.npz"), **save)


def golden_view(ref):
    rows, other, _ = synth.synth_infer_rows(21, 60)
    rays = torch.from_numpy(rows)
    oth = torch.from_numpy(other)
    S, Ni = 64, 128
    for method in (2, 1):
        mc, mf, emb = load_models(ref, 42, 43, train=False)
        with torch.no_grad(), ref_shim.cuda0_to_cpu():
            res = ref.render.render_rays_view_0525_2_2(mc, mf, emb, rays, oth, N_samples=S, N_importance=Ni,
                                                        perturb=0, noise_std=0, chunk=8192,
                                                        depth_inference_method=method)
        save = dict(rays=rows, other=other, S=S, Ni=Ni, chunk=8192, method=method)
        for k, v in res.items():
            save["out_" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, "view_m%d.npz" % method), **save)
        print("view", method, int(res["rays_effective_flag_fine"].sum()), "winners of", rays.shape[0], "rows")


def golden_heads(ref):
    """Head-only fixtures with given p (no MLP): inference_train / inference_0525_2 through a stub model, so
    masks / flags can be compared bit-exactly, plus sample_pdf."""
    N, S = 200, 64
    rays = torch.from_numpy(synth.synth_train_rays(31, N, K=8))
    g = torch.Generator().manual_seed(31)
    logits = torch.randn(N, S, generator=g) * 2 - 2
    # put a bump near the return so that weights look like a trained net
    z = orc.sample_z(rays, S, 1, 0.1, 0, None)
    logits = logits + 6 * torch.exp(-0.5 * ((z - rays[:, 14:15]) / 0.3) ** 2)
    p = torch.sigmoid(logits)

    class Stub(torch.nn.Module):
        def __init__(self, p):
            super().__init__()
            self.p = torch.nn.Parameter(p.reshape(-1, 1).clone())
            self.i = 0

        def forward(self, x):
            out = self.p[self.i:self.i + x.shape[0]]
            self.i += x.shape[0]
            return out

    stub = Stub(p)
    emb = ref.networks.Embedding(3, 10)
    pts = rays[:, :3].unsqueeze(1) + rays[:, 3:6].unsqueeze(1) * z.unsqueeze(2)
    fl, dl, depth, w = ref.render.inference_train(stub, emb, pts, rays, z, rays[:, 10:12], rays[:, 12:14],
                                                   rays[:, -1].view(-1, 1), rays[:, 8].view(-1, 1), chunk=4096,
                                                   noise_std=0, epsilon=1e-10, use_child_nerf_divide=0,
                                                   use_child_nerf_loss=1)
    sl1 = torch.nn.SmoothL1Loss(reduction="mean")
    loss = 0.1 * sl1(10 * depth, 10 * rays[:, 14]) + 1e6 * fl + 1e5 * dl
    loss.backward()
    gp = stub.p.grad.reshape(N, S)
    # uniform-z variant exercises the expand-until-non-empty loop
    zu = orc.sample_z(rays, 24, 0, 0.1, 0, None)
    pu = torch.sigmoid(torch.randn(N, 24, generator=g))
    stub2 = Stub(pu)
    ptsu = rays[:, :3].unsqueeze(1) + rays[:, 3:6].unsqueeze(1) * zu.unsqueeze(2)
    flu, dlu, depthu, wu = ref.render.inference_train(stub2, emb, ptsu, rays, zu, rays[:, 10:12], rays[:, 12:14],
                                                       rays[:, -1].view(-1, 1), rays[:, 8].view(-1, 1), chunk=4096,
                                                       noise_std=0, epsilon=1e-10, use_child_nerf_divide=0,
                                                       use_child_nerf_loss=1)
    (1e6 * flu + 1e5 * dlu + depthu.sum()).backward()
    # sample_pdf (det and with given u)
    mid = .5 * (z[..., 1:] + z[..., :-1])
    with ref_shim.cuda0_to_cpu():
        zs_det = ref.render.sample_pdf(mid, w.detach()[..., 1:-1], 128, det=True)
        torch.manual_seed(77)
        u = torch.rand(N, 128)
        torch.manual_seed(77)
        zs_rnd = ref.render.sample_pdf(mid, w.detach()[..., 1:-1], 128, det=False)
    np.savez_compressed(os.path.join(OUT, "head_train.npz"), rays=rays.numpy(), z=z.numpy(), p=p.numpy(),
                        free=fl.item(), depthloss=dl.item(), depth=depth.detach().numpy(), w=w.detach().numpy(),
                        grad_p=gp.numpy(), zu=zu.numpy(), pu=pu.numpy(), free_u=flu.item(), depthloss_u=dlu.item(),
                        depth_u=depthu.detach().numpy(), grad_pu=stub2.p.grad.reshape(N, 24).numpy(),
                        zs_det=zs_det.numpy(), u=u.numpy(), zs_rnd=zs_rnd.numpy())

    # search head
    rows, other, r = synth.synth_infer_rows(41, 80)
    rv = torch.from_numpy(rows)
    oth = torch.from_numpy(other)
    Nv = rv.shape[0]
    zv = orc.sample_z(rv, 96, 0, 0.5, 0, None, near_col=9, far_col=10)
    lg = torch.randn(Nv, 96, generator=g) * 1.5 - 3
    lg = lg + 5 * torch.exp(-0.5 * ((zv - 0.5 * (rv[:, 6:7] + rv[:, 7:8])) / 0.5) ** 2) * (torch.rand(Nv, 1, generator=g) > 0.4)
    pv = torch.sigmoid(lg)
    ptsv = rv[:, :3].unsqueeze(1) + rv[:, 3:6].unsqueeze(1) * zv.unsqueeze(2)
    save = dict(rays=rows, other=other, z=zv.numpy(), p=pv.numpy())
    for method in (2, 1):
        with torch.no_grad():
            d_, w_, op_, fl_ = ref.render.inference_0525_2(Stub(pv), emb, ptsv, zv, oth, rv[:, 6:8], chunk=1 << 20,
                                                            noise_std=0, epsilon=1e-10, depth_inference_method=method)
        save.update({"depth_m%d" % method: d_.numpy(), "w_m%d" % method: w_.numpy(), "opacity_m%d" % method: op_.item(),
                     "flag_m%d" % method: fl_.numpy()})
    np.savez_compressed(os.path.join(OUT, "head_search.npz"), **save)


def golden_c1(ref):
    """BASELINE.json configs[0] ("C1"): MaiCity-00-shaped block, 1 parent + 8 child AABBs, 4,096 rays x 64 coarse (+128
    importance, render_rays_train's own default, nof/render.py:417) samples, chunk 32,768, shipped training flags (segmented
    sampling ratio 0.1, child losses on), perturb 0 (deterministic) -- the reference's own render_rays_train + the six-term
    loss of train_kitti.py:145-155 + backward, executed TWICE: as shipped (float32) and with every tensor in float64
    (torch default dtype float64: same code, same inputs, same weights) as the ground truth that arbitrates which
    float32 differences are the reference's own rounding noise."""
    import time
    N, S, Ni, chunk, K = 4096, 64, 128, 32768, 8
    rays32 = torch.from_numpy(synth.synth_train_rays(101, N, K=K))
    lam = (1.0, 1e6, 1e5)
    save = dict(rays=rays32.numpy(), S=S, Ni=Ni, chunk=chunk, K=K, lam=np.array(lam))

    sd_c, sd_f = orc.init_state_dict(42), orc.init_state_dict(43)      # float32 values (drawn BEFORE the dtype switch)

    def run(dtype, tag):
        old = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            # parameters in the default dtype, holding exactly the float32 values
            mc, mf, emb = ref.networks.NOF_coarse(), ref.networks.NOF_fine(), ref.networks.Embedding(3, 10)
            mc.load_state_dict(sd_c)
            mf.load_state_dict(sd_f)
            mc.train()
            mf.train()
            rays = rays32.to(dtype)
            t0 = time.time()
            with ref_shim.cuda0_to_cpu():
                res = ref.render.render_rays_train(mc, mf, emb, rays, N_samples=S, N_importance=Ni, perturb=0,
                                                   noise_std=0, chunk=chunk, issegmentated=1, childnerf_ratio=0.1,
                                                   use_child_nerf_divide=0, use_child_nerf_loss=1)
            gt = rays[:, 14]
            sl1 = torch.nn.SmoothL1Loss(reduction="mean")
            lr_c = 1e-1 * lam[0] * sl1(1e1 * res["depth"], 1e1 * gt)
            lr_f = 1e-1 * lam[0] * sl1(1e1 * res["depth_fine"], 1e1 * gt)
            loss = lr_c + lr_f + lam[1] * res["child_free_loss_fine"] + lam[1] * res["child_free_loss"] \
                + lam[2] * res["child_depth_loss_fine"] + lam[2] * res["child_depth_loss"]
            loss.backward()
            dt = time.time() - t0
        finally:
            torch.set_default_dtype(old)
        save["%s_loss" % tag] = loss.item()
        save["%s_loss_range" % tag] = lr_c.item()
        save["%s_loss_range_fine" % tag] = lr_f.item()
        save["%s_seconds" % tag] = dt
        for k, v in res.items():
            save["%s_%s" % (tag, k)] = v.detach().numpy()
        for t_, m in (("c", mc), ("f", mf)):
            for k, g in compress_grads({n: p.grad for n, p in m.named_parameters()}).items():
                save["%s_grad_%s_%s" % (tag, t_, k)] = g
        print("c1", tag, "loss", loss.item(), "%.1f s" % dt, {k: float(v.detach().sum()) for k, v in res.items()})
        return res

    r32 = run(torch.float32, "f32")
    r64 = run(torch.float64, "f64")
    for k in ("depth", "depth_fine"):
        a, b = r32[k].detach().double().numpy(), r64[k].detach().numpy()
        rel = np.abs(a - b) / np.abs(b)
        print("c1 reference float32 vs float64 %s: median %.2e p99 %.2e max %.2e" % (k, np.median(rel), np.quantile(rel, .99), rel.max()))
    np.savez_compressed(os.path.join(OUT, "c1_train.npz"), **save)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.import_reference()
    torch.set_num_threads(8)
    jobs = {
        "aabb": lambda: golden_aabb(ref),
        "heads": lambda: golden_heads(ref),
        "train_seg": lambda: golden_train(ref, "train_seg", 1, 96, 64, 128, 4096, 0, 1, 0.1, 1),
        "train_perturb": lambda: golden_train(ref, "train_perturb", 2, 64, 64, 128, 8192, 1.0, 1, 0.1, 1),
        "train_plain": lambda: golden_train(ref, "train_plain", 3, 64, 32, 64, 1024, 0, 0, 0.5, 0),
        "system": lambda: golden_system(ref),
        "val": lambda: golden_val(ref),
        "view": lambda: golden_view(ref),
        "c1": lambda: golden_c1(ref),
    }
    for k, fn in jobs.items():
        if a.only and a.only != k:
            continue
        print("==", k)
        fn()


if __name__ == "__main__":
    main()
