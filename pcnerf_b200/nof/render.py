"""Drop-in for nof/render.py: same function names, positional/keyword parameters, defaults and result-dict keys
(reference signatures at render.py:13-15, :38-40, :166-167, :229-231, :371, :416-418, :485-486, :538-539, :614-616).
Every computation is a kernel of libpcnerf_b200.so; tensors must live on a CUDA device.

Extra keyword-only arguments (all optional, default = reference behaviour):
  U        pre-drawn stratified-jitter numbers (N,S)   -- the reference draws torch.rand on rays.device (:453)
  u        pre-drawn sample_pdf numbers (N,N_importance) -- the reference draws torch.rand on the CPU generator and
           copies them to "cuda:0" (:383,:397); here they are drawn on rays.device unless RNG_ON_CPU is set
  noise, noise_fine   pre-drawn torch.randn tensors for noise_std != 0 (:57)
Random streams: with noise_std == 0 the reference still consumes torch.randn numbers (:57); this path does not.
"""
import torch

from .. import ops
from .networks import NOF, Embedding, NOF_coarse, NOF_fine, NOF_plusfine  # noqa: F401

__all__ = ['render_rays']

RNG_ON_CPU = False      # True: draw sample_pdf's u with the CPU generator like nof/render.py:383 (slow: H2D each call)


def _tc(model):
    return model.mlp_precision() == 1


def _lazy(model):
    """Pass of a closed-form (precision 2) model in training mode, or in eval mode without autograd: K2 / K2' write depths
    only, the engine re-derives the encodings from (rays, z)."""
    return model.mlp_precision() == 2 and (model.training or not torch.is_grad_enabled())


def _draw_u(n, Ni, device):
    if RNG_ON_CPU:
        return torch.rand(n, Ni).to(device)
    return torch.rand(n, Ni, device=device)


def _noise(shape, device, noise_std, given):
    if noise_std == 0:
        return None
    return given if given is not None else torch.randn(shape, device=device)


def _segments(N_samples, issegmentated, childnerf_ratio):
    if issegmentated == 0:
        return N_samples, 0
    n_parent = int(N_samples * (1 - childnerf_ratio))      # render.py:434-435
    return n_parent, N_samples - n_parent


def _embed_samples(samples_xy, f16):
    enc = ops.embed(samples_xy.reshape(-1, 3).contiguous(), 64)
    return enc.to(torch.float16) if f16 else enc


# ----------------------------------------------------------------------------------------------------- inference_*


def inference_val(model, embedding_xy, samples_xy, rays, z_vals, near_far_child, near_far_point, range_readings,
                  ray_class, chunk=1024 * 32, noise_std=1, epsilon=1e-10, isval=False, sub_nerf_test_num=4, *,
                  noise=None, _enc=None):
    """nof/render.py:13-36."""
    N_rays, N_samples = z_vals.shape
    enc = _enc if _enc is not None else _embed_samples(samples_xy, _tc(model))
    p = model.forward_encoded(enc, chunk).view(N_rays, N_samples)
    nz = _noise(z_vals.shape, z_vals.device, noise_std, noise)
    w, depth = ops.composite(p, z_vals, None, (0, 0, 0), nz, noise_std, epsilon, 0)[:2]
    return depth, w


def inference_train(model, embedding_xy, samples_xy, rays, z_vals, near_far_child, near_far_point, range_readings,
                    ray_class, chunk=1024 * 32, noise_std=1, epsilon=1e-10, isval=False, sub_nerf_test_num=4,
                    use_child_nerf_divide=1, use_child_nerf_loss=0, *, noise=None, _enc=None, _extra=None):
    """nof/render.py:38-163.  `near_far_child` / `range_readings` are read from `rays` columns 10:12 / -1 exactly as
    the reference's caller packs them (render.py:424-427).
    _extra (dict, optional): receives 'range_sl1' = SmoothL1Loss(mean)(10 depth, 10 rays[:, -1]) -- the scene-level range term
    of train_kitti.py:145-146 before its 0.1 * lambda_loss factor -- accumulated by K4 in the same pass (and back-propagated
    by K4's backward), so the caller needs no separate loss kernels."""
    N_rays, N_samples = z_vals.shape
    enc = _enc if _enc is not None else _embed_samples(samples_xy, _tc(model))
    p = model.forward_encoded(enc, chunk).view(N_rays, N_samples)
    nz = _noise(z_vals.shape, z_vals.device, noise_std, noise)
    ld = rays.shape[1]
    rl = ops.COMP_RANGE_LOSS if (_extra is not None and noise_std == 0) else 0
    if use_child_nerf_loss == 1:
        divide = use_child_nerf_divide == 1
        w, depth, fl, dl, free_r, sl1_r, _, rsl = ops.composite(p, z_vals, rays, (10, 11, ld - 1), nz, noise_std, epsilon,
                                                               ops.COMP_CHILD_LOSS | rl, divide)
        if divide:
            # render.py:106-119, :135-152: per-child means summed over sub_nerf_test_num children (column 9)
            sub = rays[:, 9]
            fl = torch.zeros(1, device=rays.device)
            dl = torch.zeros(1, device=rays.device)
            for i in range(sub_nerf_test_num):
                sel = (sub > (i + 0.5)) & (sub < (i + 1.5))
                cnt = sel.sum()
                if cnt >= 1:
                    fl = fl + free_r[sel].sum() / cnt
                    dl = dl + 1 / cnt * 0.1 * sl1_r[sel].mean()
    else:
        out = ops.composite(p, z_vals, rays if rl else None, (0, 0, ld - 1) if rl else (0, 0, 0), nz, noise_std, epsilon, rl)
        w, depth, rsl = out[0], out[1], out[7]
        fl = torch.tensor(0.0)          # CPU scalars, as in the reference (render.py:123-125,157-159)
        dl = torch.tensor(0.0)
    if rl:
        _extra["range_sl1"] = rsl
    return fl, dl, depth, w


def inference(model, embedding_xy, samples_xy, z_vals, chunk=1024 * 32, noise_std=1, epsilon=1e-10, isval=False, *,
              noise=None, _enc=None):
    """nof/render.py:166-226."""
    N_rays, N_samples = z_vals.shape
    enc = _enc if _enc is not None else _embed_samples(samples_xy, _tc(model))
    p = model.forward_encoded(enc, chunk).view(N_rays, N_samples)
    nz = _noise(z_vals.shape, z_vals.device, noise_std, noise)
    if isval is not False:
        raise NotImplementedError("inference(isval=True) returns un-normalised weights in the reference; no caller "
                                  "reaches that branch (render_rays passes isval into the epsilon slot)")
    out = ops.composite(p, z_vals, None, (0, 0, 0), nz, noise_std, float(epsilon), ops.COMP_OPACITY)
    return out[1], out[0], out[6]


def inference_0525_2(model, embedding_xy, samples_xy, z_vals, other_interest_sub_nerf_number, near_far_child,
                     chunk=1024 * 32, noise_std=1, epsilon=0, isval=False, is_fine=0, depth_inference_method=0, *,
                     _enc=None):
    """nof/render.py:229-368."""
    N_rays, N_samples = z_vals.shape
    enc = _enc if _enc is not None else _embed_samples(samples_xy, _tc(model))
    p = model.forward_encoded(enc, chunk).view(N_rays, N_samples)
    nfc = near_far_child.contiguous().to(torch.float32)
    depth, w, opacity, peak, wsum = ops.search_rows(p, z_vals, nfc, 0, 1, epsilon, depth_inference_method)
    flag = ops.search_select(other_interest_sub_nerf_number, peak, wsum)
    return depth, w, opacity, flag


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """nof/render.py:371-412."""
    if pytest:
        raise NotImplementedError("sample_pdf(pytest=True) is a dead nerf-pytorch leftover (never enabled by a caller)")
    u = None if det else _draw_u(bins.shape[0], N_samples, bins.device)
    return ops.sample_pdf(bins, weights, N_samples, u=u, det=det)


# ------------------------------------------------------------------------------------------------------- render_*


def _two_pass(model, model_fine, rays, n_a, n_b, N_importance, use_disp, perturb, U, u, near_col, far_col, head):
    """Shared skeleton of the four render_* entry points: coarse sample+encode -> head -> resample+encode -> head."""
    tc = _tc(model)
    lazy, lazy_f = _lazy(model), _lazy(model_fine)
    z, enc = ops.sample_encode_coarse(rays, n_a, n_b, near_col, far_col, 10, 11, use_disp, float(perturb), U, not lazy, tc)
    out_c = head(model, ops.LazyEnc(rays, z) if lazy else enc, z, 0)
    w = out_c["w"]
    det = (perturb == 0.)
    if not det and u is None:
        u = _draw_u(rays.shape[0], N_importance, rays.device)
    zf, encf = ops.sample_encode_fine(rays, z, w, N_importance, u, det, not lazy_f, _tc(model_fine))
    out_f = head(model_fine, ops.LazyEnc(rays, zf) if lazy_f else encf, zf, 1)
    return z, zf, out_c, out_f


def render_rays_train(model, model_fine, embedding_xy, rays, sub_nerf_test_num=4, N_samples=64, N_importance=128,
                      use_disp=False, perturb=0, noise_std=1, chunk=1024 * 3, isval=False, issegmentated=0,
                      childnerf_ratio=0.5, use_child_nerf_divide=0, use_child_nerf_loss=0, *, U=None, u=None,
                      noise=None, noise_fine=None):
    """nof/render.py:416-482."""
    n_a, n_b = _segments(N_samples, issegmentated, childnerf_ratio)
    noises = (noise, noise_fine)

    def head(net, enc, z, which):
        extra = {}
        fl, dl, depth, w = inference_train(net, embedding_xy, None, rays, z, None, None, None, None, chunk, noise_std,
                                           1e-10, isval, sub_nerf_test_num, use_child_nerf_divide, use_child_nerf_loss,
                                           noise=noises[which], _enc=enc, _extra=extra)
        return {"fl": fl, "dl": dl, "depth": depth, "w": w, "rsl": extra.get("range_sl1")}

    _, _, c, f = _two_pass(model, model_fine, rays, n_a, n_b, N_importance, False, perturb, U, u, 6, 7, head)
    out = {'child_free_loss_fine': f["fl"], 'child_depth_loss_fine': f["dl"], "depth_fine": f["depth"],
           'child_free_loss': c["fl"], 'child_depth_loss': c["dl"], 'depth': c["depth"]}
    if c["rsl"] is not None:
        # beyond the reference's six keys: SmoothL1Loss(mean)(10 depth, 10 rays[:, -1]) of both passes, fused into K4
        # (train_kitti.py:145-146; `batch['ranges']` IS rays[:, 14], ipb2dmapping.py:447-453)
        out["range_sl1"], out["range_sl1_fine"] = c["rsl"], f["rsl"]
    return out


def render_rays_val(model, model_fine, embedding_xy, rays, sub_nerf_test_num=4, N_samples=64, N_importance=128,
                    use_disp=False, perturb=0, noise_std=1, chunk=1024 * 3, isval=False, *, U=None, u=None,
                    noise=None, noise_fine=None):
    """nof/render.py:485-536."""
    noises = (noise, noise_fine)

    def head(net, enc, z, which):
        depth, w = inference_val(net, embedding_xy, None, rays, z, None, None, None, None, chunk, noise_std, 1e-10,
                                 isval, sub_nerf_test_num, noise=noises[which], _enc=enc)
        return {"depth": depth, "w": w}

    _, _, c, f = _two_pass(model, model_fine, rays, N_samples, 0, N_importance, False, perturb, U, u, 6, 7, head)
    return {"depth_fine": f["depth"], 'depth': c["depth"]}


def render_rays(model, model_fine, embedding_xy, rays, N_samples=64, N_importance=128, use_disp=False, perturb=0,
                noise_std=1, chunk=1024 * 3, isval=False, *, U=None, u=None, noise=None, noise_fine=None):
    """nof/render.py:538-611 (legacy API).  The reference passes `isval` positionally into inference()'s `epsilon`
    slot (:585 vs :166-167): epsilon = float(isval) and the weights are always normalised."""
    noises = (noise, noise_fine)
    eps = float(isval)

    def head(net, enc, z, which):
        depth, w, opacity = inference(net, embedding_xy, None, z, chunk, noise_std, eps, False, noise=noises[which],
                                      _enc=enc)
        return {"depth": depth, "w": w, "opacity": opacity}

    _, zf, c, f = _two_pass(model, model_fine, rays, N_samples, 0, N_importance, use_disp, perturb, U, u, 6, 7, head)
    weights = f["w"]
    weights_mask = weights.argsort(dim=-1, descending=True).eq(weights.shape[1] - 1)     # render.py:598-600
    return {'depth_fine': f["depth"], 'weights': weights, 'opacity': c["opacity"], 'z_vals': zf, "depth": c["depth"],
            "depth2": zf[weights_mask], "opacity_fine": f["opacity"]}


GROUP_RAYS = True       # depth inference: evaluate samples once per physical ray instead of once per candidate row


def render_rays_view_0525_2_2(model, model_fine, embedding_xy, rays, other_interest_sub_nerf_number, N_samples=64,
                              N_importance=128, use_disp=False, perturb=0, noise_std=1, chunk=1024 * 3, isval=False,
                              depth_inference_method=0, *, U=None, u=None, plan=None):
    """nof/render.py:614-699: z uniform over the PARENT segment (cols 9,10), child interval in cols 6,7.

    All candidate rows of a group share origin, direction and the parent segment (eval_kitti_render.py:379-390), so with
    perturb == 0 their samples, occupancies and weights are identical; the reference evaluates them once per ROW, this
    path once per PHYSICAL RAY (2.9 rows per ray on the shipped frames) whenever that holds for the given
    rows (checked on the device, ops.GroupPlan) -- the per-row quantities (child masks, peak test, in-child sum, depth,
    flags) are computed per row from the ray's samples.  Results are bit-identical to the per-row evaluation."""
    if GROUP_RAYS and perturb == 0 and U is None and u is None and rays.shape[0] > 0:
        if plan is None:
            plan = ops.GroupPlan(rays, other_interest_sub_nerf_number)
        if plan.uniform and plan.G < rays.shape[0]:
            return _view_grouped(model, model_fine, rays, plan, N_samples, N_importance, chunk, depth_inference_method)
    nfc = rays[:, 6:8]

    def head(net, enc, z, which):
        depth, w, opacity, flag = inference_0525_2(net, embedding_xy, None, z, other_interest_sub_nerf_number, nfc,
                                                   chunk, noise_std, 1e-10, isval, which, depth_inference_method,
                                                   _enc=enc)
        return {"depth": depth, "w": w, "opacity": opacity, "flag": flag}

    _, zf, c, f = _two_pass(model, model_fine, rays, N_samples, 0, N_importance, False, perturb, U, u, 9, 10, head)
    return {'depth_fine': f["depth"], 'weights': f["w"], 'opacity': c["opacity"], 'z_vals': zf, "depth": c["depth"],
            "opacity_fine": f["opacity"], "points_inference_fine": ops.points(rays, f["depth"]),
            "points_inference": ops.points(rays, c["depth"]), "rays_effective_flag": c["flag"],
            "rays_effective_flag_fine": f["flag"]}


def _view_grouped(model, model_fine, rays, plan, N_samples, N_importance, chunk, method):
    """render_rays_view_0525_2_2 with the samples evaluated once per physical ray (same result dict, per candidate row)."""
    rays = rays.contiguous()
    hr = rays.index_select(0, plan.head_rows)                       # (G,13) one row per physical ray
    G = plan.G
    lazy, lazy_f = _lazy(model), _lazy(model_fine)
    z, enc = ops.sample_encode_coarse(hr, N_samples, 0, 9, 10, 10, 11, False, 0.0, None, not lazy, _tc(model))
    p = model.forward_encoded(ops.LazyEnc(hr, z) if lazy else enc, chunk).view(G, N_samples)
    depth_c, w, op_c, peak, wsum = ops.search_rows(p, z, rays, 6, 7, 1e-10, method, row_ray=plan.row_ray)
    flag_c = ops.search_select(plan.other, peak, wsum)
    zf, encf = ops.sample_encode_fine(hr, z, w, N_importance, None, True, not lazy_f, _tc(model_fine))
    pf = model_fine.forward_encoded(ops.LazyEnc(hr, zf) if lazy_f else encf, chunk).view(G, N_samples + N_importance)
    depth_f, wf, op_f, peak, wsum = ops.search_rows(pf, zf, rays, 6, 7, 1e-10, method, row_ray=plan.row_ray)
    flag_f = ops.search_select(plan.other, peak, wsum)
    rr = plan.row_ray.to(torch.int64)
    return {'depth_fine': depth_f, 'weights': wf.index_select(0, rr), 'opacity': op_c, 'z_vals': zf.index_select(0, rr),
            "depth": depth_c, "opacity_fine": op_f, "points_inference_fine": ops.points(rays, depth_f),
            "points_inference": ops.points(rays, depth_c), "rays_effective_flag": flag_c,
            "rays_effective_flag_fine": flag_f}
