"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU) of PC-NeRF's ray-rendering hot path.

This file is the *oracle* (checker) for the CUDA path in `pcnerf_b200/`.  It is imported only by
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`.
The product package never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against fixtures in
`tests/golden/*.npz` that were produced by executing the reference itself (imported from
/root/reference through `oracle/ref_shim.py`) on seeded synthetic inputs -- see `oracle/make_golden.py`.

All `file:line` citations are relative to the reference tree (biter0088/pc-nerf).
The restatement is vectorised over rays (the reference loops per ray in Python); arithmetic order,
dtypes (fp32 for the renderer, fp64 for the AABB stage) and comparison operators follow the cited lines.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# AABB stage (fp64, numpy)
# --------------------------------------------------------------------------------------------------


def compute_far_bound(ray_o, ray_d, x_max, x_min, y_max, y_min, z_max, z_min):
    """nof/dataset/ipb2dmapping.py:36-77.  Vectorised over rays: ray_o (3,) or (N,3), ray_d (N,3).
    Returns (N,) float64; +inf where the reference returns None."""
    ray_d = np.asarray(ray_d, dtype=np.float64).reshape(-1, 3)
    ray_o = np.broadcast_to(np.asarray(ray_o, dtype=np.float64).reshape(-1, 3), ray_d.shape)
    planes = np.array([[x_max, x_min], [y_max, y_min], [z_max, z_min]], dtype=np.float64)
    ts = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax in range(3):
            for j in range(2):
                t = (planes[ax, j] - ray_o[:, ax]) / ray_d[:, ax]
                t = np.where(ray_d[:, ax] != 0, t, np.inf)
                t = np.where(t < 0, np.inf, t)
                ts.append(t)
    return np.min(np.stack(ts, 0), axis=0)


def _plane_hits(p, d, p_min, p_max):
    """Shared body of compute_far_bound0406/0606/0429 (ipb2dmapping.py:82-145, eval_kitti_render.py:170-196).
    Broadcasts p,d,p_min,p_max to (...,3).  Returns (valid (...,6) bool, dist (...,6) float64) in the
    reference's append order: axis0-min, axis0-max, axis1-min, axis1-max, axis2-min, axis2-max."""
    p, d, p_min, p_max = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (p, d, p_min, p_max)))
    valid, dist = [], []
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for i in range(3):
            for plane in (p_min[..., i], p_max[..., i]):
                cond = d[..., i] * (plane - p[..., i]) > 0
                distance = (plane - p[..., i]) / d[..., i]
                p_end = p + distance[..., None] * d
                count = np.zeros(cond.shape, dtype=np.int64)
                for k in range(3):
                    if k == i:
                        continue
                    count += ((p_end[..., k] >= p_min[..., k]) & (p_end[..., k] <= p_max[..., k])).astype(np.int64)
                valid.append(cond & (count >= 2))
                dist.append(distance)
    return np.stack(valid, -1), np.stack(dist, -1)


def _first_two(valid, dist):
    order = np.argsort(~valid, axis=-1, kind="stable")
    d_sorted = np.take_along_axis(dist, order, -1)
    return d_sorted[..., 0], d_sorted[..., 1]


def compute_far_bound0429(p, d, p_min, p_max):
    """eval_kitti_render.py:170-211: exactly two valid hits required.  Returns (flag, near, far)."""
    valid, dist = _plane_hits(p, d, p_min, p_max)
    n = valid.sum(-1)
    a, b = _first_two(valid, dist)
    flag = n == 2
    near = np.where(flag, np.minimum(a, b), 0.0)
    far = np.where(flag, np.maximum(a, b), 0.0)
    return flag, near, far


def compute_far_bound0406(p, d, p_min, p_max):
    """ipb2dmapping.py:82-114: first two valid hits (reference raises IndexError with fewer than two;
    here such rows come back as NaN)."""
    valid, dist = _plane_hits(p, d, p_min, p_max)
    n = valid.sum(-1)
    a, b = _first_two(valid, dist)
    ok = n >= 2
    near = np.where(ok, np.minimum(a, b), np.nan)
    far = np.where(ok, np.maximum(a, b), np.nan)
    return near, far


def compute_far_bound0606(p, d, p_min, p_max):
    """ipb2dmapping.py:119-172: 0 hits -> (False,0,0); 1 -> (d,d); 2 -> sorted; >2 -> (min,max)."""
    valid, dist = _plane_hits(p, d, p_min, p_max)
    n = valid.sum(-1)
    dm = np.where(valid, dist, np.inf).min(-1)
    dM = np.where(valid, dist, -np.inf).max(-1)
    flag = n > 0
    return flag, np.where(flag, dm, 0.0), np.where(flag, dM, 0.0)


def find_aabb_box(centres, aabb_list, query_points, k=10):
    """ipb2dmapping.py:174-197, batched over query points.  The reference builds an sklearn KDTree per
    query and walks the k nearest centres (ascending euclidean distance) returning the first whose box
    contains the point.  Equivalent statement used here: the containing box with the smallest centre
    distance, provided fewer than k centres are strictly closer.  Returns (inside (Q,) bool, idx (Q,) int64,
    -1 where outside).  Squared distances are accumulated x,y,z in fp64 like sklearn's reduced distance."""
    centres = np.asarray(centres, dtype=np.float64)
    aabb_list = np.asarray(aabb_list, dtype=np.float64)
    q = np.asarray(query_points, dtype=np.float64).reshape(-1, 3)
    K = centres.shape[0]
    if K < k:
        raise ValueError("k must be less than or equal to the number of training points")
    inside = np.zeros(q.shape[0], dtype=bool)
    idx = -np.ones(q.shape[0], dtype=np.int64)
    for s in range(0, q.shape[0], 4096):
        qq = q[s:s + 4096]
        diff = qq[:, None, :] - centres[None, :, :]
        rd = diff[..., 0] * diff[..., 0]
        rd = rd + diff[..., 1] * diff[..., 1]
        rd = rd + diff[..., 2] * diff[..., 2]
        contains = np.ones(rd.shape, dtype=bool)
        for a in range(3):
            contains &= (qq[:, None, a] >= aabb_list[None, :, a]) & (qq[:, None, a] <= aabb_list[None, :, 3 + a])
        rd_c = np.where(contains, rd, np.inf)
        best = rd_c.argmin(1)
        best_rd = rd_c[np.arange(qq.shape[0]), best]
        rank = (rd < best_rd[:, None]).sum(1)
        ok = np.isfinite(best_rd) & (rank < k)
        inside[s:s + 4096] = ok
        idx[s:s + 4096] = np.where(ok, best, -1)
    return inside, idx


def pack_train_rays(origin, points, centres, child_bounds, child_bounds_bigger, parent_box, surface_expand,
                    variant="maicity"):
    """Per-point loop body of ipb2dmapping.py:367-397 (MaiCity, compute_far_bound0406) /
    :736-768 (KITTI, compute_far_bound0606 + drop on no-intersection), followed by the 15-column packing of
    :447-452 / :819-824.  origin (3,), points (P,3) float64.  parent_box = (x_min,x_max,y_min,y_max,z_min,z_max).
    Returns rays (N,15) float32 (torch.Tensor(np.float64) rounding) and the kept point indices."""
    origin = np.asarray(origin, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    vec = points - origin
    dist_vec = np.linalg.norm(vec, axis=1)
    dir_vec = vec / dist_vec[:, None]
    return pack_train_rays_from_dirs(origin, dir_vec, dist_vec, points, centres, child_bounds,
                                     child_bounds_bigger, parent_box, surface_expand, variant)


def pack_train_rays_from_dirs(origin, dir_vec, dist_vec, points, centres, child_bounds, child_bounds_bigger,
                              parent_box, surface_expand, variant="maicity"):
    inside, idx = find_aabb_box(centres, child_bounds, points)
    keep = inside.copy()
    bb = np.asarray(child_bounds_bigger, dtype=np.float64)[np.where(inside, idx, 0)]
    if variant == "maicity":
        near, far = compute_far_bound0406(origin, dir_vec, bb[:, :3], bb[:, 3:6])
    else:
        flag, near, far = compute_far_bound0606(origin, dir_vec, bb[:, :3], bb[:, 3:6])
        keep &= flag
    near = near - surface_expand
    far = far + surface_expand
    near_pt = dist_vec - surface_expand
    x_min, x_max, y_min, y_max, z_min, z_max = parent_box
    far_parent = compute_far_bound(origin, dir_vec, x_max, x_min, y_max, y_min, z_max, z_min)
    far_parent = np.where(far_parent < far, far, far_parent)
    n = int(keep.sum())
    rays = np.zeros((n, 15), dtype=np.float64)
    rays[:, 0:3] = origin
    rays[:, 3:6] = dir_vec[keep]
    rays[:, 6] = 0.0
    rays[:, 7] = far_parent[keep]
    rays[:, 8] = 3
    rays[:, 9] = idx[keep] + 1
    rays[:, 10] = near[keep]
    rays[:, 11] = far[keep]
    rays[:, 12] = near_pt[keep]
    rays[:, 13] = far[keep]          # ipb2dmapping.py:443 concatenates far_bound, not far_bound_point
    rays[:, 14] = dist_vec[keep]
    return rays.astype(np.float32), np.nonzero(keep)[0]


def ray_aabb_distances(ray_origin, ray_dirs, aabb_min, aabb_max):
    """eval_kitti_render.py:213-235 (already vectorised in the reference)."""
    o = np.asarray(ray_origin, dtype=np.float64)
    dirs = np.asarray(ray_dirs, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        tmins, tmaxs = [], []
        for a in range(3):
            t1 = (aabb_min[a] - o[a]) / dirs[:, a]
            t2 = (aabb_max[a] - o[a]) / dirs[:, a]
            tmins.append(np.minimum(t1, t2))
            tmaxs.append(np.maximum(t1, t2))
        tmin = np.max(np.vstack(tmins), axis=0)
        tmax = np.min(np.vstack(tmaxs), axis=0)
    return np.where(tmax >= tmin, tmax, np.inf)


def distance_to_ray(ray_origin, ray_dir, points):
    """eval_kitti_render.py:237-244 (one ray, all centres)."""
    v = points - np.asarray(ray_origin, dtype=np.float64)
    dist = np.sqrt(np.sum(v ** 2, axis=1))
    cos_angle = np.sum(v * ray_dir, axis=1) / dist
    with np.errstate(invalid="ignore"):
        sin_angle = np.sqrt(1 - cos_angle ** 2)
    return dist * sin_angle


def build_candidate_groups(origin, dir_vec, dist_vec, child_bounds, child_bounds_larger, parent_min, parent_max,
                           depth_inference_method=2, grow_step=0.005, prefilter=0.65):
    """Per-ray loop body of eval_kitti_render.py:353-461 (MaiCity, grow_step 0.005) / :681-803 (KITTI, 0.05),
    followed by the column drop of :520-522.  Returns (rays (N',13) float32, ranges (N',1) float32,
    other_interest_sub_nerf_number (N',1) int64, kept physical-ray indices)."""
    origin = np.asarray(origin, dtype=np.float64)
    dir_vec = np.asarray(dir_vec, dtype=np.float64)
    sb = np.asarray(child_bounds, dtype=np.float64)
    sbl = np.asarray(child_bounds_larger, dtype=np.float64)
    parent_far = ray_aabb_distances(origin, dir_vec, parent_min, parent_max)
    center = (sb[:, :3] + sb[:, 3:]) / 2
    rows, others, kept = [], [], []
    for i in range(dir_vec.shape[0]):
        d = dir_vec[i]
        with np.errstate(invalid="ignore", divide="ignore"):
            dtr = distance_to_ray(origin, d, center)
            filt = sbl[dtr <= prefilter].copy()
        cand = []
        hit = False

        def scan(boxes):
            out = []
            if boxes.shape[0] == 0:
                return out
            flag, nb, fb = compute_far_bound0429(origin, d, boxes[:, :3], boxes[:, 3:6])
            for k in range(boxes.shape[0]):
                if flag[k]:
                    if depth_inference_method == 1:
                        out.append((0.0, parent_far[i]))
                        break
                    out.append((nb[k], fb[k]))
            return out

        cand = scan(filt)
        hit = len(cand) > 0
        extend_iter = 0
        dropped = False
        while not hit:
            if extend_iter > 0.5:
                dropped = True
                break
            extend_iter = extend_iter + grow_step
            filt[:, :3] = filt[:, :3] - extend_iter
            filt[:, 3:6] = filt[:, 3:6] + extend_iter
            cand = scan(filt)
            hit = len(cand) > 0
        if dropped or len(cand) == 0:
            continue
        n = len(cand)
        nears = np.array([c[0] for c in cand], dtype=np.float64)
        order = np.argsort(nears)
        for r, j in enumerate(order):
            # 14 columns of :379-390,:439-447 with range (col 9) removed by :520-522
            rows.append([origin[0], origin[1], origin[2], d[0], d[1], d[2], cand[j][0], cand[j][1], 3.0,
                         0.0, parent_far[i], float(r + 1), float(n - 1) if r == 0 else -1.0])
            others.append(n - 1 if r == 0 else 0)
        kept.append((i, n))
    rays = np.asarray(rows, dtype=np.float64).reshape(-1, 13).astype(np.float32)
    ranges = np.concatenate([np.full((n,), dist_vec[i]) for i, n in kept]).astype(np.float32).reshape(-1, 1) \
        if kept else np.zeros((0, 1), np.float32)
    return rays, ranges, np.asarray(others, dtype=np.int64).reshape(-1, 1), np.asarray([i for i, _ in kept])


# --------------------------------------------------------------------------------------------------
# Model (fp32, torch CPU)
# --------------------------------------------------------------------------------------------------


def embedding(x, N_freq=10):
    """nof/networks/models.py:27-41: [x, sin(2^k x), cos(2^k x)]_{k<N_freq}."""
    out = [x]
    for k in range(N_freq):
        f = float(2 ** k)
        out.append(torch.sin(f * x))
        out.append(torch.cos(f * x))
    return torch.cat(out, -1)


LAYER1_LIN = ("layer1.0", "layer1.3", "layer1.6", "layer1.9")
LAYER1_BN = ("layer1.1", "layer1.4", "layer1.7", "layer1.10")
LAYER2_LIN = ("layer2.0", "layer2.2", "layer2.4", "layer2.6")
LAYER2_BN = ("layer2.1", "layer2.3", "layer2.5", "layer2.7")


def nof_forward(sd, x, training, momentum=0.1, eps=1e-5, use_skip=True):
    """nof/networks/models.py:183-203 as actually constructed: every LeakyReLU has
    negative_slope == True == 1.0 (identity, models.py:152) and the layer2 activations were appended to
    layer1 (models.py:172).  `sd` is a dict of tensors keyed like the reference state_dict; BN running
    buffers are updated in place when training (one update per call = per chunk)."""

    def bn(h, name):
        return F.batch_norm(h, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                            sd[name + ".bias"], training, momentum, eps)

    h = x
    for lin, b in zip(LAYER1_LIN, LAYER1_BN):
        h = bn(F.linear(h, sd[lin + ".weight"], sd[lin + ".bias"]), b)
        if training:
            sd[b + ".num_batches_tracked"] += 1
    if use_skip:
        h = torch.cat([x, h], dim=1)
    for lin, b in zip(LAYER2_LIN, LAYER2_BN):
        h = bn(F.linear(h, sd[lin + ".weight"], sd[lin + ".bias"]), b)
        if training:
            sd[b + ".num_batches_tracked"] += 1
    return torch.sigmoid(F.linear(h, sd["occ_out.0.weight"], sd["occ_out.0.bias"]))


def init_state_dict(seed, feature_size=256, in_ch=63, randomize_bn=True):
    """Default torch init of the reference modules restated functionally (kaiming_uniform(a=sqrt5) for
    Linear = U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both W and b), plus non-trivial BN state."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, fin, fout):
        bound = 1.0 / math.sqrt(fin)
        sd[name + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound

    def bnp(name, f):
        if randomize_bn:
            sd[name + ".weight"] = torch.rand(f, generator=g) + 0.5
            sd[name + ".bias"] = torch.rand(f, generator=g) - 0.5
            sd[name + ".running_mean"] = torch.rand(f, generator=g) - 0.5
            sd[name + ".running_var"] = torch.rand(f, generator=g) + 0.5
        else:
            sd[name + ".weight"] = torch.ones(f)
            sd[name + ".bias"] = torch.zeros(f)
            sd[name + ".running_mean"] = torch.zeros(f)
            sd[name + ".running_var"] = torch.ones(f)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for i, (l, b) in enumerate(zip(LAYER1_LIN, LAYER1_BN)):
        lin(l, in_ch if i == 0 else feature_size, feature_size)
        bnp(b, feature_size)
    for i, (l, b) in enumerate(zip(LAYER2_LIN, LAYER2_BN)):
        lin(l, in_ch + feature_size if i == 0 else feature_size, feature_size)
        bnp(b, feature_size)
    lin("occ_out.0", feature_size, 1)
    return sd


def param_names():
    names = []
    for l, b in zip(LAYER1_LIN, LAYER1_BN):
        names += [l + ".weight", l + ".bias", b + ".weight", b + ".bias"]
    for l, b in zip(LAYER2_LIN, LAYER2_BN):
        names += [l + ".weight", l + ".bias", b + ".weight", b + ".bias"]
    names += ["occ_out.0.weight", "occ_out.0.bias"]
    return names


def mlp_chunks(sd, samples, chunk, training):
    """nof/render.py:47-49: embed and evaluate in chunks of `chunk` rows (one BN batch per chunk)."""
    outs = []
    B = samples.shape[0]
    for i in range(0, B, chunk):
        outs.append(nof_forward(sd, embedding(samples[i:i + chunk]), training))
    return torch.cat(outs, 0)


# --------------------------------------------------------------------------------------------------
# Renderer (fp32, torch CPU)
# --------------------------------------------------------------------------------------------------


def lerp_z(near, far, n):
    """nof/render.py:430-432: near*(1-s)+far*s with s = torch.linspace(0,1,n)."""
    s = torch.linspace(0, 1, n).expand(near.shape[0], n)
    return near * (1 - s) + far * s


def sample_z(rays, N_samples, issegmentated, childnerf_ratio, perturb, U=None, near_col=6, far_col=7):
    """nof/render.py:429-454.  U = pre-drawn torch.rand(N, S) (reference draws it on rays.device)."""
    near, far = rays[:, near_col].view(-1, 1), rays[:, far_col].view(-1, 1)
    if issegmentated == 0:
        z = lerp_z(near, far, N_samples)
    else:
        n_parent = int(N_samples * (1 - childnerf_ratio))
        n_child = N_samples - n_parent
        zp = lerp_z(near, far, n_parent)
        zc = lerp_z(rays[:, 10].view(-1, 1), rays[:, 11].view(-1, 1), n_child)
        z, _ = torch.sort(torch.cat([zp, zc], -1), -1)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * U)
    return z


def composite(p, noise=None, noise_std=0.0, epsilon=1e-10, normalize=True):
    """nof/render.py:51-61."""
    free = 1 - p
    shift = torch.cat([torch.ones_like(free[:, :1]), free], -1)
    T = torch.cumprod(shift, -1)[:, :-1]
    w = T * p
    if noise is not None:
        w = w + noise * noise_std
    if normalize:
        w = w / (torch.sum(w, -1).reshape(-1, 1) + epsilon)
    return w


def child_mask(z, near_far_child, gamma0, strict, step=0.01, max_iter=200000):
    """nof/render.py:77-84 (gamma0=0.0, closed), :91-97 (gamma0=2, closed), :252-263 (gamma0=0.01, strict).
    The python float `expand_threshold` accumulates in double; `interval[k] -/+ expand_threshold` is an fp32
    tensor-scalar op (scalar rounded to fp32, fp32 subtract)."""
    N = z.shape[0]
    lo0, hi0 = near_far_child[:, 0:1], near_far_child[:, 1:2]
    mask = torch.zeros_like(z, dtype=torch.bool)
    bound = torch.zeros(N, 2)
    todo = torch.ones(N, dtype=torch.bool)
    g = float(gamma0)
    for _ in range(max_iter):
        lo, hi = lo0 - g, hi0 + g
        m = ((lo < z) & (z < hi)) if strict else ((lo <= z) & (z <= hi))
        mask[todo] = m[todo]
        bound[todo, 0] = lo[todo, 0]
        bound[todo, 1] = hi[todo, 0]
        todo = todo & (mask.sum(-1) == 0)
        if not bool(todo.any()):
            break
        g = g + step
    return mask, bound


def smooth_l1_mean(a, b):
    d = (a - b).abs()
    return torch.where(d < 1.0, 0.5 * d * d, d - 0.5).mean()


def train_head(p, z, rays, noise=None, noise_std=0.0, epsilon=1e-10, use_child_nerf_loss=1,
               use_child_nerf_divide=0, sub_nerf_test_num=4):
    """nof/render.py:51-163 after the MLP: weights, masks, child free / depth losses, depth.
    With float64 p / z / rays (exact images of the fp32 inputs) the masks are still decided in fp32, everything else runs
    in double: the gradient "truth" the fp32 kernels and fp32 autograd are both measured against in tests."""
    near_far_child = rays[:, 10:12]
    range_readings = rays[:, -1]
    N, S = z.shape
    w = composite(p, noise, noise_std, epsilon)
    if use_child_nerf_loss == 1:
        m0, _ = child_mask(z.float(), near_far_child.float(), 0.0, strict=False)
        m2, _ = child_mask(z.float(), near_far_child.float(), 2, strict=False)
        w_non = w * (~m0).to(w.dtype)
        w_child = w * m2.to(w.dtype)
        z_child = z * m2.to(w.dtype)
        w_child = w_child / (torch.sum(w_child, -1).reshape(-1, 1) + epsilon)
        d_hat = torch.sum(w_child * z_child, -1)
        if use_child_nerf_divide == 1:
            sub = rays[:, 9]
            free_loss = torch.zeros(1)
            depth_loss = torch.zeros(1)
            for i in range(sub_nerf_test_num):
                sel = (sub > (i + 0.5)) & (sub < (i + 1.5))
                cnt = sel.sum()
                if cnt >= 1:
                    free_loss = free_loss + torch.sum(torch.square(w_non[sel])) / cnt
                    depth_loss = depth_loss + 1 / cnt * 0.1 * smooth_l1_mean(
                        10 * d_hat[sel].reshape(-1, 1).squeeze(), 10 * range_readings[sel].reshape(-1, 1).squeeze())
        else:
            free_loss = torch.sum(torch.square(w_non)) / N
            depth_loss = 1 / N * 0.1 * smooth_l1_mean(10 * d_hat, 10 * range_readings)
    else:
        free_loss = torch.tensor(0.0)
        depth_loss = torch.tensor(0.0)
    depth = torch.sum(w * z, -1)
    return free_loss, depth_loss, depth, w


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """nof/render.py:371-412.  `u` replaces the CPU-generator torch.rand draw when det is False."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0., 1., steps=N_samples).expand(list(cdf.shape[:-1]) + [N_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bins_b, bins_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b)


def fine_z(z, w, N_importance, det, u=None):
    """nof/render.py:463-467."""
    mid = .5 * (z[..., 1:] + z[..., :-1])
    zs = sample_pdf(mid, w[..., 1:-1], N_importance, det=det, u=u).detach()
    zf, _ = torch.sort(torch.cat([z, zs], -1), -1)
    return zf


def render_rays_train(sd_c, sd_f, rays, N_samples=64, N_importance=128, perturb=0, noise_std=1, chunk=1024 * 3,
                      issegmentated=0, childnerf_ratio=0.5, use_child_nerf_divide=0, use_child_nerf_loss=0,
                      sub_nerf_test_num=4, U=None, u_fine=None, noise_c=None, noise_f=None, training=True):
    """nof/render.py:416-482.  sd_c/sd_f: state dicts (leaf tensors may require grad)."""
    o, d = rays[:, :3], rays[:, 3:6]
    z = sample_z(rays, N_samples, issegmentated, childnerf_ratio, perturb, U)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    p = mlp_chunks(sd_c, pts.view(-1, 3), chunk, training).view(z.shape)
    fl, dl, depth, w = train_head(p, z, rays, noise_c, noise_std, 1e-10, use_child_nerf_loss,
                                  use_child_nerf_divide, sub_nerf_test_num)
    zf = fine_z(z, w, N_importance, det=(perturb == 0.), u=u_fine)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * zf.unsqueeze(2)
    pf = mlp_chunks(sd_f, pts.view(-1, 3), chunk, training).view(zf.shape)
    flf, dlf, depthf, wf = train_head(pf, zf, rays, noise_f, noise_std, 1e-10, use_child_nerf_loss,
                                      use_child_nerf_divide, sub_nerf_test_num)
    return {"child_free_loss_fine": flf, "child_depth_loss_fine": dlf, "depth_fine": depthf,
            "child_free_loss": fl, "child_depth_loss": dl, "depth": depth,
            "_z": z, "_z_fine": zf, "_w": w, "_w_fine": wf}


def render_rays_val(sd_c, sd_f, rays, N_samples=64, N_importance=128, perturb=0, noise_std=1, chunk=1024 * 3,
                    U=None, u_fine=None, noise_c=None, noise_f=None, training=False):
    """nof/render.py:485-536."""
    o, d = rays[:, :3], rays[:, 3:6]
    z = sample_z(rays, N_samples, 0, 0.5, perturb, U)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    p = mlp_chunks(sd_c, pts.view(-1, 3), chunk, training).view(z.shape)
    w = composite(p, noise_c, noise_std, 1e-10)
    depth = torch.sum(w * z, -1)
    zf = fine_z(z, w, N_importance, det=(perturb == 0.), u=u_fine)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * zf.unsqueeze(2)
    pf = mlp_chunks(sd_f, pts.view(-1, 3), chunk, training).view(zf.shape)
    wf = composite(pf, noise_f, noise_std, 1e-10)
    return {"depth_fine": torch.sum(wf * zf, -1), "depth": depth}


def opacity_reg(p):
    """nof/render.py:224."""
    return torch.mean(torch.log(0.1 + p) + torch.log(0.1 + (1 - p)) + 2.20727)


def render_rays(sd_c, sd_f, rays, N_samples=64, N_importance=128, use_disp=False, perturb=0, noise_std=1,
                chunk=1024 * 3, isval=False, U=None, u_fine=None, noise_c=None, noise_f=None, training=False):
    """nof/render.py:538-611 (legacy API).  `isval` lands in inference()'s epsilon slot (:585 vs :166-167):
    epsilon = float(isval) and the weights are always normalised."""
    o, d = rays[:, :3], rays[:, 3:6]
    near, far = rays[:, 6].view(-1, 1), rays[:, 7].view(-1, 1)
    s = torch.linspace(0, 1, N_samples).expand(rays.shape[0], N_samples)
    z = 1 / (1 / near * (1 - s) + 1 / far * s) if use_disp else near * (1 - s) + far * s
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * U)
    eps = float(isval)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    p = mlp_chunks(sd_c, pts.view(-1, 3), chunk, training).view(z.shape)
    w = composite(p, noise_c, noise_std, eps)
    depth = torch.sum(w * z, -1)
    opacity = opacity_reg(p)
    zf = fine_z(z, w, N_importance, det=(perturb == 0.), u=u_fine)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * zf.unsqueeze(2)
    pf = mlp_chunks(sd_f, pts.view(-1, 3), chunk, training).view(zf.shape)
    wf = composite(pf, noise_f, noise_std, eps)
    depth_fine = torch.sum(wf * zf, -1)
    wmask = wf.argsort(dim=-1, descending=True).eq(wf.shape[1] - 1)
    return {"depth_fine": depth_fine, "weights": wf, "opacity": opacity, "z_vals": zf, "depth": depth,
            "depth2": zf[wmask], "opacity_fine": opacity_reg(pf)}


# ---- two-step depth-inference search -----------------------------------------------------------------


def gaussian_kernel1d(sigma=5.0, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0): radius=int(truncate*sigma+0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def gaussian_filter_reflect(w, sigma=5.0):
    """scipy.ndimage.gaussian_filter(w_i, sigma) for 1-D fp32 rows (nof/render.py:303-307), restated:
    NI_Correlate1D symmetric branch -- double accumulation  tmp = x[l]*k[c]; for j=-r..-1: tmp += (x[l+j]+x[l-j])*k[j+r],
    'reflect' (half-sample symmetric) boundary, result rounded to fp32."""
    k, r = gaussian_kernel1d(sigma)
    x = np.asarray(w, dtype=np.float32).astype(np.float64)
    n = x.shape[-1]
    idx = np.arange(-r, n + r)
    per = 2 * n
    idx = np.mod(idx, per)
    idx = np.where(idx >= n, per - 1 - idx, idx)
    xe = x[..., idx]
    out = xe[..., r:r + n] * k[r]
    for j in range(-r, 0):
        out = out + (xe[..., r + j:r + j + n] + xe[..., r - j:r - j + n]) * k[j + r]
    return out.astype(np.float32)


def search_head(p, z, other, near_far_child, epsilon=1e-10, depth_inference_method=0):
    """nof/render.py:241-368 after the MLP.  Returns depth, weights, opacity, flag (N,1) bool."""
    N = z.shape[0]
    w = composite(p, None, 0.0, epsilon)
    mask_child, _ = child_mask(z, near_far_child, 0.01, strict=True)
    sm = torch.from_numpy(gaussian_filter_reflect(w.detach().numpy(), 5.0))
    max_idx = torch.argmax(sm, dim=1)
    peak_in = mask_child[torch.arange(N), max_idx].float().reshape(-1, 1)        # mask2 (:308-313)
    wsum = torch.sum(w * mask_child.float(), -1).reshape(-1, 1)                    # :314-315
    flag = torch.zeros((N, 1), dtype=torch.bool)
    i = 0
    oth = other.reshape(-1).tolist()
    pk = peak_in.reshape(-1).tolist()
    ws = wsum.reshape(-1)
    while i < N:
        if abs(oth[i] - 0) < 0.5:
            flag[i] = True
            i += 1
        elif oth[i] > 0.5:
            k = int(oth[i])
            win = i
            if not abs(pk[i] - 1) < 0.1:
                found = False
                for j in range(k):
                    if abs(pk[i + j + 1] - 1) < 0.1:
                        win = i + j + 1
                        found = True
                        break
                if not found:
                    for j in range(k):
                        if ws[i + j + 1] > ws[win]:
                            win = i + j + 1
            flag[win] = True
            i += k + 1
        else:
            i += 1
    if depth_inference_method == 2:
        wc = w * mask_child.float()
        wc = wc / (torch.sum(wc, -1).reshape(-1, 1) + epsilon)
        depth = torch.sum(wc * z, -1)
    else:
        depth = torch.sum(w * z, -1)
    return depth, w, opacity_reg(p), flag


def render_rays_view(sd_c, sd_f, rays, other, N_samples=64, N_importance=128, perturb=0, noise_std=1,
                     chunk=1024 * 3, depth_inference_method=0, U=None, u_fine=None, training=False):
    """nof/render.py:614-699 (render_rays_view_0525_2_2)."""
    o, d = rays[:, :3], rays[:, 3:6]
    nfc = rays[:, 6:8]
    z = sample_z(rays, N_samples, 0, 0.5, perturb, U, near_col=9, far_col=10)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    p = mlp_chunks(sd_c, pts.view(-1, 3), chunk, training).view(z.shape)
    depth, w, opacity, flag = search_head(p, z, other, nfc, 1e-10, depth_inference_method)
    zf = fine_z(z, w, N_importance, det=(perturb == 0.), u=u_fine)
    pts = o.unsqueeze(1) + d.unsqueeze(1) * zf.unsqueeze(2)
    pf = mlp_chunks(sd_f, pts.view(-1, 3), chunk, training).view(zf.shape)
    depth_f, wf, opacity_f, flag_f = search_head(pf, zf, other, nfc, 1e-10, depth_inference_method)
    return {"depth_fine": depth_f, "weights": wf, "opacity": opacity, "z_vals": zf, "depth": depth,
            "opacity_fine": opacity_f, "points_inference_fine": o + depth_f.unsqueeze(1) * d,
            "points_inference": o + depth.unsqueeze(1) * d, "rays_effective_flag": flag,
            "rays_effective_flag_fine": flag_f}


# ---- loss assembly (train_kitti.py:117-156) ------------------------------------------------------------


def training_loss(results, gt, lambda_loss=1.0, lambda_child_free_loss=1.0, lambda_child_depth_loss=1.0):
    """train_kitti.py:145-155, use_child_nerf_divide == 0 branch (both range terms use lambda_loss)."""
    lr_c = 1e-1 * lambda_loss * smooth_l1_mean(1e1 * results["depth"], 1e1 * gt)
    lr_f = 1e-1 * lambda_loss * smooth_l1_mean(1e1 * results["depth_fine"], 1e1 * gt)
    return lr_c + lr_f + lambda_child_free_loss * results["child_free_loss_fine"] \
        + lambda_child_free_loss * results["child_free_loss"] \
        + lambda_child_depth_loss * results["child_depth_loss_fine"] \
        + lambda_child_depth_loss * results["child_depth_loss"]


def eval_batches(rays, batch_size_set):
    """eval_kitti_render.py:979-1005 / :1111-1136: group-aligned batching.  Returns [(start, stop)]."""
    n = rays.shape[0]
    out = []
    i = 0
    while i < n:
        if i == n - 1:
            break
        if i + batch_size_set < n - 0.5 * batch_size_set:
            extra = 0
            while rays[i + batch_size_set + extra, -1] < -0.5:
                extra += 1
                if i + batch_size_set + extra == n:
                    break
            out.append((i, i + batch_size_set + extra))
            i = i + batch_size_set + extra
        else:
            out.append((i, n))
            i = n
    return out


# --------------------------------------------------------------------------------------------------
# Point-cloud metrics.  Parity note: nn_correspondance needs Open3D (KDTreeFlann, not
# installed here; no pinned version in the reference), so this row is anchored on the algorithm it runs --
# FLANN's exact kd-tree 1-NN search (SearchParams checks = -1, eps = 0) in float64 -- restated with SciPy's
# exact cKDTree, not on executed reference output.
# --------------------------------------------------------------------------------------------------


def nn_correspondance(verts1, verts2):
    """nof/criteria/pointcloud_metrics.py:5-33 -> (indices, distances) as numpy arrays."""
    from scipy.spatial import cKDTree
    verts1 = np.asarray(verts1, dtype=np.float64).reshape(-1, 3)
    verts2 = np.asarray(verts2, dtype=np.float64).reshape(-1, 3)
    if len(verts1) == 0 or len(verts2) == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    dist, idx = cKDTree(verts1).query(verts2, k=1)
    return idx, dist


def eval_pts(pts1, pts2, threshold=0.2):
    """nof/criteria/pointcloud_metrics.py:39-49."""
    _, dist1 = nn_correspondance(pts1, pts2)
    _, dist2 = nn_correspondance(pts2, pts1)
    precision = np.mean((dist1 < threshold).astype('float'))
    recall = np.mean((dist2 < threshold).astype('float'))
    fscore = 2 * precision * recall / (precision + recall)
    cd = np.mean(dist1) + np.mean(dist2)
    return cd, fscore


# ---------------------------------------------------------------------------------------------------------------
# KITTI dataset build: pose file, per-frame filtering / transform, then the packing above.
# Pinned by tests/golden/kitti_dataset.npz (oracle/make_golden_dataset.py executes the reference class itself).
# ---------------------------------------------------------------------------------------------------------------
T_VELO2CAM = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                       [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                       [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01],
                       [0, 0, 0, 1]])


def kitti_poses(pose_lines, data_start):
    """ipb2dmapping.py:566-591: every pose is (P @ T_velo2cam), re-expressed in the frame of pose data_start+1; the
    product is taken in float32 (torch.Tensor(...)).  Returns a float32 torch tensor (n,4,4)."""
    poses = []
    for row in pose_lines:
        P = np.append(np.array([float(i) for i in row.strip("\n").split(" ")]).reshape(3, 4), np.array([[0, 0, 0, 1]]), axis=0)
        poses.append(np.matmul(P, T_VELO2CAM))
    poses = np.array(poses)
    T_start_inv = torch.from_numpy(np.linalg.inv(poses[data_start + 1])).float()
    return T_start_inv @ torch.Tensor(poses)


def kitti_frame_returns(points_f32, poses, j, data_start, data_end, range_delete_x, range_delete_y, range_delete_z,
                        over_height, over_low, interest_x, interest_y):
    """ipb2dmapping.py:662-711 for frame file j+1: near-sensor box removal, 120 m range gate, height window (all on the
    float32 sensor-frame points), pose transform in float64 (float32 pose entries), interest region around any pose of the
    run (float32 arithmetic: numpy-scalar minus 0-d float32 tensor), then ray directions / ranges from the sensor position.
    Returns (world points (M,3) f64, dir (M,3) f64, dist (M,) f64, position (3,) f32 tensor)."""
    p = np.asarray(points_f32, dtype=np.float32)
    mask1 = np.logical_or.reduce((np.abs(p[:, 0]) >= range_delete_x, np.abs(p[:, 1]) >= range_delete_y,
                                  np.abs(p[:, 2]) >= range_delete_z))
    p = p[mask1]
    p = p[np.linalg.norm(p, axis=1) <= 120]
    p = p[p[:, 2] <= over_height]
    p = p[p[:, 2] >= over_low]
    pe = np.vstack((p.T, np.ones((1, p.shape[0]))))
    pe = (poses[j + 1].numpy() @ pe).T[:, :3]                  # float32 pose promoted to float64 by numpy
    px = poses[data_start + 1:data_end + 1, 0, -1]
    py = poses[data_start + 1:data_end + 1, 1, -1]
    x32 = torch.from_numpy(pe[:, 0].astype(np.float32))        # np.float64 scalar - float32 tensor -> float32 arithmetic
    y32 = torch.from_numpy(pe[:, 1].astype(np.float32))
    near = ((x32[:, None] - px[None, :]).abs() <= interest_x) & ((y32[:, None] - py[None, :]).abs() <= interest_y)
    pe = pe[near.any(1).numpy()]
    pos = poses[j + 1][:3, -1]
    vec = pe - np.array([pos[0], pos[1], pos[2]])
    dist_vec = np.linalg.norm(vec, axis=1)
    dir_vec = np.apply_along_axis(lambda x: x / np.linalg.norm(x), 1, vec) if len(vec) else vec
    return pe, dir_vec, dist_vec, pos


def kitti_child_boxes(child_clouds, extend=0.025):
    """ipb2dmapping.py:596-626: child bounds = axis-aligned bounds of each child cloud +- 0.025, centre of the raw bounds."""
    K = len(child_clouds)
    bound, centre = np.zeros((K, 6)), np.zeros((K, 3))
    for i, c in enumerate(child_clouds):
        c = np.asarray(c, dtype=np.float64)
        lo, hi = c.min(0), c.max(0)
        bound[i, :3], bound[i, 3:] = lo - extend, hi + extend
        centre[i] = (lo + hi) / 2.0
    return bound, bound.copy(), centre


def kitti_train_frames(data_start, data_end, split="train"):
    """Frame selection of ipb2dmapping.py:643-658 (frame sparsity 20 %): file numbers j+1."""
    out = []
    for j in range(data_start, data_end):
        if split == "train" and (j + 1 - 3 - data_start) % 5 != 0:
            out.append(j)
        elif split == "val" and (j + 1 - 3) % 5 == 0:
            out.append(j)
    return out


def kitti_build_rays(frames, pose_lines, child_clouds, parent_cloud, data_start, data_end, split="train", **kw):
    """kitti_dataload(re_loaddata=1) end to end.  frames: {file number: (N,3) float32}.  Returns (rays (N,15) f32, ranges)."""
    poses = kitti_poses(pose_lines, data_start)
    bound, bigger, centre = kitti_child_boxes(child_clouds)
    par = np.asarray(parent_cloud, dtype=np.float64)
    lo, hi = par.min(0), par.max(0)
    parent_box = (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
    rays = []
    for j in kitti_train_frames(data_start, data_end, split):
        pe, dir_vec, dist_vec, pos = kitti_frame_returns(frames[j + 1], poses, j, data_start, data_end,
                                                         kw["range_delete_x"], kw["range_delete_y"], kw["range_delete_z"],
                                                         kw["over_height"], kw["over_low"], kw["interest_x"], kw["interest_y"])
        r, _ = pack_train_rays_from_dirs(pos.numpy().astype(np.float64), dir_vec, dist_vec, pe, centre, bound, bigger,
                                         parent_box, kw["surface_expand"], "kitti")
        r[:, 0:3] = pos.numpy()                                 # rays_o is the float32 pose translation itself (:788)
        rays.append(r)
    rays = np.concatenate(rays) if rays else np.zeros((0, 15), np.float32)
    return rays, rays[:, 14].copy()


def maicity_build_rays(frames, pose_lines, child_clouds, data_start, data_end, split="train", **kw):
    """maicity_dataload(re_loaddata=1), ipb2dmapping.py:200-463: raw poses (no calibration / re-basing), frame file j+1 uses
    pose j, near-sensor box and `< 120 m` gates on the float32 points, pose transform in float64 (float32 pose entries),
    closed parent-box test on the transformed points, float64 sensor position, compute_far_bound0406 (every kept ray hits its
    box).  Returns (rays (N,15) f32, ranges)."""
    P = np.array([np.append(np.array([float(i) for i in r.strip("\n").split(" ")]).reshape(3, 4), np.array([[0, 0, 0, 1]]), axis=0)
                  for r in pose_lines])
    positions = P[:, :3, -1]
    poses32 = torch.Tensor(P)
    bound, bigger, centre = kitti_child_boxes(child_clouds)
    parent_box = (kw["nerf_length_min"], kw["nerf_length_max"], kw["nerf_width_min"], kw["nerf_width_max"],
                  kw["nerf_height_min"], kw["nerf_height_max"])
    rays = []
    for j in range(data_start, data_end):
        if not ((split == "train" and (j + 1 - 3 - data_start) % 5 != 0) or (split == "val" and (j + 1 - 3 - data_start) % 5 == 0)):
            continue
        p = np.asarray(frames[j + 1], dtype=np.float32)
        p = p[np.logical_or.reduce((np.abs(p[:, 0]) >= kw["range_delete_x"], np.abs(p[:, 1]) >= kw["range_delete_y"],
                                    np.abs(p[:, 2]) >= kw["range_delete_z"]))]
        p = p[np.linalg.norm(p, axis=1) < 120]
        pe = (poses32[j].numpy() @ np.vstack((p.T, np.ones((1, p.shape[0]))))).T[:, :3]
        m = (pe[:, 0] >= parent_box[0]) & (pe[:, 1] >= parent_box[2]) & (pe[:, 2] >= parent_box[4]) & \
            (pe[:, 0] <= parent_box[1]) & (pe[:, 1] <= parent_box[3]) & (pe[:, 2] <= parent_box[5])
        pe = pe[m]
        vec = pe - positions[j]
        dist_vec = np.linalg.norm(vec, axis=1)
        dir_vec = np.apply_along_axis(lambda x: x / np.linalg.norm(x), 1, vec) if len(vec) else vec
        r, _ = pack_train_rays_from_dirs(positions[j], dir_vec, dist_vec, pe, centre, bound, bigger, parent_box,
                                         kw["surface_expand"], "maicity")
        r[:, 0:3] = poses32[j][:3, -1].numpy()
        rays.append(r)
    rays = np.concatenate(rays) if rays else np.zeros((0, 15), np.float32)
    return rays, rays[:, 14].copy()
