"""Tensor-level wrappers over the C ABI (include/pcnerf_b200.h) + the autograd glue.

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every computation on the path is a
kernel of libpcnerf_b200.so.  All tensors must be CUDA tensors; there is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MlpGrads, MlpParams, check, lib

COMP_CHILD_LOSS, COMP_OPACITY, COMP_RANGE_LOSS = 1, 2, 4
_f64p = ctypes.POINTER(ctypes.c_double)

def launch_count(reset=False):
    """Kernels launched by libpcnerf_b200.so since the last reset (counted inside the library at every launch site)."""
    return int(lib().pcnerf_launch_count(1 if reset else 0))


def profile(enable):
    """Switch the library's per-kernel-class CUDA-event timers on (resetting them) or off."""
    lib().pcnerf_prof_enable(1 if enable else 0)


def profile_read():
    """{class name: (device ms, launches, algorithmic work)} for every kernel class (synchronises the events)."""
    out = {}
    for i in range(lib().pcnerf_prof_classes()):
        ms, n, work = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        check(lib().pcnerf_prof_read(i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(work)))
        out[lib().pcnerf_prof_name(i).decode()] = (ms.value, n.value, work.value)
    return out


# Writes the LIBRARY makes to parameters / BN running statistics go through raw pointers and never bump torch's tensor
# version counters (pcnerf_adam_step on FlatAdam's flat buffer, k_bn_fold's running-statistics update in every
# training-mode forward, and any replay of a captured step).  Everything that caches values derived from those tensors
# (the folded eval-mode weights of MLPFunction) keys on this generation as well.
_PARAM_GEN = [0]


def note_param_write():
    """Tell the caches that parameters / running statistics were (or may have been) written by the library."""
    _PARAM_GEN[0] += 1


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("pcnerf_b200: %s must be a CUDA tensor (the B200 path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("pcnerf_b200: %s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _h3(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    return a, a.ctypes.data_as(_f64p)


_LINSPACE = {}


def linspace01(n, device):
    """torch.linspace(0, 1, n) evaluated on the CPU (the values the reference's CPU path and the golden fixtures see),
    cached on the device."""
    key = (int(n), str(device))
    t = _LINSPACE.get(key)
    if t is None:
        t = torch.linspace(0, 1, int(n)).to(device)
        _LINSPACE[key] = t
    return t


# ----------------------------------------------------------------------------------------------------------- K1 AABB


def _f64(t, device="cuda"):
    if isinstance(t, torch.Tensor):
        return t.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.asarray(t, dtype=np.float64), device=device).contiguous()


def _rays_od(ray_o, ray_d):
    d = _f64(ray_d).reshape(-1, 3)
    o = _f64(ray_o).reshape(-1, 3)
    if o.shape[0] == 1 and d.shape[0] != 1:
        o = o.expand(d.shape[0], 3).contiguous()
    return o, d


def aabb_far_bound(ray_o, ray_d, x_max, x_min, y_max, y_min, z_max, z_min):
    o, d = _rays_od(ray_o, ray_d)
    out = torch.empty(d.shape[0], dtype=torch.float64, device=d.device)
    keep, hp = _h3([x_max, x_min, y_max, y_min, z_max, z_min])
    check(lib().pcnerf_aabb_far_bound(_p(o), _p(d), d.shape[0], hp, _p(out), _stream()))
    return out


def aabb_slab(ray_o, ray_d, aabb_min, aabb_max):
    o, d = _rays_od(ray_o, ray_d)
    out = torch.empty(d.shape[0], dtype=torch.float64, device=d.device)
    k1, pmin = _h3(aabb_min)
    k2, pmax = _h3(aabb_max)
    check(lib().pcnerf_aabb_slab(_p(o), _p(d), d.shape[0], pmin, pmax, _p(out), _stream()))
    return out


def aabb_child_pairs(variant, ray_o, ray_d, boxes):
    o, d = _rays_od(ray_o, ray_d)
    b = _f64(boxes).reshape(-1, 6)
    n, K = d.shape[0], b.shape[0]
    flag = torch.empty((n, K), dtype=torch.uint8, device=d.device)
    near = torch.empty((n, K), dtype=torch.float64, device=d.device)
    far = torch.empty((n, K), dtype=torch.float64, device=d.device)
    check(lib().pcnerf_aabb_child_pairs(int(variant), _p(o), _p(d), n, _p(b), K, _p(flag), _p(near), _p(far), _stream()))
    return flag.bool(), near, far


def aabb_dist_to_ray(ray_o, ray_d, centres):
    o, d = _rays_od(ray_o, ray_d)
    c = _f64(centres).reshape(-1, 3)
    out = torch.empty((d.shape[0], c.shape[0]), dtype=torch.float64, device=d.device)
    check(lib().pcnerf_aabb_dist_to_ray(_p(o), _p(d), d.shape[0], _p(c), c.shape[0], _p(out), _stream()))
    return out


def aabb_find_box(centres, boxes, points, knn=10):
    c = _f64(centres).reshape(-1, 3)
    b = _f64(boxes).reshape(-1, 6)
    q = _f64(points).reshape(-1, 3)
    out = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    check(lib().pcnerf_aabb_find_box(_p(c), _p(b), c.shape[0], _p(q), q.shape[0], int(knn), _p(out), _stream()))
    return out


def aabb_pack_train(variant, ray_o, ray_d, dist, points, centres, boxes, boxes_bigger, parent, surface_expand, knn=10,
                    compact=True):
    """parent = (x_min, x_max, y_min, y_max, z_min, z_max).  Returns (rays (N,15) f32, keep mask) compacted in order.
    compact=False returns all N rows (rows with keep == 0 are undefined): no data-dependent shape, so the call can be
    captured in a CUDA graph when the caller knows every ray is kept."""
    o, d = _rays_od(ray_o, ray_d)
    dist = _f64(dist).reshape(-1)
    pts = _f64(points).reshape(-1, 3)
    c, b, bb = _f64(centres).reshape(-1, 3), _f64(boxes).reshape(-1, 6), _f64(boxes_bigger).reshape(-1, 6)
    n = d.shape[0]
    rays = torch.empty((n, 15), dtype=torch.float32, device=d.device)
    keep = torch.empty(n, dtype=torch.uint8, device=d.device)
    x_min, x_max, y_min, y_max, z_min, z_max = parent
    k, hp = _h3([x_max, x_min, y_max, y_min, z_max, z_min])
    check(lib().pcnerf_aabb_pack_train(int(variant), _p(o), _p(d), _p(dist), _p(pts), n, _p(c), _p(b), _p(bb), c.shape[0],
                                       hp, float(surface_expand), int(knn), _p(rays), _p(keep), _stream()))
    keep = keep.bool()
    return (rays[keep] if compact else rays), keep


def frame_returns(points_f32, pose, pose_xy, range_delete, max_range, over_height, over_low, interest_x, interest_y,
                  parent_box=None, position=None):
    """K0 (ipb2dmapping.py:662-711): raw sensor-frame points (n,3) float32 -> (world points (M,3), dir (M,3), dist (M,))
    float64 CUDA tensors of the points the block keeps, in their original order.  pose: (4,4) host array (float32 values),
    pose_xy: (P,2) float32 x / y translations of the run's poses, range_delete = (x, y, z)."""
    pts = points_f32 if isinstance(points_f32, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points_f32, dtype=np.float32))
    pts = _cuda_f32(pts.to("cuda") if not pts.is_cuda else pts, "points").reshape(-1, 3)
    n = pts.shape[0]
    dev = pts.device
    if pose_xy is None:
        pose_xy = np.zeros((0, 2), dtype=np.float32)
    pxy = pose_xy if isinstance(pose_xy, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pose_xy, dtype=np.float32))
    pxy = pxy.to(device=dev, dtype=torch.float32).contiguous().reshape(-1, 2)
    keep_h, hp = _h3(np.asarray(pose, dtype=np.float64).reshape(16))
    keep_b, hb = _h3(parent_box) if parent_box is not None else (None, None)
    keep_s, hs = _h3(position) if position is not None else (None, None)      # sensor position (default: pose[:3, 3])
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    world = torch.empty((n, 3), dtype=torch.float64, device=dev)
    dirs = torch.empty((n, 3), dtype=torch.float64, device=dev)
    dist = torch.empty(n, dtype=torch.float64, device=dev)
    check(lib().pcnerf_frame_returns(_p(pts), n, hp, _p(pxy), pxy.shape[0], float(range_delete[0]), float(range_delete[1]),
                                     float(range_delete[2]), float(max_range), float(over_height), float(over_low),
                                     float(interest_x), float(interest_y), hb, hs, _p(keep), _p(world), _p(dirs), _p(dist),
                                     _stream()))
    sel = keep.bool()
    return world[sel], dirs[sel], dist[sel]


def route_points(points, parent_boxes, origins=None, frame_id=None):
    """Multi-parent scenes: (which (n,) int32 = first parent box containing each return or -1, origin (n,3), dir (n,3),
    dist (n,) float64 rays from the sensor position of each return's frame).  parent_boxes (P,6) min xyz, max xyz."""
    pts = _f64(points).reshape(-1, 3)
    b = _f64(parent_boxes).reshape(-1, 6)
    n = pts.shape[0]
    dev = pts.device
    which = torch.empty(n, dtype=torch.int32, device=dev)
    o = d = r = None
    if origins is not None:
        origins = _f64(origins).reshape(-1, 3)
        if frame_id is not None:
            frame_id = frame_id.to(device=dev, dtype=torch.int32).contiguous()
        o = torch.empty((n, 3), dtype=torch.float64, device=dev)
        d = torch.empty((n, 3), dtype=torch.float64, device=dev)
        r = torch.empty(n, dtype=torch.float64, device=dev)
    check(lib().pcnerf_route_points(_p(pts), n, _p(b), b.shape[0], _p(origins), _p(frame_id), _p(which), _p(o), _p(d),
                                    _p(r), _stream()))
    return which, o, d, r


GROUPS_GRID_MIN_K = 1024       # child boxes from which aabb_build_groups bins the box centres on a uniform grid
_GRID_CACHE = {}


class BoxGrid:
    """Uniform x/y grid over the centres of the child boxes (cell size `h`), in CSR form for K1's group builder: the
    prefilter of eval_kitti_render.py:367-369 keeps boxes whose centre lies within 0.65 m of the ray's line, so a ray only
    has to look at the cells along its projection.  Built once per set of boxes (device-side sort + histogram)."""

    def __init__(self, boxes, h=1.0):
        c = (boxes[:, :3] + boxes[:, 3:]) / 2
        lo = c[:, :2].min(0).values - h                      # one cell of margin: no centre is ever clamped
        hi = c[:, :2].max(0).values + h
        x0, y0 = float(lo[0]), float(lo[1])
        self.nx = int(np.floor((float(hi[0]) - x0) / h)) + 1
        self.ny = int(np.floor((float(hi[1]) - y0) / h)) + 1
        ix = torch.floor((c[:, 0] - x0) / h).to(torch.int64).clamp_(0, self.nx - 1)
        iy = torch.floor((c[:, 1] - y0) / h).to(torch.int64).clamp_(0, self.ny - 1)
        cell = iy * self.nx + ix
        order = torch.argsort(cell, stable=True)              # ascending box index within a cell
        counts = torch.bincount(cell, minlength=self.nx * self.ny)
        start = torch.zeros(self.nx * self.ny + 1, dtype=torch.int64, device=boxes.device)
        start[1:] = torch.cumsum(counts, 0)
        self.cell_start = start.to(torch.int32).contiguous()
        self.cell_boxes = order.to(torch.int32).contiguous()
        self.h5 = np.array([x0, y0, float(h), float(self.nx), float(self.ny)], dtype=np.float64)


def _box_grid(b):
    key = (b.data_ptr(), b._version, b.shape[0])
    g = _GRID_CACHE.get(key)
    if g is None:
        if len(_GRID_CACHE) > 64:
            _GRID_CACHE.clear()
        g = _GRID_CACHE[key] = (BoxGrid(b), b)               # (keep the tensor alive: the key is its address)
    return g[0]


def aabb_build_groups(ray_o, ray_d, dist, boxes, boxes_larger, parent_min, parent_max, depth_inference_method=2,
                      grow_step=0.005, prefilter=0.65, grid=None):
    """Returns (rays (N',13) f32, ranges (N',1) f32, other (N',1) i64, kept ray mask (N,)).
    grid: None = a BoxGrid over the box centres when there are at least GROUPS_GRID_MIN_K boxes (identical results, an order
    of magnitude fewer boxes per ray at the shipped scenes' K), False = scan every box, or a prebuilt BoxGrid."""
    o, d = _rays_od(ray_o, ray_d)
    dist = _f64(dist).reshape(-1)
    b, bl = _f64(boxes).reshape(-1, 6), _f64(boxes_larger).reshape(-1, 6)
    n, K = d.shape[0], b.shape[0]
    if grid is None:
        grid = _box_grid(b) if (K >= GROUPS_GRID_MIN_K and isinstance(boxes, torch.Tensor)) else (BoxGrid(b) if K >= GROUPS_GRID_MIN_K else False)
    if grid is False:
        g5, gs, gb = None, None, None
    else:
        g5, gs, gb = grid.h5.ctypes.data_as(_f64p), _p(grid.cell_start), _p(grid.cell_boxes)
    count = torch.empty(n, dtype=torch.int32, device=d.device)
    pfar = torch.empty(n, dtype=torch.float64, device=d.device)
    k1, pmin = _h3(parent_min)
    k2, pmax = _h3(parent_max)
    check(lib().pcnerf_aabb_groups_count(_p(o), _p(d), n, _p(b), _p(bl), K, pmin, pmax, int(depth_inference_method),
                                         float(grow_step), float(prefilter), g5, gs, gb, _p(count), _p(pfar), _stream()))
    csum = torch.cumsum(count.to(torch.int64), 0)
    total = int(csum[-1].item()) if n > 0 else 0        # output size is data dependent: one host sync
    offset = (csum - count).contiguous()
    rays = torch.empty((total, 13), dtype=torch.float32, device=d.device)
    ranges = torch.empty((total, 1), dtype=torch.float32, device=d.device)
    other = torch.empty((total, 1), dtype=torch.int64, device=d.device)
    scratch = torch.empty((max(total, 1), 3), dtype=torch.float64, device=d.device)
    check(lib().pcnerf_aabb_groups_fill(_p(o), _p(d), _p(dist), n, _p(b), _p(bl), K, int(depth_inference_method),
                                        float(grow_step), float(prefilter), g5, gs, gb, _p(count), _p(offset), _p(pfar),
                                        _p(scratch), _p(rays), _p(ranges), _p(other), _stream()))
    return rays, ranges, other, count > 0


# ----------------------------------------------------------------------------------------------- K2 sample + encode


def sample_encode_coarse(rays, n_a, n_b=0, near_col=6, far_col=7, cnear_col=10, cfar_col=11, use_disp=False,
                         perturb=0.0, U=None, want_enc=True, f16=False):
    rays = _cuda_f32(rays, "rays")
    n, ld = rays.shape
    S = n_a + n_b
    sa = linspace01(n_a, rays.device)
    sb = linspace01(n_b, rays.device) if n_b > 0 else None
    if perturb > 0:
        if U is None:
            U = torch.rand((n, S), device=rays.device)
        U = _cuda_f32(U, "U")
    z = torch.empty((n, S), dtype=torch.float32, device=rays.device)
    enc = enc_bf = None
    if want_enc:
        if f16:
            enc_bf = torch.empty((n * S, 64), dtype=torch.float16, device=rays.device)
        else:
            enc = torch.empty((n * S, 64), dtype=torch.float32, device=rays.device)
    check(lib().pcnerf_sample_encode_coarse(_p(rays), ld, n, near_col, far_col, cnear_col, cfar_col, _p(sa), n_a, _p(sb),
                                            n_b, int(bool(use_disp)), float(perturb), _p(U) if perturb > 0 else None,
                                            _p(z), _p(enc), _p(enc_bf), _stream()))
    return z, (enc_bf if f16 else enc)


def sample_encode_fine(rays, z, w, Ni, u=None, det=True, want_enc=True, f16=False):
    rays = _cuda_f32(rays, "rays")
    z = _cuda_f32(z, "z")
    w = _cuda_f32(w.detach(), "w")
    n, S = z.shape
    if det:
        u = linspace01(Ni, rays.device)
        u_ld = 0
    else:
        u = _cuda_f32(u, "u")
        u_ld = Ni
    zf = torch.empty((n, S + Ni), dtype=torch.float32, device=rays.device)
    enc = enc_bf = None
    if want_enc:
        if f16:
            enc_bf = torch.empty((n * (S + Ni), 64), dtype=torch.float16, device=rays.device)
        else:
            enc = torch.empty((n * (S + Ni), 64), dtype=torch.float32, device=rays.device)
    check(lib().pcnerf_sample_encode_fine(_p(rays), rays.shape[1], n, _p(z), _p(w), S, _p(u), u_ld, Ni, _p(zf), _p(enc),
                                          _p(enc_bf), _stream()))
    return zf, (enc_bf if f16 else enc)


def sample_pdf(bins, weights, Ni, u=None, det=False):
    bins = _cuda_f32(bins, "bins")
    weights = _cuda_f32(weights, "weights")
    n, nb = bins.shape
    if weights.shape != (n, nb - 1):
        raise ValueError("sample_pdf: weights must be (N, len(bins)-1)")
    if det:
        u, u_ld = linspace01(Ni, bins.device), 0
    else:
        u, u_ld = _cuda_f32(u, "u"), Ni
    out = torch.empty((n, Ni), dtype=torch.float32, device=bins.device)
    check(lib().pcnerf_sample_pdf(_p(bins), _p(weights), n, nb, _p(u), u_ld, Ni, _p(out), _stream()))
    return out


def embed(x, out_ld=63):
    x = _cuda_f32(x, "x")
    if x.dim() != 2 or x.shape[1] != 3:
        raise ValueError("embed: x must be (B,3)")
    out = torch.empty((x.shape[0], out_ld), dtype=torch.float32, device=x.device)
    check(lib().pcnerf_embed(_p(x), x.shape[0], _p(out), out_ld, _stream()))
    return out


# ------------------------------------------------------------------------------------------------------- K3 MLP

_SCRATCH = {}
_SCRATCH_RETIRED = []             # outgrown scratch buffers, kept alive for the graphs that may still reference them
EVAL_CHUNK_FLOOR = 1 << 20        # rows per launch of the eval-mode tensor-core MLP (tests lower it to cover chunking)


TC_LANES = int(__import__('os').environ.get('PCNERF_TC_LANES', '4'))   # BN chunks in flight in the training-mode tensor-core MLP (1..4; measured 61.4 / 60.1 / 59.8 ms per C2 step with 2 / 3 / 4)


def _scratch(rows, precision, device, lane=0):
    need = lib().pcnerf_mlp_scratch_bytes(rows, precision)
    key = (str(device), precision, lane)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        if buf is not None:
            # a CUDA graph captured earlier has the old buffer's address baked into its kernel nodes: never free it
            _SCRATCH_RETIRED.append(buf)
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


def _mlp_params(tensors, buffers, training, precision, momentum=0.1, eps=1e-5):
    """tensors: 34 parameter tensors in mlp_param_order; buffers: (running_mean[8], running_var[8], nbt[8])."""
    P = MlpParams()
    for l in range(9):
        P.W[l] = tensors[2 * l].data_ptr()
        P.b[l] = tensors[2 * l + 1].data_ptr()
    for l in range(8):
        P.gamma[l] = tensors[18 + 2 * l].data_ptr()
        P.beta[l] = tensors[19 + 2 * l].data_ptr()
        P.running_mean[l] = buffers[0][l].data_ptr()
        P.running_var[l] = buffers[1][l].data_ptr()
        P.num_batches_tracked[l] = buffers[2][l].data_ptr()
    P.momentum, P.eps, P.training, P.precision = momentum, eps, int(training), int(precision)
    P.prepared = 0
    return P


DIRECT_PARAM_GRADS = True     # MLP backward accumulates into existing .grad tensors instead of returning fresh gradients

GRAD_SIZES = [256 * 63, 256] + [256 * 256, 256] * 3 + [256 * 319, 256] + [256 * 256, 256] * 3 + [256, 1] + [256, 256] * 8


def tc_fused_eval(on=None):
    """Get / set the eval-mode engine of the precision-1 MLP: fused single kernel (default) or layered row GEMMs."""
    if on is not None:
        lib().pcnerf_tc_set_fused_eval(int(on))
    return int(lib().pcnerf_tc_get_fused_eval())


def tc_row_pairs(on=None):
    """Get / set the training-mode GEMM forms of the precision-1 MLP (bit mask): bit 0 = forward, bit 1 = data gradient
    (k_tc_rowgemm2: cta_group::2), bit 2 = weight gradient (k_tc_wgrad2) on CTA pairs; 0 = the single-CTA kernels."""
    if on is not None:
        lib().pcnerf_tc_set_row_pairs(int(on))
    return int(lib().pcnerf_tc_get_row_pairs())


def tc_weight_correction(on=None):
    """Get / set the linear correction of the fp16 weight rounding in the layered precision-1 forward (k_tc_fold)."""
    if on is not None:
        lib().pcnerf_tc_set_weight_correction(int(on))
    return int(lib().pcnerf_tc_get_weight_correction())


class MLPFunction(torch.autograd.Function):
    """p = NOF(enc) evaluated chunk by chunk (one BN batch per chunk, nof/render.py:47-49)."""

    @staticmethod
    def forward(ctx, enc, chunk, training, precision, buffers, cache, *params):
        rows = enc.shape[0]
        dev = enc.device
        P = _mlp_params(params, buffers, training, precision)
        out = torch.empty(rows, dtype=torch.float32, device=dev)
        need_grad = training and any(ctx.needs_input_grad[6:])
        if training:
            note_param_write()          # k_bn_fold updates the BN running statistics through raw pointers
        if not training and precision == 1:
            # eval-mode BN is row-wise (running statistics): `chunk` (an OOM guard in the reference, nof/render.py:21-24)
            # does not change any value, and the row GEMMs run closer to their steady-state rate on >= 1 M-row launches
            chunk = max(int(chunk), EVAL_CHUNK_FLOOR)
        saved = []
        if need_grad and precision == 1:
            # training pass of the tensor-core engine: the library walks the chunks itself, TC_LANES of them in flight
            nch = -(-rows // chunk)
            saved = [torch.empty(lib().pcnerf_mlp_saved_bytes(min(chunk, rows - c * chunk), 1), dtype=torch.uint8, device=dev)
                     for c in range(nch)]
            lanes = TC_LANES if nch > 1 else 1
            scr = [_scratch(min(chunk, rows), 1, dev, k) for k in range(lanes)]
            sv_arr = (ctypes.c_void_p * nch)(*[t.data_ptr() for t in saved])
            sb_arr = (ctypes.c_size_t * nch)(*[t.numel() for t in saved])
            sc_arr = (ctypes.c_void_p * lanes)(*[t.data_ptr() for t in scr])
            check(lib().pcnerf_mlp_tc_forward_chunks(ctypes.byref(P), _p(enc), rows, chunk, _p(out), sv_arr, sb_arr, sc_arr,
                                                     min(t.numel() for t in scr), lanes, _stream()))
            ctx.saved_chunks = saved
            ctx.meta = (chunk, precision, buffers, rows)
            ctx.save_for_backward(enc, out, *params)
            return out
        shared = None
        esz = 2 if precision == 1 else 4
        fused_eval = (not training) and precision == 1 and bool(lib().pcnerf_tc_get_fused_eval())
        if fused_eval:
            # The folded fp16 weights of an eval-mode model are a function of its parameters and running statistics only:
            # they live in a small per-model scratch and are re-derived only when one of those tensors was written to
            # (tensor version counters), not on every call -- 18 small launches per call otherwise.
            # `cache` is a dict owned by the model (its lifetime bounds the cached copies)
            cache = cache if cache is not None else {}
            ver = (_PARAM_GEN[0],) + tuple((t.data_ptr(), t._version) for t in params) + \
                tuple((b.data_ptr(), b._version) for grp in buffers[:2] for b in grp)
            scratch = cache.get("scratch")
            if scratch is None or scratch.device != dev:
                if scratch is not None:
                    _SCRATCH_RETIRED.append(scratch)
                scratch = torch.empty(lib().pcnerf_mlp_scratch_bytes(1, 1), dtype=torch.uint8, device=dev)
                cache["scratch"], cache["ver"] = scratch, None
            P.prepared = 2 if cache.get("ver") == ver else 0       # 2: weight copies valid, per-call constants to be loaded
            cache["ver"] = ver
        else:
            scratch = _scratch(min(chunk, rows), precision, dev)
        for i in range(0, rows, chunk):
            r = min(chunk, rows - i)
            # (the fused eval kernel keeps every activation on chip: nothing is saved)
            nbytes = 256 if fused_eval else lib().pcnerf_mlp_saved_bytes(r, precision)
            if need_grad:
                sv = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                saved.append(sv)
            else:
                if shared is None or shared.numel() < nbytes:
                    shared = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                sv = shared
            check(lib().pcnerf_mlp_forward(ctypes.byref(P), ctypes.c_void_p(enc.data_ptr() + i * 64 * esz), r,
                                           ctypes.c_void_p(out.data_ptr() + i * 4), _p(sv), sv.numel(), _p(scratch),
                                           scratch.numel(), _stream()))
            P.prepared = 1              # later chunks of this pass reuse the weight copies in `scratch`
        ctx.saved_chunks = saved
        ctx.meta = (chunk, precision, buffers, rows)
        ctx.save_for_backward(enc, out, *params)
        return out

    @staticmethod
    def backward(ctx, gp):
        chunk, precision, buffers, rows = ctx.meta
        enc, out, *params = ctx.saved_tensors
        dev = enc.device
        gp = gp.contiguous()
        P = _mlp_params(params, buffers, True, precision)
        # The library ACCUMULATES (+=) parameter gradients.  When every parameter already owns a dense fp32 .grad (the flat
        # GradBucket of FlatAdam / parallel, or any earlier backward), it accumulates straight into those tensors and autograd
        # is handed None -- otherwise 68 tiny `grad += new` kernels per step (one per parameter) follow every backward.
        # Only for parameters whose .grad aliases a parallel.GradBucket (FlatAdam): that is an explicit opt-in to "gradients
        # live in the bucket"; torch.autograd.grad() on such parameters is not supported.
        direct = DIRECT_PARAM_GRADS and all(
            getattr(p, "_pcnerf_bucketed", False) and isinstance(p.grad, torch.Tensor) and p.grad.dtype == torch.float32 and
            p.grad.is_contiguous() and p.grad.device == dev and p.grad.shape == p.shape for p in params)
        if direct:
            views = [p.grad for p in params]
        else:
            flat = torch.zeros(sum(GRAD_SIZES), dtype=torch.float32, device=dev)
            views, o = [], 0
            for sz_, p in zip(GRAD_SIZES, params):
                views.append(flat[o:o + sz_].view_as(p))
                o += sz_
        ret = (None,) * len(views) if direct else tuple(views)
        G = MlpGrads()
        for l in range(9):
            G.dW[l] = views[2 * l].data_ptr()
            G.db[l] = views[2 * l + 1].data_ptr()
        for l in range(8):
            G.dgamma[l] = views[18 + 2 * l].data_ptr()
            G.dbeta[l] = views[19 + 2 * l].data_ptr()
        if precision == 1:
            nch = len(ctx.saved_chunks)
            lanes = TC_LANES if nch > 1 else 1
            scr = [_scratch(min(chunk, rows), 1, dev, k) for k in range(lanes)]
            sv_arr = (ctypes.c_void_p * nch)(*[t.data_ptr() for t in ctx.saved_chunks])
            sb_arr = (ctypes.c_size_t * nch)(*[t.numel() for t in ctx.saved_chunks])
            sc_arr = (ctypes.c_void_p * lanes)(*[t.data_ptr() for t in scr])
            check(lib().pcnerf_mlp_tc_backward_chunks(ctypes.byref(P), ctypes.byref(G), _p(enc), rows, chunk, _p(out), _p(gp),
                                                      sv_arr, sb_arr, sc_arr, min(t.numel() for t in scr), lanes, _stream()))
            ctx.saved_chunks = None
            return (None, None, None, None, None, None) + ret
        scratch = _scratch(min(chunk, rows), precision, dev)
        esz = 2 if precision == 1 else 4
        for ci_, i in enumerate(range(0, rows, chunk)):
            r = min(chunk, rows - i)
            sv = ctx.saved_chunks[ci_]
            check(lib().pcnerf_mlp_backward(ctypes.byref(P), ctypes.byref(G), ctypes.c_void_p(enc.data_ptr() + i * 64 * esz),
                                            r, ctypes.c_void_p(out.data_ptr() + i * 4),
                                            ctypes.c_void_p(gp.data_ptr() + i * 4), _p(sv), sv.numel(), _p(scratch),
                                            scratch.numel(), _stream()))
            P.prepared = 1
        ctx.saved_chunks = None
        return (None, None, None, None, None, None) + ret


# ------------------------------------------------------------------------------------ K3' closed-form ("affine") MLP


def affine_moments(enc, chunk):
    """Per-chunk first and second moments of the (rows,64) fp32 encodings.
    Returns (m (nc,64) f64, C (nc,64,64) f64 covariance, counts (nc,) f64)."""
    enc = _cuda_f32(enc, "enc")
    rows = enc.shape[0]
    nc = -(-rows // chunk)
    P = lib().pcnerf_affine_parts()
    part = torch.empty((nc, P, 65, 64), dtype=torch.float64, device=enc.device)
    check(lib().pcnerf_affine_moments(_p(enc), rows, int(chunk), _p(part), _stream()))
    tot = part.sum(1)
    cnt = torch.full((nc,), float(chunk), dtype=torch.float64, device=enc.device)
    cnt[-1:].fill_(float(rows - (nc - 1) * chunk))           # (fill_, not setitem: no host scalar tensor -> graph-safe)
    shift = enc[::chunk].to(torch.float64)
    m1 = tot[:, 64, :] / cnt[:, None]
    C = tot[:, :64, :] / cnt[:, None, None] - m1[:, :, None] * m1[:, None, :]
    return shift + m1, C, cnt


class AffineApplyFunction(torch.autograd.Function):
    """p_r = sigmoid(alpha[chunk(r)] . x_r + c[chunk(r)]); backward returns sum_r g_r x_r and sum_r g_r per chunk."""

    @staticmethod
    def forward(ctx, enc, alpha, c, chunk):
        enc = _cuda_f32(enc, "enc")
        alpha = _cuda_f32(alpha, "alpha")
        c = _cuda_f32(c, "c")
        rows = enc.shape[0]
        p = torch.empty(rows, dtype=torch.float32, device=enc.device)
        check(lib().pcnerf_affine_apply(_p(enc), rows, int(chunk), _p(alpha), _p(c), _p(p), _stream()))
        ctx.save_for_backward(enc, p)
        ctx.chunk = int(chunk)
        ctx.nc = alpha.shape[0]
        return p

    @staticmethod
    def backward(ctx, gp):
        enc, p = ctx.saved_tensors
        gp = gp.contiguous()
        P = lib().pcnerf_affine_parts()
        part = torch.empty((ctx.nc, P, 65), dtype=torch.float64, device=enc.device)
        check(lib().pcnerf_affine_grad(_p(enc), _p(p), _p(gp), enc.shape[0], ctx.chunk, _p(part), _stream()))
        g = part.sum(1)
        return None, g[:, :64].to(torch.float32), g[:, 64].to(torch.float32), None


class LazyEnc:
    """The (rows,64) encoding tensor of a sampling pass, NOT materialised: rows are (ray, depth) pairs and every kernel of the
    closed-form engine that needs a row's encoding re-derives it from rays[:, 0:6] and z (csrc/affine_rays.cu).  Produced by
    nof/render.py for precision-2 models (training passes, and eval-mode passes without autograd); `materialise()` gives the
    tensor K2 would have written."""

    def __init__(self, rays, z):
        self.rays = _cuda_f32(rays, "rays")
        self.z = _cuda_f32(z, "z")
        if self.rays.dim() != 2 or self.rays.shape[1] < 6 or self.z.dim() != 2 or self.z.shape[0] != self.rays.shape[0]:
            raise ValueError("LazyEnc: rays (N, >=6) and z (N, S) expected")
        self.shape = (self.z.shape[0] * self.z.shape[1], 64)
        self.dtype = torch.float32
        self.device = self.z.device

    def materialise(self):
        pts = self.rays[:, None, 0:3] + self.rays[:, None, 3:6] * self.z[..., None]        # nof/render.py:458
        return embed(pts.reshape(-1, 3).contiguous(), 64)


def _grad_views(params, dev):
    """(views, returned): where a backward accumulates the 34 parameter gradients -- straight into the existing .grad tensors
    when every parameter lives in a parallel.GradBucket (autograd is handed None), else into a fresh flat buffer."""
    direct = DIRECT_PARAM_GRADS and all(
        getattr(p, "_pcnerf_bucketed", False) and isinstance(p.grad, torch.Tensor) and p.grad.dtype == torch.float32 and
        p.grad.is_contiguous() and p.grad.device == dev and p.grad.shape == p.shape for p in params)
    if direct:
        views = [p.grad for p in params]
    else:
        flat = torch.zeros(sum(GRAD_SIZES), dtype=torch.float32, device=dev)
        views, o = [], 0
        for sz_, p in zip(GRAD_SIZES, params):
            views.append(flat[o:o + sz_].view_as(p))
            o += sz_
    G = MlpGrads()
    for l in range(9):
        G.dW[l] = views[2 * l].data_ptr()
        G.db[l] = views[2 * l + 1].data_ptr()
    for l in range(8):
        G.dgamma[l] = views[18 + 2 * l].data_ptr()
        G.dbeta[l] = views[19 + 2 * l].data_ptr()
    return G, views, ((None,) * len(views) if direct else tuple(views))


class AffineRaysFunction(torch.autograd.Function):
    """p = NOF(embed(o + d z)) of a TRAINING pass of the closed-form engine, from (rays, z): batch moments, float64 algebra
    and per-row dot + sigmoid in repo kernels, hand-derived backward (pcnerf_affine_forward_rays / _backward_rays)."""

    @staticmethod
    def forward(ctx, rays, z, chunk, buffers, *params):
        dev = z.device
        n, S = z.shape
        rows = n * S
        nc = -(-rows // chunk)
        P = _mlp_params(params, buffers, True, 2)
        out = torch.empty(rows, dtype=torch.float32, device=dev)
        work = torch.empty(lib().pcnerf_affine_work_bytes(nc), dtype=torch.uint8, device=dev)
        note_param_write()              # the BN running statistics are updated through raw pointers
        check(lib().pcnerf_affine_forward_rays(ctypes.byref(P), _p(rays), rays.shape[1], n, _p(z), S, int(chunk), _p(out),
                                               _p(work), work.numel(), _stream()))
        ctx.meta = (int(chunk), buffers)
        ctx.work = work
        ctx.save_for_backward(rays, z, out, *params)
        return out

    @staticmethod
    def backward(ctx, gp):
        chunk, buffers = ctx.meta
        rays, z, out, *params = ctx.saved_tensors
        n, S = z.shape
        P = _mlp_params(params, buffers, True, 2)
        G, views, ret = _grad_views(params, z.device)
        gp = gp.contiguous()
        work = ctx.work
        check(lib().pcnerf_affine_backward_rays(ctypes.byref(P), ctypes.byref(G), _p(rays), rays.shape[1], n, _p(z), S, chunk,
                                                _p(out), _p(gp), _p(work), work.numel(), _stream()))
        ctx.work = None
        return (None, None, None, None) + ret


def affine_eval_alpha(params, buffers, cache):
    """alpha (64,) f32 on the device: the eval-mode logit of the closed-form engine is alpha . (x, 1) (running statistics:
    a function of the parameters alone).  Re-derived (17 small float64 kernels, pcnerf_affine_eval_alpha) only when a parameter
    or a running statistic was written to since the last call -- same versioning as the folded weights of the tensor-core
    engine (tensor version counters + note_param_write for library-side writes).  `cache`: a dict owned by the model."""
    dev = params[0].device
    ver = (_PARAM_GEN[0],) + tuple((t.data_ptr(), t._version) for t in params) + \
        tuple((b.data_ptr(), b._version) for grp in buffers[:2] for b in grp)
    alpha = cache.get("affine_alpha")
    if alpha is None or alpha.device != dev:
        if alpha is not None:
            _SCRATCH_RETIRED.append((alpha, cache.get("affine_work")))     # a captured graph may still reference them
        alpha = torch.empty(64, dtype=torch.float32, device=dev)
        cache["affine_alpha"] = alpha
        cache["affine_work"] = torch.empty(lib().pcnerf_affine_work_bytes(1), dtype=torch.uint8, device=dev)
        cache["affine_ver"] = None
    if cache.get("affine_ver") != ver:
        P = _mlp_params(params, buffers, False, 2)
        work = cache["affine_work"]
        check(lib().pcnerf_affine_eval_alpha(ctypes.byref(P), _p(alpha), _p(work), work.numel(), _stream()))
        cache["affine_ver"] = ver
    return alpha


def affine_apply_rays(rays, z, alpha):
    """p (n S,) = sigmoid(alpha . (embed(o + d z), 1)) per (ray, depth) row, no encoding tensor (pcnerf_affine_apply_rays)."""
    rays = _cuda_f32(rays, "rays")
    z = _cuda_f32(z, "z")
    n, S = z.shape
    out = torch.empty(n * S, dtype=torch.float32, device=z.device)
    check(lib().pcnerf_affine_apply_rays(_p(rays), rays.shape[1], n, _p(z), S, _p(alpha), _p(out), _stream()))
    return out


# -------------------------------------------------------------------------------------------- K4 composite + losses


class CompositeFunction(torch.autograd.Function):
    """(w, depth, child_free_loss, child_depth_loss, free_r, sl1_r, opacity, range_sl1) = composite(p, z, rays)."""

    @staticmethod
    def forward(ctx, p, z, rays, cols, noise, noise_std, epsilon, flags, want_per_ray=False):
        p = _cuda_f32(p, "p")
        z = _cuda_f32(z, "z")
        n, P_ = z.shape
        dev = z.device
        child = bool(flags & COMP_CHILD_LOSS)
        rloss = bool(flags & COMP_RANGE_LOSS)
        if rays is not None:
            rays = _cuda_f32(rays, "rays")
        ld = rays.shape[1] if rays is not None else 0
        cn, cf, rc = cols
        w = torch.empty((n, P_), dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        per_ray = torch.empty((n, 8), dtype=torch.float32, device=dev) if child else None
        sums = torch.empty(5, dtype=torch.float64, device=dev)
        if noise is not None:
            noise = _cuda_f32(noise, "noise")
        # (the three loss scalars are written by the last block of the forward kernel: no finaliser launch)
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        check(lib().pcnerf_composite_fwd(_p(p), _p(z), _p(rays), ld, n, P_, cn, cf, rc, _p(noise), float(noise_std),
                                         float(epsilon), int(flags), _p(w), _p(depth), _p(per_ray), _p(sums), _p(losses),
                                         _stream()))
        if flags & COMP_OPACITY:
            opacity = (sums[2] / max(n * P_, 1)).to(torch.float32)
        else:
            opacity = torch.zeros((), dtype=torch.float32, device=dev)
        ctx.save_for_backward(p, z, w, rays, per_ray, depth if rloss else None)
        ctx.meta = (rc, float(noise_std), float(epsilon), int(flags), n, P_, ld)
        ctx.set_materialize_grads(False)
        free_r = per_ray[:, 0].contiguous() if (child and want_per_ray) else None
        sl1_r = per_ray[:, 2].contiguous() if (child and want_per_ray) else None
        ctx.mark_non_differentiable(w, opacity)
        # (views of one 3-element tensor: no clone kernels; autograd hands back one gradient per view)
        return w, depth, losses[0], losses[1], free_r, sl1_r, opacity, losses[2]

    @staticmethod
    def backward(ctx, g_w, g_depth, g_free, g_dl, g_free_r, g_sl1_r, g_op, g_range):
        p, z, w, rays, per_ray, depth = ctx.saved_tensors
        rc, noise_std, epsilon, flags, n, P_, ld = ctx.meta
        if g_w is not None or g_op is not None:
            raise NotImplementedError("pcnerf_b200: gradients through `weights` / `opacity` outputs are not supported "
                                      "(the reference never back-propagates through them)")
        gp = torch.empty((n, P_), dtype=torch.float32, device=p.device)

        def c(t):
            return None if t is None else t.contiguous().to(torch.float32)

        g_depth, g_free, g_dl, g_free_r, g_sl1_r, g_range = c(g_depth), c(g_free), c(g_dl), c(g_free_r), c(g_sl1_r), c(g_range)
        check(lib().pcnerf_composite_bwd(_p(p), _p(z), _p(w), _p(rays), ld, n, P_, rc, noise_std, epsilon, flags,
                                         _p(per_ray), _p(g_depth), _p(g_free), _p(g_dl), _p(g_free_r), _p(g_sl1_r), n,
                                         _p(depth), _p(g_range), _p(gp), _stream()))
        return gp, None, None, None, None, None, None, None, None


def composite(p, z, rays=None, cols=(10, 11, 14), noise=None, noise_std=0.0, epsilon=1e-10, flags=0, want_per_ray=False):
    """-> (w, depth, child_free_loss, child_depth_loss, free_r, sl1_r, opacity, range_sl1).  range_sl1 (flags &
    COMP_RANGE_LOSS) = SmoothL1Loss(mean)(10 depth, 10 rays[:, cols[2]]), the scene-level range term of
    train_kitti.py:145-146 before its 0.1 * lambda_loss factor, computed (and back-propagated) inside K4."""
    return CompositeFunction.apply(p, z, rays, cols, noise, noise_std, epsilon, flags, want_per_ray)


# ------------------------------------------------------------------------------------ masked-mean losses / metrics

LOSS_KINDS = {"smoothl1": 0, "mse": 1, "l1": 2, "abs_error": 3, "acc_thres": 4}


def _mask_u8(valid_mask, n, device):
    if valid_mask is None:
        return None
    m = valid_mask.reshape(-1)
    if m.shape[0] != n:
        raise ValueError("valid_mask must have one entry per element")
    return m.to(device=device, dtype=torch.uint8).contiguous()


class MaskedLossFunction(torch.autograd.Function):
    """mean over the selected elements of SmoothL1 / MSE / L1 (pred - target): one reduction kernel + a finaliser forward,
    one elementwise kernel backward (nof/criteria/loss.py:7-50)."""

    @staticmethod
    def forward(ctx, pred, target, mask, kind):
        pred = _cuda_f32(pred.reshape(-1), "pred")
        target = _cuda_f32(target.reshape(-1), "target")
        n = pred.shape[0]
        if target.shape[0] != n:
            raise ValueError("pred and target must have the same number of elements")
        acc = torch.empty(2, dtype=torch.float64, device=pred.device)
        out = torch.empty(1, dtype=torch.float32, device=pred.device)
        check(lib().pcnerf_masked_loss_fwd(int(kind), _p(pred), _p(target), _p(mask), n, _p(acc), _p(out), _stream()))
        ctx.save_for_backward(pred, target, mask, acc)
        ctx.kind = int(kind)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        pred, target, mask, acc = ctx.saved_tensors
        g = g.reshape(1).contiguous().to(torch.float32)
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gp = torch.empty_like(pred) if need_p else None
        gt = torch.empty_like(target) if need_t else None
        if need_p or need_t:
            check(lib().pcnerf_masked_loss_bwd(ctx.kind, _p(pred), _p(target), _p(mask), pred.shape[0], _p(acc), _p(g), _p(gp),
                                               _p(gt), _stream()))
        return gp, gt, None, None


def masked_loss(pred, target, valid_mask=None, kind="smoothl1"):
    """Masked mean of the elementwise loss / metric `kind` (LOSS_KINDS) of pred - target; differentiable for the losses."""
    k = LOSS_KINDS[kind]
    shape = pred.shape
    mask = _mask_u8(valid_mask, pred.numel(), pred.device)
    if k >= 3:
        with torch.no_grad():
            return MaskedLossFunction.apply(pred.detach(), target.detach().expand(shape), mask, k)
    return MaskedLossFunction.apply(pred, target.expand(shape), mask, k)


# -------------------------------------------------------------------------------------------------------- K5 search


def search_rows(p, z, rays, cnear_col=6, cfar_col=7, epsilon=1e-10, method=0, row_ray=None, want_w=True):
    """K5 per candidate row.  row_ray (N',) int32: p / z are (G,P) per PHYSICAL ray and row r of `rays` reads ray
    row_ray[r] (the weights come back per ray as well); None: one p / z row per candidate row (the reference's layout)."""
    p, z, rays = _cuda_f32(p, "p"), _cuda_f32(z, "z"), _cuda_f32(rays, "rays")
    n = rays.shape[0] if row_ray is not None else z.shape[0]
    P_ = z.shape[1]
    dev = z.device
    w = torch.empty(z.shape, dtype=torch.float32, device=dev) if (want_w or row_ray is None) else None
    depth = torch.empty(n, dtype=torch.float32, device=dev)
    peak = torch.empty(n, dtype=torch.uint8, device=dev)
    wsum = torch.empty(n, dtype=torch.float32, device=dev)
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    if row_ray is None:
        check(lib().pcnerf_search_rows(_p(p), _p(z), _p(rays), rays.shape[1], n, P_, cnear_col, cfar_col, float(epsilon),
                                       int(method), _p(w), _p(depth), _p(peak), _p(wsum), _p(sums), _stream()))
    else:
        if row_ray.dtype != torch.int32 or row_ray.shape[0] != n:
            raise TypeError("search_rows: row_ray must be an int32 tensor with one entry per candidate row")
        check(lib().pcnerf_search_rows_grouped(_p(p), _p(z), _p(rays), rays.shape[1], n, P_, cnear_col, cfar_col,
                                               float(epsilon), int(method), _p(row_ray.contiguous()), _p(w), _p(depth),
                                               _p(peak), _p(wsum), _p(sums), _stream()))
    opacity = (sums[0] / max(n * P_, 1)).to(torch.float32)
    return depth, w, opacity, peak, wsum


def search_select(other, peak, wsum, n_rendered=None):
    other = other.reshape(-1).to(torch.int64).contiguous()
    n = other.shape[0]
    flag = torch.empty(n, dtype=torch.uint8, device=other.device)
    check(lib().pcnerf_search_select(_p(other), _p(peak), _p(wsum), n, _p(n_rendered), _p(flag), _stream()))
    return flag.bool().reshape(-1, 1)


class GroupPlan:
    """Group structure of a set of candidate rows (nof/render.py:317-340 walks it row by row on the host): which rows start
    a group (= one physical LiDAR ray), the row -> ray map, and whether every row of a group really carries its head's ray
    (origin, direction, parent segment) -- the condition under which the samples are evaluated once per ray.
    Building it costs ONE host synchronisation (the number of physical rays sizes every later launch)."""

    def __init__(self, rays, other, pnear_col=9, pfar_col=10):
        rays = _cuda_f32(rays, "rays")
        other = other.reshape(-1).to(torch.int64).contiguous()
        n = rays.shape[0]
        if other.shape[0] != n:
            raise ValueError("GroupPlan: `other` must have one entry per candidate row")
        dev = rays.device
        head = torch.empty(n, dtype=torch.uint8, device=dev)
        check(lib().pcnerf_group_heads(_p(other), n, _p(head), _stream()))
        self.row_ray = (torch.cumsum(head, 0, dtype=torch.int32) - 1).contiguous()      # (N',) physical-ray index of a row
        self.head_rows = torch.nonzero(head).reshape(-1)                                # (G,) first row of every group (sync)
        self.G = int(self.head_rows.shape[0])
        mism = torch.zeros(1, dtype=torch.int32, device=dev)
        if n:
            check(lib().pcnerf_group_uniform(_p(rays), rays.shape[1], n, _p(self.row_ray), _p(self.head_rows), pnear_col,
                                             pfar_col, _p(mism), _stream()))
        self.uniform = int(mism.item()) == 0
        self.n = n
        self.other = other


def eval_rows_rendered(rays, batch_size_set, tag_col=-1):
    """Device scalar (int64): the number of leading rows the reference's batch loop hands to the renderer
    (eval_kitti_render.py:979-1005; n, or n - 1 when a batch ends one row before the end)."""
    rays = _cuda_f32(rays, "rays")
    ld = rays.shape[1]
    out = torch.empty(1, dtype=torch.int64, device=rays.device)
    check(lib().pcnerf_eval_rows_rendered(_p(rays), ld, tag_col % ld, rays.shape[0], int(batch_size_set), _p(out), _stream()))
    return out


def points(rays, depth):
    rays, depth = _cuda_f32(rays, "rays"), _cuda_f32(depth, "depth")
    out = torch.empty((rays.shape[0], 3), dtype=torch.float32, device=rays.device)
    check(lib().pcnerf_points(_p(rays), rays.shape[1], rays.shape[0], _p(depth), _p(out), _stream()))
    return out


# ------------------------------------------------------------------------------- tensor-core building blocks (tests)


def tc_rowgemm(mode, A0, B, A1=None, vec=None, E=None, want_bf16_copy=False):
    """C (rows,256) = [A0 | A1] @ B.T on tcgen05.
    mode 0: fp16 operands; returns (out = fp16(C + vec), None, stats (2,256) f64 = col sums of out and out^2).
    mode 1: bf16 operands, E fp16, vec = (4,256) [c0, c1, c2, mean]; returns (bf16(c0*C - c1 - (E-mean)*c2), None, stats)."""
    dt = torch.float16 if mode == 0 else torch.bfloat16
    for t in (A0, B) + ((A1,) if A1 is not None else ()):
        if not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
            raise TypeError("tc_rowgemm: operands must be contiguous CUDA %s tensors" % dt)
    vec = _cuda_f32(vec, "vec")
    rows, k0 = A0.shape
    k1 = 0 if A1 is None else A1.shape[1]
    out = torch.empty((rows, 256), dtype=dt, device=A0.device)
    out2 = torch.empty((rows, 256), dtype=torch.bfloat16, device=A0.device) if (mode == 0 and want_bf16_copy) else None
    stats = torch.empty((2, 256), dtype=torch.float64, device=A0.device)
    work = torch.empty(lib().pcnerf_tc_rowgemm_work_bytes(), dtype=torch.uint8, device=A0.device)
    check(lib().pcnerf_tc_rowgemm(int(mode), _p(A0), k0, _p(A1), k1, _p(B), _p(vec), _p(E), rows, _p(out), _p(out2),
                                  _p(stats), _p(work), _stream()))
    return out, out2, stats


def tc_wgrad(DH, X, ncols, out, col_off=0):
    """out[256, ldo] window [col_off, col_off+ncols) += DH.T @ X[:, :ncols] on tcgen05 (DH bf16, X fp16 or bf16)."""
    if not (DH.is_cuda and DH.dtype == torch.bfloat16 and DH.is_contiguous() and DH.shape[1] == 256):
        raise TypeError("tc_wgrad: DH must be a contiguous CUDA bf16 (rows,256) tensor")
    if X.dtype not in (torch.float16, torch.bfloat16) or not X.is_contiguous():
        raise TypeError("tc_wgrad: X must be contiguous fp16 / bf16")
    check(lib().pcnerf_tc_wgrad(_p(DH), _p(X), X.shape[1], int(ncols), int(X.dtype == torch.bfloat16), DH.shape[0], _p(out),
                                out.shape[1], int(col_off), _stream()))
    return out


# ------------------------------------------------------------------------------------------------- K6 point-cloud metrics


def nn_correspondance(verts1, verts2):
    """For each vertex of verts2 the exact nearest vertex of verts1 -> (indices (n2,) i32, distances (n2,) f64)."""
    v1 = _f64(verts1).reshape(-1, 3)
    v2 = _f64(verts2).reshape(-1, 3)
    idx = torch.empty(v2.shape[0], dtype=torch.int32, device=v2.device)
    dist = torch.empty(v2.shape[0], dtype=torch.float64, device=v2.device)
    check(lib().pcnerf_nn_correspondance(_p(v1), v1.shape[0], _p(v2), v2.shape[0], _p(idx), _p(dist), _stream()))
    return idx, dist


def dist_stats(dist, threshold):
    """(sum of distances, number below threshold) as a (2,) f64 device tensor."""
    dist = dist.contiguous()
    out = torch.empty(2, dtype=torch.float64, device=dist.device)
    check(lib().pcnerf_dist_stats(_p(dist), dist.shape[0], float(threshold), _p(out), _stream()))
    return out
