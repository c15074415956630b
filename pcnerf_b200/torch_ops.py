"""`torch.ops.pcnerf.*`: the stages of the path registered with the PyTorch dispatcher for the CUDA key ONLY (SURVEY.md
section 8b) -- a CPU tensor finds no kernel and the dispatcher raises; there is no fallback to register.  Every op is a thin
wrapper over the `extern "C"` entry points of include/pcnerf_b200.h (through `pcnerf_b200.ops`): callers that prefer
dispatcher ops (TorchScript-free export, `torch.library` tooling, op-level profiling by name) get the same kernels as the
reference-shaped Python API.

    import pcnerf_b200.torch_ops                   # registers the library once
    z, enc = torch.ops.pcnerf.sample_encode_coarse(rays, 57, 7, 6, 7, 1.0, U, True)
"""
import torch

from . import ops

_LIB = torch.library.Library("pcnerf", "DEF")


def _def(schema, fn):
    name = schema.split("(")[0]
    _LIB.define(schema)
    _LIB.impl(name, fn, "CUDA")


# K1 -- nof/dataset/ipb2dmapping.py:367-452 (variant 406 MaiCity / 606 KITTI); parent = x_min,x_max,y_min,y_max,z_min,z_max
def _aabb_pack_train(variant, ray_o, ray_d, dist, points, centres, boxes, boxes_bigger, parent, surface_expand, knn):
    rays, keep = ops.aabb_pack_train(variant, ray_o, ray_d, dist, points, centres, boxes, boxes_bigger, tuple(parent),
                                     surface_expand, knn, compact=False)
    return rays, keep


_def("aabb_pack_train(int variant, Tensor ray_o, Tensor ray_d, Tensor dist, Tensor points, Tensor centres, Tensor boxes, "
     "Tensor boxes_bigger, float[] parent, float surface_expand, int knn) -> (Tensor, Tensor)", _aabb_pack_train)


# K1 -- eval_kitti_render.py:353-461 / :681-803
def _aabb_build_groups(ray_o, ray_d, dist, boxes, boxes_larger, parent_min, parent_max, method, grow_step, prefilter):
    rays, ranges, other, kept = ops.aabb_build_groups(ray_o, ray_d, dist, boxes, boxes_larger, list(parent_min),
                                                      list(parent_max), method, grow_step, prefilter)
    return rays, ranges, other, kept


_def("aabb_build_groups(Tensor ray_o, Tensor ray_d, Tensor dist, Tensor boxes, Tensor boxes_larger, float[] parent_min, "
     "float[] parent_max, int method, float grow_step, float prefilter) -> (Tensor, Tensor, Tensor, Tensor)", _aabb_build_groups)


# K2 -- nof/render.py:429-458 + nof/networks/models.py:27-41
def _sample_encode_coarse(rays, n_a, n_b, near_col, far_col, perturb, U, f16):
    return ops.sample_encode_coarse(rays, n_a, n_b, near_col, far_col, 10, 11, False, perturb, U, True, f16)


_def("sample_encode_coarse(Tensor rays, int n_a, int n_b, int near_col, int far_col, float perturb, Tensor? U, bool f16) "
     "-> (Tensor, Tensor)", _sample_encode_coarse)


# K2' -- nof/render.py:371-412, :463-468
def _sample_encode_fine(rays, z, w, n_importance, u, f16):
    return ops.sample_encode_fine(rays, z, w, n_importance, u, u is None, True, f16)


_def("sample_encode_fine(Tensor rays, Tensor z, Tensor w, int n_importance, Tensor? u, bool f16) -> (Tensor, Tensor)",
     _sample_encode_fine)


# K3 -- nof/networks/models.py:183-203 in eval mode (running statistics); params / buffers as NOF.kernel_params() / _buffers3()
def _mlp_eval(enc, params, running_mean, running_var, num_batches_tracked, precision):
    with torch.no_grad():
        return ops.MLPFunction.apply(enc.contiguous(), enc.shape[0], False, precision,
                                     (list(running_mean), list(running_var), list(num_batches_tracked)), None, *params)


_def("mlp_eval(Tensor enc, Tensor[] params, Tensor[] running_mean, Tensor[] running_var, Tensor[] num_batches_tracked, "
     "int precision) -> Tensor", _mlp_eval)


# K4 -- nof/render.py:51-61, :75-161 (+ train_kitti.py:145-146 with flags & 4); forward and backward as separate ops
def _composite_fwd(p, z, rays, cnear_col, cfar_col, range_col, epsilon, flags):
    with torch.no_grad():
        w, depth, fl, dl, _, _, opacity, rsl = ops.composite(p, z, rays, (cnear_col, cfar_col, range_col), None, 0.0,
                                                             epsilon, flags)
    return w, depth, torch.stack([fl, dl, rsl, opacity.to(fl.dtype)])


_def("composite_fwd(Tensor p, Tensor z, Tensor? rays, int cnear_col, int cfar_col, int range_col, float epsilon, int flags) "
     "-> (Tensor, Tensor, Tensor)", _composite_fwd)


# K5 -- nof/render.py:229-368
def _search_rows(p, z, rays, cnear_col, cfar_col, epsilon, method, row_ray):
    depth, w, opacity, peak, wsum = ops.search_rows(p, z, rays, cnear_col, cfar_col, epsilon, method, row_ray=row_ray)
    return depth, w, opacity, peak, wsum


_def("search_rows(Tensor p, Tensor z, Tensor rays, int cnear_col, int cfar_col, float epsilon, int method, Tensor? row_ray) "
     "-> (Tensor, Tensor, Tensor, Tensor, Tensor)", _search_rows)
_def("search_select(Tensor other, Tensor peak, Tensor wsum) -> Tensor", lambda other, peak, wsum: ops.search_select(other, peak, wsum))
_def("points(Tensor rays, Tensor depth) -> Tensor", lambda rays, depth: ops.points(rays, depth))

OPS = ("aabb_pack_train", "aabb_build_groups", "sample_encode_coarse", "sample_encode_fine", "mlp_eval", "composite_fwd",
       "search_rows", "search_select", "points")
