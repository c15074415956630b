"""ctypes binding of libpcnerf_b200.so (the C ABI declared in include/pcnerf_b200.h).

There is no fallback: if the library is missing or a kernel fails, the call raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpcnerf_b200.so")

NLIN, NBN = 9, 8
ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3

vp, ci, i64, f32, f64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_size_t


class MlpParams(ctypes.Structure):
    _fields_ = [("W", vp * NLIN), ("b", vp * NLIN), ("gamma", vp * NBN), ("beta", vp * NBN),
                ("running_mean", vp * NBN), ("running_var", vp * NBN), ("num_batches_tracked", vp * NBN),
                ("momentum", f32), ("eps", f32), ("training", ci), ("precision", ci), ("prepared", ci)]


class MlpGrads(ctypes.Structure):
    _fields_ = [("dW", vp * NLIN), ("db", vp * NLIN), ("dgamma", vp * NBN), ("dbeta", vp * NBN)]


PD = ctypes.POINTER(f64)

SIGNATURES = {
    "pcnerf_version": (ci, []),
    "pcnerf_last_error": (ctypes.c_char_p, []),
    "pcnerf_launch_count": (ctypes.c_longlong, [ci]),
    "pcnerf_prof_enable": (None, [ci]),
    "pcnerf_prof_classes": (ci, []),
    "pcnerf_prof_name": (ctypes.c_char_p, [ci]),
    "pcnerf_prof_read": (ci, [ci, ctypes.POINTER(f64), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(f64)]),
    "pcnerf_aabb_far_bound": (ci, [vp, vp, i64, PD, vp, vp]),
    "pcnerf_aabb_slab": (ci, [vp, vp, i64, PD, PD, vp, vp]),
    "pcnerf_aabb_child_pairs": (ci, [ci, vp, vp, i64, vp, ci, vp, vp, vp, vp]),
    "pcnerf_aabb_dist_to_ray": (ci, [vp, vp, i64, vp, ci, vp, vp]),
    "pcnerf_aabb_find_box": (ci, [vp, vp, ci, vp, i64, ci, vp, vp]),
    "pcnerf_aabb_pack_train": (ci, [ci, vp, vp, vp, vp, i64, vp, vp, vp, ci, PD, f64, ci, vp, vp, vp]),
    "pcnerf_aabb_groups_count": (ci, [vp, vp, i64, vp, vp, ci, PD, PD, ci, f64, f64, PD, vp, vp, vp, vp, vp]),
    "pcnerf_aabb_groups_fill": (ci, [vp, vp, vp, i64, vp, vp, ci, ci, f64, f64, PD, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pcnerf_sample_encode_coarse": (ci, [vp, ci, i64, ci, ci, ci, ci, vp, ci, vp, ci, ci, f32, vp, vp, vp, vp, vp]),
    "pcnerf_sample_encode_fine": (ci, [vp, ci, i64, vp, vp, ci, vp, ci, ci, vp, vp, vp, vp]),
    "pcnerf_sample_pdf": (ci, [vp, vp, i64, ci, vp, ci, ci, vp, vp]),
    "pcnerf_embed": (ci, [vp, i64, vp, ci, vp]),
    "pcnerf_mlp_saved_bytes": (sz, [i64, ci]),
    "pcnerf_mlp_scratch_bytes": (sz, [i64, ci]),
    "pcnerf_mlp_forward": (ci, [ctypes.POINTER(MlpParams), vp, i64, vp, vp, sz, vp, sz, vp]),
    "pcnerf_mlp_backward": (ci, [ctypes.POINTER(MlpParams), ctypes.POINTER(MlpGrads), vp, i64, vp, vp, vp, sz, vp, sz, vp]),
    "pcnerf_affine_parts": (ci, []),
    "pcnerf_affine_moments": (ci, [vp, i64, i64, vp, vp]),
    "pcnerf_affine_apply": (ci, [vp, i64, i64, vp, vp, vp, vp]),
    "pcnerf_affine_grad": (ci, [vp, vp, vp, i64, i64, vp, vp]),
    "pcnerf_affine_work_bytes": (sz, [i64]),
    "pcnerf_affine_forward_rays": (ci, [ctypes.POINTER(MlpParams), vp, ci, i64, vp, ci, i64, vp, vp, sz, vp]),
    "pcnerf_affine_backward_rays": (ci, [ctypes.POINTER(MlpParams), ctypes.POINTER(MlpGrads), vp, ci, i64, vp, ci, i64, vp, vp,
                                        vp, sz, vp]),
    "pcnerf_affine_eval_alpha": (ci, [ctypes.POINTER(MlpParams), vp, vp, sz, vp]),
    "pcnerf_affine_apply_rays": (ci, [vp, ci, i64, vp, ci, vp, vp, vp]),
    "pcnerf_tc_rowgemm_work_bytes": (sz, []),
    "pcnerf_tc_rowgemm": (ci, [ci, vp, ci, vp, ci, vp, vp, vp, i64, vp, vp, vp, vp, vp]),
    "pcnerf_tc_wgrad": (ci, [vp, vp, ci, ci, ci, i64, vp, ci, ci, vp]),
    "pcnerf_tc_last_fault": (ci, []),
    "pcnerf_mlp_tc_forward_chunks": (ci, [ctypes.POINTER(MlpParams), vp, i64, i64, vp, vp, vp, vp, ctypes.c_size_t, ci, vp]),
    "pcnerf_mlp_tc_backward_chunks": (ci, [ctypes.POINTER(MlpParams), ctypes.POINTER(MlpGrads), vp, i64, i64, vp, vp, vp, vp,
                                          vp, ctypes.c_size_t, ci, vp]),
    "pcnerf_tc_set_fused_eval": (None, [ci]),
    "pcnerf_tc_get_fused_eval": (ci, []),
    "pcnerf_tc_set_row_pairs": (None, [ci]),
    "pcnerf_tc_get_row_pairs": (ci, []),
    "pcnerf_tc_set_weight_correction": (None, [ci]),
    "pcnerf_tc_get_weight_correction": (ci, []),
    "pcnerf_composite_fwd": (ci, [vp, vp, vp, ci, i64, ci, ci, ci, ci, vp, f32, f32, ci, vp, vp, vp, vp, vp, vp]),
    "pcnerf_composite_losses": (ci, [vp, i64, vp, vp]),
    "pcnerf_composite_bwd": (ci, [vp, vp, vp, vp, ci, i64, ci, ci, f32, f32, ci, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp]),
    "pcnerf_search_rows": (ci, [vp, vp, vp, ci, i64, ci, ci, ci, f32, ci, vp, vp, vp, vp, vp, vp]),
    "pcnerf_search_rows_grouped": (ci, [vp, vp, vp, ci, i64, ci, ci, ci, f32, ci, vp, vp, vp, vp, vp, vp, vp]),
    "pcnerf_search_select": (ci, [vp, vp, vp, i64, vp, vp, vp]),
    "pcnerf_masked_loss_fwd": (ci, [ci, vp, vp, vp, i64, vp, vp, vp]),
    "pcnerf_masked_loss_bwd": (ci, [ci, vp, vp, vp, i64, vp, vp, vp, vp, vp]),
    "pcnerf_route_points": (ci, [vp, i64, vp, ci, vp, vp, vp, vp, vp, vp, vp]),
    "pcnerf_group_heads": (ci, [vp, i64, vp, vp]),
    "pcnerf_group_uniform": (ci, [vp, ci, i64, vp, vp, ci, ci, vp, vp]),
    "pcnerf_eval_rows_rendered": (ci, [vp, ci, ci, i64, i64, vp, vp]),
    "pcnerf_points": (ci, [vp, ci, i64, vp, vp, vp]),
    "pcnerf_adam_step": (ci, [vp, vp, vp, vp, i64, vp, vp, vp, f32, f32, f32, f32, f32, vp]),
    "pcnerf_frame_returns": (ci, [vp, i64, PD, vp, ci, f32, f32, f32, f32, f32, f32, f32, f32, PD, PD, vp, vp, vp, vp, vp]),
    "pcnerf_nn_correspondance": (ci, [vp, i64, vp, i64, vp, vp, vp]),
    "pcnerf_dist_stats": (ci, [vp, i64, f64, vp, vp]),
}

_lib = None


def header_symbols():
    """Entry points declared in include/pcnerf_b200.h (parsed, so tests can check header <-> library <-> binding)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "pcnerf_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(pcnerf_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libpcnerf_b200.so is not built (%s missing): run `python -m pcnerf_b200.build`; "
                               "there is no CPU fallback" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc == 0:
        return
    msg = lib().pcnerf_last_error().decode("utf-8", "replace")
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
