"""ctypes binding of libdfprior_b200.so (include/dfprior_b200.h).

There is no CPU fallback: if the CUDA library has not been built, ``lib()`` raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``python -m attentive_dfprior_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfprior_b200.so")

ABI_VERSION = 9
FLAG_SAVE_ACT = 1
FLAG_SIMT = 2
FLAG_FROZEN = (4, 8, 16)          # low, high, colour decoder: no parameter gradient wanted (DFP_FLAG_FROZEN_*)
MAX_RAY_SAMPLES = 64
N_PARAM_TENSORS = 79
N_PARAMS = 86029
STAGES = {"low": 0, "high": 1, "color": 2}

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class DfpGrid(C.Structure):
    _fields_ = [("data", C.c_void_p), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32)]


class DfpTsdf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("sD", C.c_int64), ("sH", C.c_int64), ("sW", C.c_int64)]


class DfpScene(C.Structure):
    _fields_ = [("low", DfpGrid), ("high", DfpGrid), ("color", DfpGrid), ("tsdf", DfpTsdf),
                ("bound", (C.c_double * 2) * 3), ("tsdf_bnds", (C.c_double * 2) * 3)]


class DfpRays(C.Structure):
    _fields_ = [("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("gt_depth", C.c_void_p),
                ("n_rays", C.c_int64), ("chunk", C.c_int64), ("depth_max", C.c_void_p)]


class DfpCfg(C.Structure):
    _fields_ = [("stage", C.c_int32), ("n_samples", C.c_int32), ("n_surface", C.c_int32), ("n_importance", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32),
                ("t_vals", C.c_float * MAX_RAY_SAMPLES), ("one_minus_t", C.c_float * MAX_RAY_SAMPLES),
                ("t_surf", C.c_double * MAX_RAY_SAMPLES), ("one_minus_t_surf", C.c_double * MAX_RAY_SAMPLES)]


class DfpRenderOut(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("uncertainty", C.c_void_p), ("color", C.c_void_p), ("weight", C.c_void_p),
                ("z_vals", C.c_void_p), ("raw", C.c_void_p)]


class DfpRenderOutGrad(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("uncertainty", C.c_void_p), ("color", C.c_void_p), ("weight", C.c_void_p)]


class DfpGrads(C.Structure):
    _fields_ = [("grid_low", C.c_void_p), ("grid_high", C.c_void_p), ("grid_color", C.c_void_p),
                ("params_packed", C.c_void_p), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("points", C.c_void_p),
                ("ev_high", C.c_void_p), ("ev_color", C.c_void_p)]


class DfpPoints(C.Structure):
    _fields_ = [("p", C.c_void_p), ("is_f64", C.c_int32), ("apply_bound_mask", C.c_int32), ("n", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class DfpAdam(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("voxel_mask", C.c_void_p), ("n_vox", C.c_int64), ("C", C.c_int32), ("step", C.c_int32),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double)]


class DfpFrustum(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("w2c", C.c_float * 16), ("cam_o", C.c_float * 3),
                ("xs", C.c_void_p), ("ys", C.c_void_p), ("zs", C.c_void_p),
                ("Dg", C.c_int32), ("Hg", C.c_int32), ("Wg", C.c_int32)]


class DfpFrame(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("color", C.c_void_p), ("c2w", C.c_float * 12)]


class DfpRayBatch(C.Structure):
    _fields_ = [("frames", C.POINTER(DfpFrame)), ("n_frames", C.c_int32), ("indices", C.c_void_p), ("n_per_frame", C.c_int32),
                ("H0", C.c_int32), ("H1", C.c_int32), ("W0", C.c_int32), ("W1", C.c_int32), ("W", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("bound", (C.c_double * 2) * 3), ("filter", C.c_int32)]


class DfpRayBatchOut(C.Structure):
    _fields_ = [("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("gt_depth", C.c_void_p), ("gt_color", C.c_void_p),
                ("n_kept", C.c_void_p)]


class DfpMapLoss(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("color", C.c_void_p), ("weight", C.c_void_p), ("gt_depth", C.c_void_p),
                ("gt_color", C.c_void_p), ("n", C.c_int64), ("S", C.c_int32), ("w_color", C.c_double),
                ("use_weight", C.c_int32), ("g_depth", C.c_void_p), ("g_color", C.c_void_p), ("g_weight", C.c_void_p),
                ("loss", C.c_void_p)]


class DfpTsdfFuse(C.Structure):
    _fields_ = [("tsdf_vol", C.c_void_p), ("weight_vol", C.c_void_p), ("color_vol", C.c_void_p), ("dim", C.c_int32 * 3),
                ("origin", C.c_float * 3), ("voxel_size", C.c_float), ("trunc_margin", C.c_float), ("depth_im", C.c_void_p),
                ("color_im", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("intr", C.c_float * 9), ("pose", C.c_float * 16),
                ("obs_weight", C.c_float)]


# every symbol include/dfprior_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "dfp_abi_version": (C.c_int, []),
    "dfp_last_error": (C.c_char_p, []),
    "dfp_packed_param_floats": (C.c_int64, []),
    "dfp_packed_grad_floats": (C.c_int64, []),
    "dfp_pack_params": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "dfp_unpack_param_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfp_grid_to_ndhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "dfp_grid_to_cdhw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "dfp_render_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32]),
    "dfp_render_fwd": (C.c_int, [C.POINTER(DfpScene), C.c_void_p, C.POINTER(DfpRays), C.POINTER(DfpCfg),
                                 C.POINTER(DfpRenderOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_render_bwd": (C.c_int, [C.POINTER(DfpScene), C.c_void_p, C.POINTER(DfpRays), C.POINTER(DfpCfg),
                                 C.POINTER(DfpRenderOutGrad), C.POINTER(DfpGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_points_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "dfp_points_fwd": (C.c_int, [C.POINTER(DfpScene), C.c_void_p, C.POINTER(DfpPoints), C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_points_bwd": (C.c_int, [C.POINTER(DfpScene), C.c_void_p, C.POINTER(DfpPoints), C.c_int32, C.c_void_p,
                                 C.c_void_p, C.POINTER(DfpGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_tsdf_sample": (C.c_int, [C.POINTER(DfpScene), C.POINTER(DfpPoints), C.c_void_p, C.c_void_p]),
    "dfp_gen_rays": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                               C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfp_masked_adam_step": (C.c_int, [C.POINTER(DfpAdam), C.c_void_p]),
    "dfp_frustum_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "dfp_frustum_mask": (C.c_int, [C.POINTER(DfpFrustum), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_assemble_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "dfp_assemble_rays": (C.c_int, [C.POINTER(DfpRayBatch), C.POINTER(DfpRayBatchOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "dfp_tsdf_integrate": (C.c_int, [C.POINTER(DfpTsdfFuse), C.c_void_p]),
    "dfp_point_masks": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfp_mapper_loss": (C.c_int, [C.POINTER(DfpMapLoss), C.c_void_p]),
    # include/dfprior_mesh.h
    "dfp_mc_count": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "dfp_mc_emit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfp_launch_count": (C.c_int64, []),
    "dfp_timing_enable": (C.c_int, [C.c_int]),
    "dfp_timing_collect": (C.c_int, [C.c_char_p, C.c_size_t]),
}

_lib = None
_lock = threading.Lock()


class DfpError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the CUDA library; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DfpError(
                    f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                    "(run `python -m attentive_dfprior_b200.build`). There is no CPU fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            got = handle.dfp_abi_version()
            if got != ABI_VERSION:
                raise DfpError(f"ABI mismatch: library {got}, binding {ABI_VERSION}")
            _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().dfp_last_error()
        raise DfpError(f"libdfprior_b200 error {rc}: {msg.decode() if msg else ''}")


def fill_bounds(dst, t):
    """Copy a [3,2] bound tensor (any device, f64) into a C double[3][2]."""
    vals = t.detach().to("cpu").double().tolist()
    for k in range(3):
        dst[k][0] = vals[k][0]
        dst[k][1] = vals[k][1]


def launch_count() -> int:
    return int(lib().dfp_launch_count())


def kernel_times(fn, *args, **kw):
    """Run ``fn`` with per-kernel CUDA-event timing on; returns {kernel: (launches, total_ms)}."""
    import torch
    l = lib()
    torch.cuda.synchronize()
    l.dfp_timing_enable(1)
    try:
        fn(*args, **kw)
        torch.cuda.synchronize()
    finally:
        l.dfp_timing_enable(0)
    buf = C.create_string_buffer(1 << 16)
    check(l.dfp_timing_collect(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
