"""ctypes binding of ``libugl_b200.so`` (the C-ABI declared in ``include/ugl.h``).

The library is the product: there is no CPU or pure-PyTorch fallback.  ``lib()`` raises
``RuntimeError`` if the shared object has not been built (``python -m <pkg>.build``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

MAX_LEVELS = 6
FLOW_NSTATS = 12
GEOM_NSTATS = 16
FLOW_BASIS_PLANES = 14
# internal forms of the single-pass forward (include/ugl.h: UGL_SINGLE_PASS_*)
SINGLE_PASS_VARIANTS = {"fused": 0, "split": 1, "split_plain": 2, "split_tma": 3}
STEP_PARTS = {"photo": 1, "norm": 2, "stencil": 4, "finalize": 8, "no_finalize": 7, "all": 15}     # include/ugl.h: UGL_STEP_*
DEPTH_BASIS_PLANES = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UGL_LIB_PATH") or os.path.join(_HERE, "libugl_b200.so")   # override: tuning experiments only

_fp = C.POINTER(C.c_float)


class UglFlowLossArgs(C.Structure):
    """Mirror of ``struct UglFlowLossArgs`` (include/ugl.h)."""

    _fields_ = [
        ("batch", C.c_int32),
        ("levels", C.c_int32),
        ("scales", C.c_int32),
        ("height", C.c_int32 * MAX_LEVELS),
        ("width", C.c_int32 * MAX_LEVELS),
        ("img_l", C.c_void_p * MAX_LEVELS),
        ("img", C.c_void_p * MAX_LEVELS),
        ("img_r", C.c_void_p * MAX_LEVELS),
        ("flow_fwd", C.c_void_p * MAX_LEVELS),
        ("flow_bwd", C.c_void_p * MAX_LEVELS),
        ("loss", C.c_void_p),
        ("stats", C.c_void_p),
        ("grad_loss", C.c_void_p),
        ("grad_flow_fwd", C.c_void_p * MAX_LEVELS),
        ("grad_flow_bwd", C.c_void_p * MAX_LEVELS),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_uint64),
        ("stream", C.c_void_p),
        ("basis", C.c_void_p * MAX_LEVELS),
    ]


class UglDepthPhotoArgs(C.Structure):
    """Mirror of ``struct UglDepthPhotoArgs`` (include/ugl.h)."""

    _L, _L2 = C.c_void_p * MAX_LEVELS, (C.c_void_p * MAX_LEVELS) * 2
    _fields_ = [
        ("batch", C.c_int32), ("scales", C.c_int32),
        ("height", C.c_int32 * MAX_LEVELS), ("width", C.c_int32 * MAX_LEVELS),
        ("img", _L), ("src_area", _L2), ("src_bil", _L2), ("disp", _L), ("Kinv", _L), ("P", _L2), ("ext_mask", _L2),
        ("valid_out", _L2), ("tex_out", _L2),
        ("loss", C.c_void_p), ("den", C.c_void_p), ("grad_loss", C.c_void_p),
        ("grad_disp", _L), ("grad_P", _L2),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("stream", C.c_void_p),
        ("ext_bytes", _L), ("ext_need", C.c_int32 * 2),
    ]


class UglDepthConsisArgs(C.Structure):
    """Mirror of ``struct UglDepthConsisArgs`` (include/ugl.h)."""

    _L, _L2 = C.c_void_p * MAX_LEVELS, (C.c_void_p * MAX_LEVELS) * 2
    _fields_ = [("batch", C.c_int32), ("scales", C.c_int32), ("height", C.c_int32 * MAX_LEVELS), ("width", C.c_int32 * MAX_LEVELS),
                ("disp", _L), ("ref_disp", _L2), ("Kinv", _L), ("P", _L2), ("loss", C.c_void_p), ("grad_loss", C.c_void_p),
                ("grad_disp", _L), ("grad_ref", _L2), ("grad_P", _L2),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("stream", C.c_void_p)]


class UglDepthSsimArgs(C.Structure):
    """Mirror of ``struct UglDepthSsimArgs`` (include/ugl.h)."""

    _fields_ = [("photo", UglDepthPhotoArgs), ("loss4", C.c_void_p), ("stats", C.c_void_p), ("basis", C.c_void_p * MAX_LEVELS),
                ("grad_loss4", C.c_void_p)]


class UglDepthPhotoGradArgs(C.Structure):
    """Mirror of ``struct UglDepthPhotoGradArgs`` (include/ugl.h)."""

    _fields_ = [("photo", UglDepthPhotoArgs), ("basis", C.c_void_p * MAX_LEVELS), ("psum", C.c_void_p)]


class UglGeomFlowArgs(C.Structure):
    """Mirror of ``struct UglGeomFlowArgs`` (include/ugl.h)."""

    _L = C.c_void_p * MAX_LEVELS
    _fields_ = [
        ("flow", UglFlowLossArgs),
        ("disp", _L), ("Kinv", _L), ("P_bwd", _L), ("P_fwd", _L), ("mask_bytes", _L),
        ("alpha", C.c_float), ("beta", C.c_float),
    ]


class UglDispSmoothArgs(C.Structure):
    """Mirror of ``struct UglDispSmoothArgs`` (include/ugl.h)."""

    MAX_LISTS = 3
    _LL = (C.c_void_p * MAX_LEVELS) * MAX_LISTS
    _fields_ = [
        ("batch", C.c_int32), ("lists", C.c_int32), ("levels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("lheight", C.c_int32 * MAX_LEVELS), ("lwidth", C.c_int32 * MAX_LEVELS),
        ("img", C.c_void_p * MAX_LISTS), ("disp", _LL), ("out", C.c_void_p), ("G", _LL), ("grad_out", C.c_void_p), ("grad_disp", _LL),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("stream", C.c_void_p), ("grad_out_shared", C.c_int32),
    ]


class UglPyramidArgs(C.Structure):
    """Mirror of ``struct UglPyramidArgs`` (include/ugl.h)."""

    MAX_IMAGES = 3
    _LL = (C.c_void_p * MAX_LEVELS) * MAX_IMAGES
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("levels", C.c_int32),
                ("images", C.c_int32), ("img", C.c_void_p * MAX_IMAGES), ("box", _LL), ("bil", _LL), ("stream", C.c_void_p)]


class UglGeomRigidArgs(C.Structure):
    """Mirror of ``struct UglGeomRigidArgs`` (include/ugl.h)."""

    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("need", C.c_int32 * 2)] + [
        (name, C.c_void_p) for name in (
            "flow_bwd", "flow_fwd", "disp", "mask_bytes", "Kinv", "P_bwd", "P_fwd", "F_bwd", "F_fwd", "loss_dfc", "loss_epi", "den",
            "grad_dfc", "grad_epi", "grad_flow_bwd", "grad_flow_fwd", "grad_disp", "grad_P_bwd", "grad_P_fwd", "grad_F_bwd", "grad_F_fwd",
            "workspace")] + [("workspace_bytes", C.c_uint64), ("stream", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/ugl.h declares
SIGNATURES = {
    "ugl_version": (C.c_int, []),
    "ugl_last_error": (C.c_char_p, []),
    "ugl_flow_loss_workspace_bytes": (C.c_uint64, [C.POINTER(UglFlowLossArgs)]),
    "ugl_flow_loss_forward": (C.c_int, [C.POINTER(UglFlowLossArgs)]),
    "ugl_flow_loss_backward": (C.c_int, [C.POINTER(UglFlowLossArgs)]),
    "ugl_flow_loss_launches": (C.c_int, [C.c_int]),
    "ugl_depth_consis_workspace_bytes": (C.c_uint64, [C.POINTER(UglDepthConsisArgs)]),
    "ugl_depth_consis_forward": (C.c_int, [C.POINTER(UglDepthConsisArgs)]),
    "ugl_depth_consis_backward": (C.c_int, [C.POINTER(UglDepthConsisArgs)]),
    "ugl_depth_ssim_workspace_bytes": (C.c_uint64, [C.POINTER(UglDepthSsimArgs)]),
    "ugl_depth_ssim_forward_grad": (C.c_int, [C.POINTER(UglDepthSsimArgs)]),
    "ugl_depth_ssim_combine": (C.c_int, [C.POINTER(UglDepthSsimArgs)]),
    "ugl_geom_flow_forward_grad": (C.c_int, [C.POINTER(UglGeomFlowArgs)]),
    "ugl_geom_flow_forward_grad_ex": (C.c_int, [C.POINTER(UglGeomFlowArgs), C.c_int32]),
    "ugl_geom_flow_combine": (C.c_int, [C.POINTER(UglGeomFlowArgs)]),
    "ugl_geom_flow_step": (C.c_int, [C.POINTER(UglGeomFlowArgs)]),
    "ugl_geom_flow_step_parts": (C.c_int, [C.POINTER(UglGeomFlowArgs), C.c_int32]),
    "ugl_flow_loss_forward_grad": (C.c_int, [C.POINTER(UglFlowLossArgs)]),
    "ugl_flow_loss_forward_grad_ex": (C.c_int, [C.POINTER(UglFlowLossArgs), C.c_int32]),
    "ugl_flow_loss_step": (C.c_int, [C.POINTER(UglFlowLossArgs)]),
    "ugl_flow_loss_step_parts": (C.c_int, [C.POINTER(UglFlowLossArgs), C.c_int32]),
    "ugl_flow_loss_combine": (C.c_int, [C.POINTER(UglFlowLossArgs)]),
    "ugl_image_pyramid": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_void_p), C.c_void_p]),
    "ugl_warp_flow_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "ugl_warp_flow_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ugl_warp_flow_backward_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p]),
    "ugl_warp_flow_backward_workspace_bytes": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
}
SCATTER_FORMS = {"tile_local": 0, "global": 1}     # include/ugl.h: UGL_SCATTER_*

_i, _u64, _p, _f, _i64 = C.c_int32, C.c_uint64, C.c_void_p, C.c_float, C.c_int64
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int32)
SIGNATURES.update({
    "ugl_forward_splat_workspace_bytes": (_u64, [_i, _i, _i, _i]),
    "ugl_forward_splat": (C.c_int, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _u64, _p]),
    "ugl_forward_splat_ex": (C.c_int, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _u64, _i, _p]),
    "ugl_selftest_packed_pairs": (C.c_int, [_p, _i, _i, _p]),
    "ugl_frames_u8_to_float": (C.c_int, [_pp, _pp, _i, _u64, _p]),
    "ugl_depth_photo_workspace_bytes": (_u64, [C.POINTER(UglDepthPhotoArgs)]),
    "ugl_depth_photo_forward": (C.c_int, [C.POINTER(UglDepthPhotoArgs)]),
    "ugl_depth_photo_backward": (C.c_int, [C.POINTER(UglDepthPhotoArgs)]),
    "ugl_depth_photo_grad_workspace_bytes": (_u64, [C.POINTER(UglDepthPhotoGradArgs)]),
    "ugl_depth_photo_forward_grad": (C.c_int, [C.POINTER(UglDepthPhotoGradArgs)]),
    "ugl_depth_photo_combine": (C.c_int, [C.POINTER(UglDepthPhotoGradArgs)]),
    "ugl_reduce_workspace_bytes": (_u64, [_i, _i, _i]),
    "ugl_masked_mean_forward": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _u64, _p]),
    "ugl_masked_mean_backward": (C.c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ugl_occlusion_weights": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "ugl_channel_mean_abs_diff_backward": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "ugl_texture_mask": (C.c_int, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "ugl_dynamic_mask_forward": (C.c_int, [_p, _p, _i, _i, _i, _f, _f, _p, _p, _p, _p]),
    "ugl_abs_diff_backward": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "ugl_mask_product": (C.c_int, [_pp, _ip, _i, _i64, _p, _p]),
    "ugl_accumulate_multi": (C.c_int, [_pp, _pp, C.POINTER(C.c_int64), _i, _p]),
    "ugl_assemble_rows": (C.c_int, [_pp, _ip, _i, _i, _p, _p]),
    "ugl_weighted_total_forward": (C.c_int, [_p, _p, _i, _i, _p, _p]),
    "ugl_weighted_total_backward": (C.c_int, [_p, _p, _i, _i, _p, _p]),
    "ugl_rigid_mask": (C.c_int, [_p, _i64, _f, _f, _p, _p, _p, _p]),
    "ugl_flow_smooth_forward": (C.c_int, [_p, _p, _i, _i, _i, _p, _p, _u64, _p]),
    "ugl_flow_smooth_backward": (C.c_int, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "ugl_flow_consis_forward": (C.c_int, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _u64, _p]),
    "ugl_flow_consis_backward": (C.c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "ugl_depth_diff_forward": (C.c_int, [_p, _p, _i64, _p, _p]),
    "ugl_depth_diff_backward": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p]),
    "ugl_cost_volume_forward": (C.c_int, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "ugl_cost_volume_backward": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ugl_image_pyramid_multi": (C.c_int, [C.POINTER(UglPyramidArgs)]),
    "ugl_geom_rigid_workspace_bytes": (_u64, [_i, _i, _i]),
    "ugl_geom_rigid_forward": (C.c_int, [C.POINTER(UglGeomRigidArgs)]),
    "ugl_geom_rigid_backward": (C.c_int, [C.POINTER(UglGeomRigidArgs)]),
    "ugl_pose_setup_forward": (C.c_int, [_p, _p, _p, _fp, _i, _i, _i, _pp, _pp, _pp, _p]),
    "ugl_pose_setup_backward": (C.c_int, [_p, _p, _p, _fp, _i, _i, _i, _pp, _pp, _p, _p]),
    "ugl_disp_smooth_forward": (C.c_int, [_p, _pp, _ip, _ip, _i, _i, _i, _i, _p, _p, _u64, _p]),
    "ugl_disp_smooth_backward_workspace_bytes": (_u64, [_i, _i, _i]),
    "ugl_disp_smooth_fused_workspace_bytes": (_u64, [C.POINTER(UglDispSmoothArgs)]),
    "ugl_disp_smooth_forward_grad": (C.c_int, [C.POINTER(UglDispSmoothArgs)]),
    "ugl_disp_smooth_combine": (C.c_int, [C.POINTER(UglDispSmoothArgs)]),
    "ugl_disp_smooth_backward": (C.c_int, [_p, _pp, _ip, _ip, _i, _p, _i, _i, _i, _pp, _p, _u64, _p]),
    "ugl_reproject_forward": (C.c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ugl_reproject_backward_workspace_bytes": (_u64, [_i, _i, _i, _i, _i, _i]),
    "ugl_reproject_backward": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _u64, _p]),
    "ugl_rigid_flow_forward": (C.c_int, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "ugl_rigid_flow_backward": (C.c_int, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _u64, _p]),
    "ugl_epipolar_forward": (C.c_int, [_p, _p, _i, _i, _i, _p, _p]),
    "ugl_epipolar_backward": (C.c_int, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _u64, _p]),
    "ugl_ssim_forward": (C.c_int, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "ugl_ssim_backward": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "ugl_ssim_loss_workspace_bytes": (_u64, [_i, _i, _i, _i]),
    "ugl_ssim_loss_forward": (C.c_int, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _u64, _p]),
    "ugl_ssim_loss_backward": (C.c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
})

_lib: Optional[C.CDLL] = None


def bind(cdll: C.CDLL, signatures=SIGNATURES) -> C.CDLL:
    for name, (res, args) in signatures.items():
        fn = getattr(cdll, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return cdll


def lib() -> C.CDLL:
    """Load (once) and return the bound library; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libugl_b200.so is not built (%s). Run `python -m unsupervised_depth_opticalflow_egomotion_b200.build`; "
                "this package has no CPU / PyTorch fallback." % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


class UglError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = lib().ugl_last_error()
    text = msg.decode("utf-8", "replace") if msg else ""
    if rc in (-1, -4):
        raise ValueError("%s failed (%d): %s" % (what, rc, text))
    raise UglError("%s failed (%d): %s" % (what, rc, text))
