"""ctypes binding of libha_b200.so (the C ABI declared in include/ha_b200.h).

The library is the product: there is no Python/CPU fallback.  If it is missing or cannot be
loaded, importing this module's `lib()` raises — loudly — instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libha_b200.so")

HA_MAX_LEVELS = 4
HA_STATS = 24
HA_VGG_N_CONV = 17
HA_GEOM_KITTI, HA_GEOM_FORD, HA_GEOM_G2SP, HA_GEOM_G2SP_NN = 0, 1, 2, 3
HA_OPT_LM, HA_OPT_SGD, HA_OPT_ADAM, HA_OPT_GN = 0, 1, 2, 3
HA_CONV_FP32_SIMT, HA_CONV_F16X3, HA_CONV_F16, HA_CONV_F16X3_1CTA = 0, 1, 2, 3
HA_STATUS_NO_INRANGE, HA_STATUS_NAN_POSE, HA_STATUS_RESET, HA_STATUS_SAMPLE_EMPTY, HA_STATUS_TIMEOUT = 1, 2, 4, 8, 16
HA_COMM_ID_BYTES = 128
HA_ABI_VERSION = 3
STAT_H, STAT_GRAD, STAT_SAT_NORM, STAT_GRD_NORM, STAT_RES_SQ, STAT_DELTA, STAT_N_INRANGE = 0, 9, 12, 13, 14, 15, 18
STAT_JTG, STAT_RESET_MASK = 19, 22

# every symbol include/ha_b200.h declares (tests check the .so exports all of them)
EXPORTS = ["ha_version", "ha_error_string", "ha_last_cuda_error", "ha_device_check", "ha_nchw_to_nhwc",
           "ha_nhwc_to_nchw", "ha_lm_workspace_bytes", "ha_lm_step", "ha_lm_run", "ha_vgg_packed_weight_bytes",
           "ha_vgg_pack_weights", "ha_vgg_workspace_bytes", "ha_vgg_forward", "ha_conv3x3_workspace_bytes",
           "ha_conv3x3_nhwc", "ha_launch_count", "ha_comm_unique_id", "ha_comm_init", "ha_comm_destroy",
           "ha_pose_allgather", "ha_lm_backward_workspace_bytes", "ha_lm_step_backward",
           "ha_pose_loss", "ha_pose_loss_backward", "ha_img_affine_u8", "ha_img_resize_workspace_bytes",
           "ha_img_resize_to_tensor", "ha_vgg_train_workspace_bytes", "ha_vgg_forward_train",
           "ha_vgg_backward_workspace_bytes", "ha_vgg_backward", "ha_conv3x3_backward_workspace_bytes",
           "ha_conv3x3_backward_nhwc", "ha_lm_residual", "ha_nn_pose_update", "ha_vgg_g2s_forward"]


class HaLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("scale", C.c_void_p), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class HaLmParams(C.Structure):
    _fields_ = [("geometry", C.c_int32), ("n_levels", C.c_int32), ("n_iters", C.c_int32), ("level_first", C.c_int32),
                ("dof", C.c_int32), ("using_weight", C.c_int32), ("use_hessian", C.c_int32), ("batch", C.c_int32),
                ("rotation_range", C.c_float), ("shift_range_lat", C.c_float), ("shift_range_lon", C.c_float),
                ("damping", C.c_float * 3), ("meter_per_pixel", C.c_float * HA_MAX_LEVELS),
                ("inv_meter_per_pixel", C.c_float * HA_MAX_LEVELS), ("sat_center", C.c_float * HA_MAX_LEVELS),
                ("ori_grd_h", C.c_int32), ("ori_grd_w", C.c_int32), ("kernel_variant", C.c_int32), ("optimizer", C.c_int32),
                ("full_height", C.c_int32), ("adam_level_mult", C.c_int32), ("adam_iter", C.c_int32), ("adam_beta1", C.c_float),
                ("adam_beta2", C.c_float), ("reserved", C.c_int32)]


class HaVggStateDict(C.Structure):
    _fields_ = [("weight", C.c_void_p * HA_VGG_N_CONV), ("bias", C.c_void_p * HA_VGG_N_CONV)]


class HaVggGrads(C.Structure):
    _fields_ = [("weight", C.c_void_p * HA_VGG_N_CONV), ("bias", C.c_void_p * HA_VGG_N_CONV)]


class HaError(RuntimeError):
    pass


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise HaError("libha_b200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `python highlyaccurate_b200/build.py`.  There is no fallback path." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.ha_version.restype = i32
    L.ha_error_string.restype = C.c_char_p
    L.ha_error_string.argtypes = [i32]
    L.ha_last_cuda_error.restype = C.c_char_p
    L.ha_device_check.argtypes = [i32]
    L.ha_launch_count.restype = C.c_ulonglong
    for f in (L.ha_nchw_to_nhwc, L.ha_nhwc_to_nchw):
        f.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.ha_lm_workspace_bytes.restype = sz
    L.ha_lm_workspace_bytes.argtypes = [i32]
    L.ha_lm_step.argtypes = [C.POINTER(HaLmParams), i32, C.POINTER(HaLevel), C.POINTER(HaLevel), vp, vp, vp, vp, vp,
                             vp, vp, vp, sz, vp]
    L.ha_lm_run.argtypes = [C.POINTER(HaLmParams), C.POINTER(HaLevel), C.POINTER(HaLevel), C.POINTER(vp),
                            C.POINTER(vp), vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ha_lm_residual.argtypes = [C.POINTER(HaLmParams), i32, C.POINTER(HaLevel), C.POINTER(HaLevel), vp, vp, vp, i32, vp, vp]
    L.ha_nn_pose_update.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.ha_vgg_packed_weight_bytes.restype = sz
    L.ha_vgg_pack_weights.argtypes = [C.POINTER(HaVggStateDict), vp, sz, vp]
    L.ha_vgg_workspace_bytes.restype = sz
    L.ha_vgg_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.ha_vgg_forward.argtypes = [vp, vp, i32, i32, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, sz, vp]
    L.ha_vgg_train_workspace_bytes.restype = sz
    L.ha_vgg_train_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.ha_vgg_forward_train.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, sz, vp]
    L.ha_vgg_backward_workspace_bytes.restype = sz
    L.ha_vgg_backward_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.ha_vgg_backward.argtypes = [C.POINTER(HaVggStateDict), vp, i32, i32, i32, i32, vp, C.POINTER(vp), C.POINTER(HaVggGrads),
                                  vp, sz, vp]
    L.ha_conv3x3_backward_workspace_bytes.restype = sz
    L.ha_conv3x3_backward_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.ha_conv3x3_backward_nhwc.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    L.ha_vgg_g2s_forward.argtypes = [vp, vp, i32, i32, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, sz, vp]
    L.ha_conv3x3_workspace_bytes.restype = sz
    L.ha_conv3x3_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.ha_conv3x3_nhwc.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]
    L.ha_lm_backward_workspace_bytes.restype = sz
    L.ha_lm_backward_workspace_bytes.argtypes = [i32]
    L.ha_lm_step_backward.argtypes = [C.POINTER(HaLmParams), i32, C.POINTER(HaLevel), C.POINTER(HaLevel), vp, vp, vp, vp, vp,
                                      vp, vp, vp, vp, vp, sz, vp]
    L.ha_pose_loss.argtypes = [vp, vp, i32, i32, i32, C.POINTER(C.c_float), vp, vp, vp]
    L.ha_pose_loss_backward.argtypes = [vp, vp, i32, i32, i32, C.POINTER(C.c_float), vp, vp, vp, vp]
    L.ha_img_affine_u8.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp]
    L.ha_img_resize_workspace_bytes.restype = sz
    L.ha_img_resize_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.ha_img_resize_to_tensor.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.ha_comm_unique_id.argtypes = [vp]
    L.ha_comm_init.argtypes = [C.POINTER(vp), i32, i32, vp, i32]
    L.ha_comm_destroy.argtypes = [vp]
    L.ha_pose_allgather.argtypes = [vp, vp, vp, i32, vp]
    if L.ha_version() != HA_ABI_VERSION:
        raise HaError("libha_b200.so has ABI version %d, this package needs %d: rebuild it" % (L.ha_version(), HA_ABI_VERSION))
    for name in EXPORTS:
        f = getattr(L, name)
        if f.restype is C.c_int and name not in ("ha_version",):
            f.restype = i32
    _LIB = L
    return L


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    L = lib()
    msg = L.ha_error_string(rc).decode()
    if rc in (-3, -5):
        msg += ": " + L.ha_last_cuda_error().decode()
    raise HaError("%s failed: %s" % (what, msg))
