"""ctypes binding of include/handmvnet_b200.h.  Fails loudly when the CUDA library is missing."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HMV_LIB_PATH") or os.path.join(HERE, "lib", "libhandmvnet_b200.so")   # HMV_LIB_PATH: A/B builds (tools/)

PRECISION = {"bf16": 0, "fp32": 1}
STAGE = {"backbone": 0, "pose": 1, "sample": 2, "fusion": 3, "gcn": 4, "softargmax": 5}
TENSOR = {"feat": 0, "heatmap": 1, "xy": 2, "tokens": 3, "fused": 4, "joints": 5, "feat1": 6, "feat2": 7, "feat3": 8}
BACKBONE = {"resnet": 0, "hrnet": 1}

# every symbol include/handmvnet_b200.h declares
EXPORTS = ["hmv_create", "hmv_destroy", "hmv_set_weight", "hmv_prepare", "hmv_forward", "hmv_forward_host",
           "hmv_forward_host_async", "hmv_host_wait", "hmv_set_input_norm", "hmv_preprocess", "hmv_forward_u8", "hmv_forward_host_u8_async",
           "hmv_synchronize", "hmv_stage_run", "hmv_tensor_get", "hmv_tensor_set", "hmv_debug_backbone",
           "hmv_debug_num_steps", "hmv_debug_step_name", "hmv_debug_step_io", "hmv_debug_step_run", "hmv_conv_bn_act", "hmv_profile_enable", "hmv_profile_read", "hmv_profile_phases",
           "hmv_launch_count", "hmv_num_sms",
           "hmv_last_error", "hmv_version"]


class HmvConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("num_views", "image_size", "heatmap_size", "use_pos2d", "use_crop", "use_sin", "fusion_layers",
                 "precision", "micro_batch", "device", "backbone")] + [("hr_channels", ctypes.c_int32 * 4)]


_lib = None


def load():
    """Load libhandmvnet_b200.so (built by handmvnet_b200/build.py or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m handmvnet_b200.build` "
            "(handmvnet_b200 has no CPU / eager-PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
    lib.hmv_create.argtypes = [ctypes.POINTER(HmvConfig), ctypes.POINTER(vp)]
    lib.hmv_destroy.argtypes = [vp]
    lib.hmv_set_weight.argtypes = [vp, ctypes.c_char_p, f32p, ctypes.POINTER(i64), i32]
    lib.hmv_prepare.argtypes = [vp]
    lib.hmv_forward.argtypes = [vp, f32p, f32p, f32p, i32, f32p, f32p, f32p, vp]
    lib.hmv_forward_host.argtypes = [vp, f32p, f32p, f32p, i32, f32p, f32p, f32p]
    lib.hmv_forward_host_async.argtypes = [vp, f32p, f32p, f32p, i32, f32p, f32p, f32p, ctypes.POINTER(i64)]
    lib.hmv_host_wait.argtypes = [vp, i64]
    lib.hmv_set_input_norm.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    lib.hmv_preprocess.argtypes = [vp, vp, vp, i32, i32, i32, f32p, vp]
    lib.hmv_forward_u8.argtypes = [vp, vp, f32p, f32p, i32, f32p, f32p, f32p, vp]
    lib.hmv_forward_host_u8_async.argtypes = [vp, vp, f32p, f32p, i32, f32p, f32p, f32p, ctypes.POINTER(i64)]
    lib.hmv_synchronize.argtypes = [vp]
    lib.hmv_stage_run.argtypes = [vp, i32, f32p, f32p, f32p, i32, vp]
    lib.hmv_tensor_get.argtypes = [vp, i32, f32p, i32, vp]
    lib.hmv_tensor_set.argtypes = [vp, i32, f32p, i32, vp]
    lib.hmv_debug_backbone.argtypes = [vp, f32p, i32, i32, f32p, ctypes.POINTER(i32), vp]
    lib.hmv_debug_num_steps.argtypes = [vp]
    lib.hmv_debug_step_name.argtypes = [vp, i32]
    lib.hmv_debug_step_name.restype = ctypes.c_char_p
    lib.hmv_debug_step_io.argtypes = [vp, i32, ctypes.c_char_p, i32]
    lib.hmv_debug_step_run.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), vp]
    lib.hmv_conv_bn_act.argtypes = [i32, f32p, f32p, f32p, f32p, f32p, f32p, i32, i32, i32, i32, i32, i32, i32, i32,
                                    ctypes.POINTER(ctypes.c_float), i32, vp]
    lib.hmv_profile_enable.argtypes = [vp, i32]
    lib.hmv_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(i64), ctypes.c_char_p]
    lib.hmv_profile_phases.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.hmv_launch_count.argtypes = [vp]
    lib.hmv_launch_count.restype = i64
    lib.hmv_num_sms.argtypes = [vp]
    lib.hmv_last_error.restype = ctypes.c_char_p
    lib.hmv_version.restype = ctypes.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int:
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc, what=""):
    """C status -> Python exception (the C ABI never throws)."""
    if rc != 0:
        msg = load().hmv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"handmvnet_b200 {what} failed: {msg}")


def ptr(t):
    """Raw pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
