"""ctypes binding of ``libotvm_sm100.so`` (the C ABI declared in ``include/otvm_b200.h``).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libotvm_sm100.so")

F32, BF16, BF16X2, BF16X3 = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

c_i32, c_i64, c_vp, c_f = C.c_int32, C.c_int64, C.c_void_p, C.c_float


class ConvParams(C.Structure):
    _fields_ = [
        ("inp", c_vp), ("in_ld", c_i64),
        ("N", c_i32), ("H", c_i32), ("W", c_i32), ("Cin", c_i32),
        ("weight", c_vp), ("bias", c_vp),
        ("Cout", c_i32), ("KH", c_i32), ("KW", c_i32), ("stride", c_i32), ("pad", c_i32), ("dil", c_i32),
        ("out", c_vp), ("out_ps", c_i64), ("out_cs", c_i64),
        ("res", c_vp), ("res_ld", c_i64),
        ("out_relu", c_vp), ("out_relu_ld", c_i64),
        ("act", c_i32), ("relu_in", c_i32), ("dtype", c_i32), ("out_f32", c_i32),
        ("gn_stats", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
        ("gn_stats_zeroed", c_i32), ("gn_eps", c_f),
        ("gn_gamma", c_vp), ("gn_beta", c_vp),
        ("w_plane_stride", c_i64),
        ("gn_group_ch", c_i32), ("groups", c_i32),
    ]


class ReadParams(C.Structure):
    _fields_ = [
        ("keys", c_vp), ("vals", c_vp), ("ldv", c_i64),
        ("query", c_vp), ("q_ld", c_i64),
        ("out", c_vp), ("out_ld", c_i64),
        ("M", c_i32), ("HW", c_i32), ("De", c_i32), ("Do", c_i32),
        ("dtype", c_i32),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
        ("force_simt", c_i32),
        ("lse", c_vp),
    ]


class ReadBwdParams(C.Structure):
    _fields_ = [
        ("keys", c_vp), ("vals", c_vp), ("ldv", c_i64),
        ("query", c_vp), ("q_ld", c_i64),
        ("out", c_vp), ("out_ld", c_i64),
        ("dout", c_vp), ("dout_ld", c_i64),
        ("lse", c_vp),
        ("dkeys", c_vp), ("dvals", c_vp), ("dldv", c_i64), ("dquery", c_vp),
        ("M", c_i32), ("HW", c_i32), ("De", c_i32), ("Do", c_i32),
        ("dtype", c_i32),
    ]


# name -> (restype, argtypes); must list every symbol include/otvm_b200.h declares
SIGNATURES = {
    "otvm_version": (C.c_int, []),
    "otvm_strerror": (C.c_char_p, [C.c_int]),
    "otvm_last_cuda_error": (C.c_char_p, []),
    "otvm_launch_count": (c_i64, []),
    "otvm_device_is_sm100": (C.c_int, [C.c_int]),
    "otvm_set_pdl": (None, [C.c_int]),
    "otvm_device_error_flags": (C.c_int, [C.c_int]),
    "otvm_zero_async": (C.c_int, [c_vp, c_i64, c_vp]),
    "otvm_conv2d": (C.c_int, [C.POINTER(ConvParams), c_vp]),
    "otvm_conv2d_uses_tensor_cores": (C.c_int, [C.POINTER(ConvParams)]),
    "otvm_conv2d_can_fuse_gn": (C.c_int, [C.POINTER(ConvParams)]),
    "otvm_gn_stats": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "otvm_gn_apply": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_f,
                                c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "otvm_upsample_bilinear": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64,
                                         c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_vp]),
    "otvm_maxpool3x3s2": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_i32, c_vp]),
    "otvm_ppm_pool": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp]),
    "otvm_memory_read_workspace": (c_i64, [c_i32, c_i32, c_i32, c_i32, c_i32]),
    "otvm_memory_read": (C.c_int, [C.POINTER(ReadParams), c_vp]),
    "otvm_memory_read_backward": (C.c_int, [C.POINTER(ReadBwdParams), c_vp]),
    "otvm_preprocess": (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                  C.POINTER(c_f), c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "otvm_trimap_encode": (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, C.POINTER(c_f), c_vp, c_i64,
                                     c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "otvm_edt_sq": (C.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "otvm_fba_head": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "otvm_head_conv_fba": (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "otvm_frame_outputs": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                                     c_i32, C.POINTER(c_f), c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "otvm_unpack_frame_u8": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "otvm_alpha_to_u8": (C.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "otvm_nchw_to_nhwc": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_i32, c_vp]),
    "otvm_nhwc_to_nchw": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp]),
}

_lib = None


class OtvmError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built: ``python -m otvm_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OtvmError(f"{LIB_PATH} not found - build it with `python -m otvm_b200.build` "
                        "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype, fn.argtypes = res, args
    if lib.otvm_version() != 5:
        raise OtvmError("ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        lib = load()
        msg = lib.otvm_strerror(rc).decode()
        if rc == -2:
            msg += ": " + lib.otvm_last_cuda_error().decode()
        raise OtvmError(f"{what} failed ({rc}): {msg}")
