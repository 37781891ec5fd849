"""ctypes binding of include/demfi_b200.h (the C-ABI shared library libdemfi_b200.so).

This is the stub a maintainer of the reference would add: plain
pointers and sizes only.  There is no fallback: if the library is missing or the device
is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEMFI_LIB") or os.path.join(_HERE, "libdemfi_b200.so")  # DEMFI_LIB: A/B measurement of two builds

MAX_SRC = 4
MAX_SEG = 4
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_SIGMOID_MUL, ACT_GRU = range(6)
STORE_NHWC, STORE_PIXEL_SHUFFLE2 = 0, 1
CONV_FFMA, CONV_TC, CONV_TC16, CONV_TC16W, CONV_TC16P = 0, 1, 2, 3, 4
FMT_F32, FMT_S16 = 0, 1
SEG_DST_S16, SEG_RES_S16, SEG_RES2_S16 = 1, 2, 4

i32 = C.c_int32
vp = C.c_void_p


class Src(C.Structure):
    _fields_ = [("ptr", vp), ("C", i32), ("ld", i32), ("up", i32), ("fmt", i32)]


class Seg(C.Structure):
    _fields_ = [("dst", vp), ("res", vp), ("res2", vp), ("dst_ld", i32), ("res_ld", i32), ("res2_ld", i32),
                ("ch0", i32), ("nch", i32), ("act", i32), ("store", i32), ("fmt", i32)]


class Part(C.Structure):
    _fields_ = [("src", vp), ("src_ld", i32), ("nch", i32), ("dst_c0", i32), ("reserved", i32)]


MAX_PARTS = 8


class Conv(C.Structure):
    _fields_ = [("N", i32), ("H", i32), ("W", i32), ("Hi", i32), ("Wi", i32),
                ("KH", i32), ("KW", i32), ("stride", i32), ("pad_h", i32), ("pad_w", i32),
                ("nsrc", i32), ("nseg", i32), ("cout_pad", i32), ("kind", i32),
                ("src", Src * MAX_SRC), ("seg", Seg * MAX_SEG), ("wpack", vp), ("bias", vp)]


# every symbol include/demfi_b200.h declares: name -> (restype, argtypes)
_F = C.POINTER(C.c_float)
SYMBOLS = {
    "demfi_version": (i32, []),
    "demfi_last_error": (C.c_char_p, []),
    "demfi_device_check": (i32, [i32]),
    "demfi_packed_weight_floats": (C.c_size_t, [i32, i32, i32, C.POINTER(i32), i32, i32]),
    "demfi_pack_weights_device": (i32, [i32, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "demfi_pack_weights": (i32, [i32, vp, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), i32,
                                 C.POINTER(i32), i32, vp]),
    "demfi_conv2d": (i32, [C.POINTER(Conv), vp]),
    "demfi_conv_describe": (i32, [C.POINTER(Conv), C.POINTER(i32)]),
    "demfi_pack_input": (i32, [vp, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp]),
    "demfi_cfr_splat": (i32, [vp, i32, vp, i32, i32, i32, vp, vp]),
    "demfi_cfr_finalize": (i32, [vp, vp, i32, i32, i32, vp, i32, vp]),
    "demfi_pwb": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, vp, i32, vp]),
    "demfi_bwarp_blend": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
    "demfi_fgac_sample": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "demfi_fgac_blend": (i32, [vp, i32, vp, i32, vp, i32, C.c_int64, i32, vp, i32, vp]),
    "demfi_copy_channels": (i32, [vp, i32, vp, i32, i32, C.c_int64, i32, vp]),
    "demfi_gather_channels": (i32, [C.POINTER(Part), i32, vp, i32, C.c_int64, vp]),
    "demfi_channel_absmean": (i32, [vp, i32, vp, i32, C.c_int64, i32, vp, vp]),
    "demfi_upsample2x": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "demfi_export_nchw": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "demfi_import_nchw": (i32, [vp, i32, i32, i32, i32, vp, i32, vp]),
    "demfi_conv2d_wgrad": (i32, [vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "demfi_conv2d_dgrad_strided": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "demfi_act_backward": (i32, [vp, i32, vp, i32, C.c_int64, i32, i32, vp, i32, vp]),
    "demfi_bwarp_blend_backward": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, i32,
                                         vp, i32, vp, i32, vp]),
    "demfi_fgac_sample_backward": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
    "demfi_cfr_backward": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]),
    "demfi_fgac_blend_backward": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, C.c_int64, i32, vp, i32, vp, i32, vp, i32, vp]),
    "demfi_upsample2x_backward": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "demfi_l1_sum_workspace": (C.c_int64, [C.c_int64]),
    "demfi_l1_sum": (i32, [vp, vp, C.c_int64, C.c_float, vp, vp, C.c_int64, vp, vp]),
    "demfi_adam_step": (i32, [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32, vp]),
    "demfi_frame_metrics_workspace": (C.c_int64, [i32, i32, i32, i32]),
    "demfi_frame_metrics": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, C.c_int64, vp, vp]),
    "demfi_launch_count": (C.c_uint64, []),
    "demfi_set_option": (i32, [C.c_char_p, i32]),
    "demfi_get_option": (i32, [C.c_char_p, C.POINTER(i32)]),
    "demfi_tc_debug_read": (i32, [C.POINTER(C.c_int64), i32]),
}

_lib = None


class DemfiError(RuntimeError):
    pass


def lib():
    """Load libdemfi_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DemfiError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                             "demfi_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise DemfiError(f"{what} failed (rc={rc}): {lib().demfi_last_error().decode()}")


def set_option(name: str, value: int):
    check(lib().demfi_set_option(name.encode(), int(value)), "demfi_set_option")


def get_option(name: str) -> int:
    v = i32(0)
    check(lib().demfi_get_option(name.encode(), C.byref(v)), "demfi_get_option")
    return v.value


def launch_count() -> int:
    return int(lib().demfi_launch_count())
