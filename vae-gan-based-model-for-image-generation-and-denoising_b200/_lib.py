"""ctypes binding of libvaegan_b200.so (the C-ABI declared in include/vaegan_b200.h).

The library is the product: if it is missing or the device is not sm_100 every call raises - there is no
PyTorch / CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_longlong, c_size_t, c_ulonglong,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvaegan_b200.so")

VG_OK = 0
VG_F32, VG_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4


class VgConvGeom(Structure):
    _fields_ = [("batch", c_int32), ("big_h", c_int32), ("big_w", c_int32), ("big_c", c_int32),
                ("small_h", c_int32), ("small_w", c_int32), ("small_c", c_int32),
                ("kernel", c_int32), ("stride", c_int32), ("pad", c_int32), ("big_c_valid", c_int32)]


class VgPackItem(Structure):
    _fields_ = [("w", c_void_p), ("wd", c_void_p), ("wu", c_void_p), ("small_c", c_int32), ("big_c", c_int32),
                ("big_c_valid", c_int32), ("kk", c_int32)]


class VgEpilogue(Structure):
    _fields_ = [("mode", c_int32), ("groups", c_int32), ("channels", c_int32), ("act", c_int32), ("slope", c_float),
                ("sums", c_void_p), ("x", c_void_p), ("stats", c_void_p)]


DP_MAX_RANKS = 8


class VgDpComm(Structure):
    """include/vaegan_b200.h: symmetric-memory mappings of one network's flat gradient / parameter buffers."""
    _fields_ = [("mc_grads", c_void_p), ("mc_params", c_void_p), ("peer_grads", c_void_p * DP_MAX_RANKS),
                ("peer_params", c_void_p * DP_MAX_RANKS), ("peer_sig", c_void_p * DP_MAX_RANKS), ("epoch", c_void_p),
                ("rank", c_int32), ("world", c_int32)]


WGRAD_OVERWRITE, WGRAD_DST_ZERO = 1, 2
EPI_NONE, EPI_BN_STATS, EPI_BN_BWD, EPI_ACT_BWD, EPI_ACT_FWD, EPI_AFFINE_ACT_FWD = 0, 1, 2, 3, 4, 5

# name -> (restype, argtypes); mirrors include/vaegan_b200.h one to one
_G = POINTER(VgConvGeom)
_E = POINTER(VgEpilogue)
_P = c_void_p
PROTOTYPES = {
    "vg_last_error": (c_char_p, []),
    "vg_version": (c_int, []),
    "vg_device_check": (c_int, []),
    "vg_launch_count": (c_longlong, []),
    "vg_pack_weights_bf16": (c_int, [_G, _P, _P, _P, _P]),
    "vg_pack_weights_multi": (c_int, [POINTER(VgPackItem), c_int, _P]),
    "vg_conv_down_workspace_bytes": (c_size_t, [_G]),
    "vg_conv_down": (c_int, [_G, c_int, _P, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "vg_conv_up": (c_int, [_G, c_int, _P, _P, _P, _P]),
    "vg_conv_epilogue_supported": (c_int, [_G, c_int, c_int, _E]),
    "vg_conv_down_ex": (c_int, [_G, c_int, _P, _P, _P, _P, _E, _P]),
    "vg_conv_up_ex": (c_int, [_G, c_int, _P, _P, _P, _E, _P]),
    "vg_conv_wgrad_workspace_bytes": (c_size_t, [_G, c_int]),
    "vg_conv_wgrad": (c_int, [_G, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "vg_conv_wgrad_ex": (c_int, [_G, c_int, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "vg_reduce_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "vg_bn_bwd_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "vg_bn_train_fwd": (c_int, [_P, c_int, c_longlong, c_int, _P, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P,
                                c_size_t, _P]),
    "vg_bn_act_train_fwd": (c_int, [_P, c_int, c_longlong, c_int, _P, _P, _P, _P, _P, c_float, c_float, c_int, c_float, _P,
                                    _P, _P]),
    "vg_bn_act_train_bwd": (c_int, [_P, _P, c_int, c_longlong, c_int, _P, c_int, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "vg_bn_apply_from_sums": (c_int, [_P, c_int, c_longlong, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float, c_float, c_int,
                                      c_float, _P, _P, _P]),
    "vg_bn_bwd_apply_from_sums": (c_int, [_P, _P, c_int, c_longlong, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "vg_bn_eval_coeffs": (c_int, [_P, _P, _P, _P, c_float, c_int, _P, _P, _P]),
    "vg_scale_shift_act": (c_int, [_P, c_int, c_longlong, c_int, _P, _P, c_int, c_float, _P, c_int, _P]),
    "vg_bn_act_bwd": (c_int, [_P, _P, c_int, c_longlong, c_int, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P,
                              c_size_t, _P]),
    "vg_act_bwd": (c_int, [_P, _P, c_int, c_longlong, c_int, c_float, _P, c_int, _P]),
    "vg_colsum": (c_int, [_P, c_int, c_longlong, c_int, _P, _P, c_size_t, _P]),
    "vg_nchw_to_nhwc": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "vg_nchw_to_s2d": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "vg_s2d_to_nchw": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "vg_u8_nhwc_to_nchw": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "vg_linear_permute": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "vg_gather_f32": (c_int, [_P, _P, _P, c_longlong, c_int, c_int, _P]),
    "vg_nhwc_to_nchw": (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "vg_reparam_fwd": (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, _P, _P]),
    "vg_reparam_bwd": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, _P, c_float, _P, _P, _P]),
    "vg_bce": (c_int, [_P, c_int, c_float, c_float, _P, c_int, _P, _P]),
    "vg_mse_workspace_bytes": (c_size_t, []),
    "vg_mse": (c_int, [_P, _P, c_longlong, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "vg_bce_pair": (c_int, [_P, c_int, c_float, c_float, c_float, _P, _P, _P]),
    "vg_mse_total": (c_int, [_P, _P, c_int, c_longlong, c_float, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_size_t, _P]),
    "vg_total_loss": (c_int, [_P, _P, _P, _P, c_float, c_float, _P, _P]),
    "vg_adam_step": (c_int, [_P, _P, _P, _P, c_longlong, c_double, c_double, c_double, c_double, _P, c_float, _P]),
    "vg_adam_tick": (c_int, [_P, _P]),
    "vg_adam_apply": (c_int, [_P, _P, _P, _P, c_longlong, c_double, c_double, c_double, c_double, _P, c_float, _P]),
    "vg_dp_max_blocks": (c_int, []),
    "vg_dp_adam_bucket": (c_int, [_P, c_longlong, c_longlong, _P, _P, c_double, c_double, c_double, c_double, _P, c_float,
                                  c_int, c_int, _P, _P]),
    "vg_randn": (c_int, [_P, c_longlong, c_ulonglong, _P, c_ulonglong, _P]),
}

_lib = None


class VaeganB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built - run `make` / __graft_entry__.build()."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise VaeganB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `make` at the repo root). "
            "This package has no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().vg_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != VG_OK:
        raise VaeganB200Error(f"{what} failed ({rc}): {last_error()}")


class LaunchTimer:
    """Optional per-call timing for bench.py's roofline: while `active`, every C-ABI call is bracketed by a CUDA-event
    pair on the current (= launching) stream and tagged with the algorithmic FLOPs the caller states.  Off by
    default - nothing is recorded on the training path."""
    active = False
    records: list = []

    @classmethod
    def start(cls) -> None:
        cls.active, cls.records = True, []

    @classmethod
    def stop(cls):
        """-> [(entry point, tag, flops, algorithmic bytes, milliseconds)] after synchronising the device."""
        import torch
        cls.active = False
        torch.cuda.synchronize()
        out = [(name, tag, flops, nb, a.elapsed_time(b)) for name, tag, flops, nb, a, b in cls.records]
        cls.records = []
        return out


def call(name: str, *args, flops: float = 0.0, tag: str = "", nbytes: float = 0.0) -> None:
    """Invoke an int-returning entry point and raise RuntimeError (like the reference's torch ops) on failure."""
    if LaunchTimer.active:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(load(), name)(*args)
        b.record()
        LaunchTimer.records.append((name, tag, flops, nbytes, a, b))
    else:
        rc = getattr(load(), name)(*args)
    if rc != VG_OK:
        raise VaeganB200Error(f"{name} failed ({rc}): {last_error()}")
