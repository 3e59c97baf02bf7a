"""ctypes binding of libclpk.so (C ABI declared in include/clpk.h).

There is no CPU implementation behind these calls: if the shared library cannot be built/loaded, or a compute entry
point is used without a CUDA device, an exception is raised — the product path never falls back to anything else.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

from . import build as _build

MAX_LEVELS = 8

CONV_3X3_S1 = 0
CONV_3X3_S2 = 1
CONVT_4X4_S2 = 2
CONV_1X1 = 3

OP_BF16 = 0
OP_F16 = 1


class ClpkError(RuntimeError):
    pass


class ConvEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("film_scale1p", C.c_void_p),
        ("film_shift", C.c_void_p),
        ("film_stride", C.c_int64),
        ("resid", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("out_op", C.c_void_p),
        ("out_nchw", C.c_void_p),
        ("cout_valid", C.c_int),
        ("gn_partial", C.c_void_p),
        ("gn_cpg", C.c_int),
        ("in_scale", C.c_void_p),
        ("in_shift", C.c_void_p),
        ("in_silu", C.c_int),
        ("resid_op", C.c_void_p),
    ]


class UnetConfig(C.Structure):
    _fields_ = [
        ("z_dim", C.c_int),
        ("base", C.c_int),
        ("n_levels", C.c_int),
        ("ch_mult", C.c_int * MAX_LEVELS),
        ("time_dim", C.c_int),
        ("img_ch", C.c_int),
        ("groups", C.c_int),
        ("op_dtype", C.c_int),
    ]


_vp, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes); every symbol include/clpk.h declares
SIGNATURES = {
    "clpk_last_error": (C.c_char_p, []),
    "clpk_version": (_i, []),
    "clpk_launch_count": (_u64, []),
    "clpk_dequant_l2norm_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "clpk_quant_encode_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "clpk_quant_fit": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "clpk_ddim_step": (_i, [_vp, _vp, _vp, C.POINTER(_f), _vp, _i64, _vp]),
    "clpk_timestep_embedding": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "clpk_timestep_embedding_table": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "clpk_plan_set_time_freqs": (_i, [_vp, _vp]),
    "clpk_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "clpk_film_apply": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "clpk_groupnorm_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "clpk_groupnorm_silu": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "clpk_groupnorm_finalize": (_i, [_vp, _vp, _i, _i, _i, C.c_double, _f, _vp]),
    "clpk_groupnorm_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "clpk_conv_gn_slots": (_i, [_i, _i, _i, _i, _i]),
    "clpk_conv_in_affine_supported": (_i, [_i, _i, _i, _i, _i]),
    "clpk_groupnorm_affine": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "clpk_pack_conv_weight": (_i64, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "clpk_head_conv_supported": (_i, [_i, _i, _i, _i]),
    "clpk_pack_head_weight": (_i, [_vp, _vp, _i, _i, _vp]),
    "clpk_head_conv": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "clpk_conv_igemm": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(ConvEpilogue), _vp]),
    "clpk_conv_direct": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(ConvEpilogue), _vp]),
    "clpk_stem_im2col": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "clpk_conv_in": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "clpk_plan_create": (_i, [C.POINTER(UnetConfig), _i, _i, _i, _i, C.POINTER(C.c_char_p), C.POINTER(_vp),
                              C.POINTER(_i64), C.POINTER(_vp)]),
    "clpk_plan_destroy": (None, [_vp]),
    "clpk_plan_device_bytes": (_i64, [_vp]),
    "clpk_plan_flops_per_forward": (C.c_double, [_vp]),
    "clpk_plan_launches_per_forward": (_i, [_vp]),
    "clpk_unet_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "clpk_plan_prepare_ddim": (_i, [_vp, _i, C.POINTER(_i64), C.POINTER(_f), _i, _vp]),
    "clpk_ddim_sample": (_i, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "clpk_plan_profile_steps": (_i, [_vp, _i, C.POINTER(_f), C.POINTER(_i), _vp]),
    "clpk_plan_work_breakdown": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "clpk_plan_groupnorm_bytes": (_i, [_vp, C.POINTER(C.c_double)]),
    "clpk_to_uint8_hwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "clpk_resample_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "clpk_u8_hwc_to_float_chw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "clpk_psnr_sqerr_u8": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "clpk_ddpm_combine": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp]),
    "clpk_ssim_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "clpk_ssim_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> Path:
    return _build.LIB


def load() -> C.CDLL:
    """Loads (building first if the sources are newer) csrc/libclpk.so and types every entry point."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        import os
        override = os.environ.get("CLPK_LIB")   # A/B experiments only: another build of the SAME library (tools/ab/)
        try:
            path = Path(override) if override else _build.build()
        except Exception as e:  # noqa: BLE001 — surfaced verbatim: there is nothing to fall back to
            raise ClpkError(f"libclpk.so is unavailable and could not be built: {e}") from e
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            if override and not hasattr(lib, name):
                continue             # A/B against an older build: entry points it lacks stay unbound
            fn = getattr(lib, name)  # AttributeError here means the .so does not match include/clpk.h
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().clpk_last_error().decode("utf-8", "replace")
        raise ClpkError(f"{what or 'libclpk'} failed (code {rc}): {msg}")


def require_cuda(*tensors) -> None:
    """The hot path is CUDA-only; refuse anything else loudly.  All tensors of one call must share one device."""
    import torch

    if not torch.cuda.is_available():
        raise ClpkError("this code path needs an sm_100a CUDA device (no CPU fallback exists)")
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ClpkError("expected CUDA tensors (no CPU fallback exists)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ClpkError(f"tensors of one call live on different devices ({dev} and {t.device})")


def on_tensor_device(fn):
    """Decorator: runs `fn` with the device of its first CUDA tensor argument made current, so that the stream handed to
    libclpk, its allocations and its per-device state all belong to the tensors' GPU (not to whatever device happens to
    be current in the caller)."""
    import functools

    @functools.wraps(fn)
    def wrap(*args, **kwargs):
        import torch

        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrap


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def op_code(dtype) -> int:
    """torch dtype of the 16-bit tensor-core operands -> CLPK_OP_* code."""
    import torch

    if dtype == torch.float16:
        return OP_F16
    if dtype == torch.bfloat16:
        return OP_BF16
    raise ValueError(f"tensor-core operand dtype must be torch.float16 or torch.bfloat16, got {dtype}")
