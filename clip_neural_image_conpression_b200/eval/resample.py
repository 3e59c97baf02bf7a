"""BICUBIC resize of the eval originals on the device, bit-identical to Pillow.

The reference's eval loop loads every original with `Image.open(p).convert("RGB").resize((S, S), Image.BICUBIC)` and maps
it to float CHW in [-1, 1] (PKG/cli/eval.py:66-67).  Decoding the file stays on the host (PIL); the resampling and the
float conversion run in `resample_u8_kernel` / `u8_hwc_to_float_chw_kernel`.  Pillow's 8-bit resampler
(src/libImaging/Resample.c) is integer arithmetic on coefficients that it precomputes in double precision:
`bicubic_coeffs` below repeats that precomputation operation by operation (same order of the double operations, the
sequential sum included), so the int32 tables — and therefore every output byte — are identical.
"""
from __future__ import annotations

import functools
import math
from typing import Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import check, on_tensor_device, ptr, require_cuda, stream_ptr

PRECISION_BITS = 32 - 8 - 2      # Resample.c: coefficients as int32 with 22 fractional bits


def _bicubic_filter(x: np.ndarray) -> np.ndarray:
    """Resample.c bicubic_filter (a = -0.5), evaluated in float64 with the same expression order."""
    a = -0.5
    x = np.abs(x)
    out = np.zeros_like(x)
    m1 = x < 1.0
    m2 = (x >= 1.0) & (x < 2.0)
    out[m1] = ((a + 2.0) * x[m1] - (a + 3.0)) * x[m1] * x[m1] + 1
    out[m2] = (((x[m2] - 5) * x[m2] + 8) * x[m2] - 4) * a
    return out


@functools.lru_cache(maxsize=256)
def bicubic_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box: (bounds int32 [out, 2] = (xmin, count),
    kk int32 [out, ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale                       # bicubic support = 2
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic_filter((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = float(np.cumsum(w)[-1]) if xmax > 0 else 0.0          # sequential double sum, like the C loop
        if ww != 0.0:
            w = w / ww
        q = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :xmax] = np.trunc(q).astype(np.int64).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


@functools.lru_cache(maxsize=64)
def _device_tables(in_size: int, out_size: int, device_index: int):
    bounds, kk = bicubic_coeffs(in_size, out_size)
    dev = torch.device("cuda", device_index)
    return torch.from_numpy(bounds).to(dev), torch.from_numpy(kk).to(dev), kk.shape[1]


@on_tensor_device
def resize_bicubic_u8(img_hwc: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 [H, W, C] (CUDA) -> uint8 [out_h, out_w, C], byte-identical to PIL's Image.resize((out_w, out_h), BICUBIC)."""
    require_cuda(img_hwc)
    assert img_hwc.dtype == torch.uint8 and img_hwc.dim() == 3
    lib = _lib.load()
    x = img_hwc.contiguous()
    h, w, c = x.shape
    di = x.device.index if x.device.index is not None else torch.cuda.current_device()
    if w != out_w:                                    # horizontal pass first (ImagingResampleInner)
        b, k, ks = _device_tables(w, out_w, di)
        y = torch.empty((h, out_w, c), dtype=torch.uint8, device=x.device)
        check(lib.clpk_resample_u8(ptr(x), ptr(y), ptr(b), ptr(k), ks, h, w, out_w, c, stream_ptr()), "clpk_resample_u8")
        x, w = y, out_w
    if h != out_h:
        b, k, ks = _device_tables(h, out_h, di)
        y = torch.empty((out_h, w, c), dtype=torch.uint8, device=x.device)
        check(lib.clpk_resample_u8(ptr(x), ptr(y), ptr(b), ptr(k), ks, 1, h, out_h, w * c, stream_ptr()), "clpk_resample_u8")
        x = y
    return x


@on_tensor_device
def u8_hwc_to_float_chw(img_hwc: torch.Tensor) -> torch.Tensor:
    """uint8 [H, W, C] -> fp32 [C, H, W] = u8 / 127.5 - 1 (eval.py:67)."""
    require_cuda(img_hwc)
    x = img_hwc.contiguous()
    h, w, c = x.shape
    out = torch.empty((c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().clpk_u8_hwc_to_float_chw(ptr(x), ptr(out), h, w, c, stream_ptr()), "clpk_u8_hwc_to_float_chw")
    return out


def load_original_device(path: str, size: int, device) -> torch.Tensor:
    """eval.py:66-67 with the resize and the float conversion on the device: decode (PIL, host) -> H2D -> BICUBIC -> CHW."""
    from PIL import Image

    rgb = np.array(Image.open(path).convert("RGB"))      # a writable copy: torch.from_numpy refuses read-only views silently
    t = torch.from_numpy(rgb).to(device)
    return u8_hwc_to_float_chw(resize_bicubic_u8(t, size, size))
