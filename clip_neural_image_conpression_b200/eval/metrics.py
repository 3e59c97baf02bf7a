"""Reconstruction metrics (interface mirror of the reference's PKG/eval/metrics.py:16-46).

psnr: both images -> uint8 domain ((x+1)*127.5, clip, truncate) -> squared error summed EXACTLY in int64 on the device
(clpk_psnr_sqerr_u8) -> 20*log10(255/sqrt(mse)).  The reference averages the squares in float32 with numpy's pairwise
sum, so its MSE carries ~1e-7 relative rounding noise; agreement is therefore to ~1e-5 dB, not bitwise.
ssim: the reference defers to scikit-image, which is not vendored and not installed here (parity
unpinned); like the reference without scikit-image, it returns NaN.  lpips / clip_similarity need pretrained networks
and are out of scope.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops
from .._lib import require_cuda


def _to_uint8(img: np.ndarray) -> np.ndarray:
    return ((img + 1.0) * 127.5).clip(0, 255).astype(np.uint8)


def psnr_from_sqerr(sq_err_sum: int, n: int) -> float:
    if sq_err_sum == 0:
        return float("inf")
    return 20.0 * math.log10(255.0 / math.sqrt(sq_err_sum / n))


def psnr_batch(a: torch.Tensor, b: torch.Tensor) -> list[float]:
    """Per-image PSNR of two [B, ...] fp32 CUDA tensors in [-1, 1]."""
    sq = ops.psnr_sqerr_u8(a, b).cpu().tolist()
    n = a.numel() // a.shape[0]
    return [psnr_from_sqerr(int(s), n) for s in sq]


def psnr(img1, img2) -> float:
    require_cuda()
    a = torch.as_tensor(np.ascontiguousarray(img1, dtype=np.float32)).cuda()[None]
    b = torch.as_tensor(np.ascontiguousarray(img2, dtype=np.float32)).cuda()[None]
    return psnr_batch(a, b)[0]


def ssim(img1, img2) -> float:
    return float("nan")


def lpips_distance(img1, img2, device: str = "cpu") -> float:
    return float("nan")


def clip_similarity(img1, img2, device: str = "cpu") -> float:
    return float("nan")
