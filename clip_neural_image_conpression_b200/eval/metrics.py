"""Reconstruction metrics (interface mirror of the reference's PKG/eval/metrics.py:16-46).

psnr: both images -> uint8 domain ((x+1)*127.5, clip, truncate) -> squared error summed EXACTLY in int64 on the device
(clpk_psnr_sqerr_u8) -> 20*log10(255/sqrt(mse)).  The reference averages the squares in float32 with numpy's pairwise
sum, so its MSE carries ~1e-7 relative rounding noise; agreement is therefore to ~1e-5 dB, not bitwise.
ssim: the reference defers to scikit-image's structural_similarity(HWC uint8, data_range=255, channel_axis=-1); here the
same quantity (7x7 uniform window, K1 0.01, K2 0.03, sample covariance, float64, 3-pixel border cropped, mean over
pixels then channels) is computed on the device with exact integer window sums (clpk_ssim_u8).  scikit-image is not
vendored, so the parity of this metric is pinned only to the restated algorithm (parity unpinned).
lpips / clip_similarity need pretrained networks and are out of scope.
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


def ssim_batch(a: torch.Tensor, b: torch.Tensor) -> list[float]:
    """Per-image SSIM of two [B, C, H, W] fp32 CUDA tensors in [-1, 1] (C > 1: mean over channels)."""
    return ops.ssim_u8(a, b).cpu().tolist()


def ssim(img1, img2) -> float:
    """metrics.py:32-46.  (C,H,W) with C in (1, 3), or (H,W,C), or (H,W) arrays in [-1, 1]."""
    require_cuda()
    x1 = np.ascontiguousarray(img1, dtype=np.float32)
    x2 = np.ascontiguousarray(img2, dtype=np.float32)
    if x1.ndim == 2:                                    # plain (H, W) image: one plane
        x1, x2 = x1[None], x2[None]
    else:
        if x1.shape[0] not in (1, 3):                   # (H, W, C): the reference leaves it as is -> planes = last axis
            x1, x2 = x1.transpose(2, 0, 1), x2.transpose(2, 0, 1)
        if x1.shape[0] == 1:
            # Single channel in a 3-D array: the reference hands skimage an (H, W, 1) array with channel_axis=None,
            # whose 7-wide window cannot fit the length-1 axis -> skimage raises.  Mirror the error, do not invent a value.
            raise ValueError("win_size exceeds image extent. Either ensure that your images are at least 7x7; or pass "
                             "win_size explicitly in the function call, with an odd value less than or equal to the "
                             "smaller side of your images.")
    a = torch.from_numpy(np.ascontiguousarray(x1)).cuda()[None]
    b = torch.from_numpy(np.ascontiguousarray(x2)).cuda()[None]
    return ssim_batch(a, b)[0]


def lpips_distance(img1, img2, device: str = "cpu") -> float:
    return float("nan")


def clip_similarity(img1, img2, device: str = "cpu") -> float:
    return float("nan")
