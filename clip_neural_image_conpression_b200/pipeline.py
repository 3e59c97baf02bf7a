"""Batched decode pipeline: uint8 codes (host) -> H2D -> dequantise + L2 renorm (device) -> DDIM -> images.

This is the body of the reference's eval loop (PKG/cli/eval.py:56-64) and of reconstruct_diffusion.main
(PKG/cli/reconstruct_diffusion.py:41-56), batched: the reference decodes one image at a time with B = 1.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import require_cuda
from .diffusion.ddim import DDIMSampler
from .models.unet import CLIPCondUNet


def codes_to_device(q_host: np.ndarray, device) -> torch.Tensor:
    """uint8 [N, D] host array -> device tensor through pinned memory (one async H2D copy)."""
    t = torch.from_numpy(np.array(q_host, dtype=np.uint8, order="C", copy=True))  # frombuffer views are read-only
    return t.pin_memory().to(device, non_blocking=True)


@torch.no_grad()
def decode_codes(net: CLIPCondUNet, sampler: DDIMSampler, q_host: np.ndarray, scale: torch.Tensor, zero: torch.Tensor,
                 size: int, steps: int = 50, batch: int = 8, x_T: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decodes N quantised CLIP vectors into N images [N, 3, size, size] (fp32, unclamped, on the device).

    q_host: uint8 [N, D] (already zstd-decoded, io.read_bitstreams); scale/zero: fp32 [D] device tensors from
    codec_meta.npz.  Images are processed in micro-batches of `batch`; a ragged last batch is padded by repeating the
    last code so that one plan / one CUDA graph serves the whole store.  x_T: optional [N, 3, size, size] start noise
    (the reference draws torch.randn per image, ddim.py:27)."""
    require_cuda(scale, zero)
    device = scale.device
    n = q_host.shape[0]
    img_ch = net.img_ch
    out = torch.empty((n, img_ch, size, size), dtype=torch.float32, device=device)
    if n == 0:
        return out
    q_dev = codes_to_device(q_host, device)
    z_all = ops.dequant_l2norm(q_dev, scale, zero, l2norm=True)
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        idx = torch.arange(lo, lo + batch, device=device).clamp_(max=n - 1)
        z = z_all.index_select(0, idx)
        if x_T is not None:
            x0 = x_T.to(device).index_select(0, idx)
        else:
            x0 = torch.randn((batch, img_ch, size, size), device=device)
        nz = None if noise is None else noise.to(device).index_select(1, idx)
        x = sampler.sample(net, z, (batch, img_ch, size, size), steps=steps, x_T=x0, noise=nz)
        out[lo:hi] = x[: hi - lo]
    return out
