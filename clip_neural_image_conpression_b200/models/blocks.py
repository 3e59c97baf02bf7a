"""FiLM and ResBlock of the diffusion decoder, executed by the sm_100a kernels of libclpk.so.

Interface mirror of the reference's PKG/models/blocks.py:14-44 (same constructor arguments, same parameter names and
creation order, so state dicts and seeded random inits are interchangeable).  The torch modules below only OWN the
parameters; forward() hands their storage to the CUDA kernels.  AttnBlock / DWConvBlock of the reference are not on
the decode path and are not provided.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .._lib import require_cuda


class FiLM(nn.Module):
    """y = x * (1 + to_scale(h)) + to_shift(h), broadcast over H, W (reference blocks.py:22-25)."""

    def __init__(self, c: int, cond_dim: int) -> None:
        super().__init__()
        self.to_scale = nn.Linear(cond_dim, c)
        self.to_shift = nn.Linear(cond_dim, c)

    def scale_shift(self, h: torch.Tensor):
        """(1 + s, b) as two [B, C] tensors, computed by the fp32 GEMV kernel."""
        sc = ops.linear(h, self.to_scale.weight, self.to_scale.bias + 1.0)
        sh = ops.linear(h, self.to_shift.weight, self.to_shift.bias)
        return sc, sh

    @torch.no_grad()
    def forward(self, x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
        require_cuda(x, h)
        sc, sh = self.scale_shift(h)
        return ops.film_apply(x, sc, sh)


class ResBlock(nn.Module):
    """GN -> SiLU -> conv3x3 -> FiLM -> GN -> SiLU -> conv3x3 -> + x   (reference blocks.py:31-44).

    Standalone forward (NCHW in / NCHW out) for API parity and tests; inside CLIPCondUNet the same kernels run from the
    plan on NHWC buffers without any layout change.
    """

    def __init__(self, c: int, cond_dim: int, groups: int = 8) -> None:
        super().__init__()
        self.norm1 = nn.GroupNorm(min(groups, c), c)
        self.conv1 = nn.Conv2d(c, c, 3, padding=1)
        self.norm2 = nn.GroupNorm(min(groups, c), c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)
        self.film = FiLM(c, cond_dim)
        self.act = nn.SiLU()
        self.operand_dtype = torch.float16  # tensor-core operand format (torch.bfloat16 also supported)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
        require_cuda(x, h)
        c = self.conv1.out_channels
        x_nhwc = x.float().permute(0, 2, 3, 1).contiguous()
        sc, sh = self.film.scale_shift(h)
        dt = self.operand_dtype
        a = ops.groupnorm_silu(x_nhwc, self.norm1.weight, self.norm1.bias, self.norm1.num_groups, self.norm1.eps, dtype=dt)
        y = ops.conv_igemm(a, ops.pack_conv_weight(self.conv1.weight, ops.CONV_3X3_S1, dt), ops.CONV_3X3_S1, c,
                           self.conv1.bias, film_scale1p=sc, film_shift=sh)["f32"]
        a = ops.groupnorm_silu(y, self.norm2.weight, self.norm2.bias, self.norm2.num_groups, self.norm2.eps, dtype=dt)
        out = ops.conv_igemm(a, ops.pack_conv_weight(self.conv2.weight, ops.CONV_3X3_S1, dt), ops.CONV_3X3_S1, c,
                             self.conv2.bias, resid=x_nhwc)["f32"]
        return out.permute(0, 3, 1, 2).contiguous()
