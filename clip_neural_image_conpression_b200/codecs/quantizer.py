"""Per-channel affine uint8 quantiser (interface mirror of the reference's PKG/codecs/quantizer.py:13-40).

fit / encode / decode run as CUDA kernels (clpk_quant_fit / clpk_quant_encode_u8 / clpk_dequant_l2norm_u8) and are
bit-exact with the reference's torch/numpy arithmetic: scale = clamp_min(max-min, eps)/255, zero = min,
q = clamp(round_half_even((x-zero)/scale), 0, 255), x' = float(q)*scale + zero (separately rounded mul, add).
Only num_bits = 8 and eps = 1e-8 (the reference defaults, the only values it ever uses) are supported.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import require_cuda


class PerChannelAffineQuantizer:
    def __init__(self, num_bits: int = 8, eps: float = 1e-8) -> None:
        if num_bits != 8 or eps != 1e-8:
            raise ValueError("the sm_100a quantiser kernels implement the reference defaults num_bits=8, eps=1e-8")
        self.num_bits = num_bits
        self.eps = eps
        self.scale: torch.Tensor | None = None
        self.zero: torch.Tensor | None = None

    @staticmethod
    def _dev(x) -> torch.Tensor:
        require_cuda()
        t = torch.as_tensor(x)
        return t if t.is_cuda else t.cuda()

    def fit(self, X: torch.Tensor) -> "PerChannelAffineQuantizer":
        scale, zero = ops.quant_fit(self._dev(X).float())
        # kept on the caller's device like the reference (its attributes live wherever X lives)
        dev = X.device if isinstance(X, torch.Tensor) else torch.device("cpu")
        self.scale, self.zero = scale.to(dev), zero.to(dev)
        return self

    def _check(self) -> None:
        if self.scale is None or self.zero is None:
            raise RuntimeError("Quantizer has not been fitted.")

    def encode(self, x: torch.Tensor) -> np.ndarray:
        self._check()
        q = ops.quant_encode(self._dev(x).float(), self._dev(self.scale), self._dev(self.zero))
        return q.cpu().numpy()

    def decode(self, q: np.ndarray) -> np.ndarray:
        self._check()
        qt = self._dev(np.ascontiguousarray(q, dtype=np.uint8))
        shape = qt.shape
        z = ops.dequant_l2norm(qt.reshape(-1, shape[-1]), self._dev(self.scale), self._dev(self.zero), l2norm=False)
        return z.reshape(shape).cpu().numpy()
