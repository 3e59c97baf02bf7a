"""DDPM noise schedule tables (interface mirror of the reference's PKG/diffusion/scheduler.py:18-68).

Only the table constructor and the attributes the DDIM sampler reads are on the decode path.  The tables are built
ONCE with torch on the host — the same op sequence as the reference (`scheduler.py:25-44`) so they are bit-identical
to its CPU tables — and then placed on `device`.  (A GPU `linspace`/`cumprod` can differ from the CPU oracle in the
last ulp)  `q_sample` / `predict_x0_from_eps` / `p_mean_variance` (training / DDPM ancestral sampling
helpers) run as ONE fused pass each on CUDA tensors (`clpk_ddpm_combine`, bit-identical to the
reference's ATen expressions); on host tensors they stay plain tensor expressions (table-level host utility only).
"""
from __future__ import annotations

import math

import torch


def _beta_table(timesteps: int, schedule: str) -> torch.Tensor:
    if schedule == "linear":
        return torch.linspace(1e-4, 0.02, timesteps)
    if schedule == "cosine":
        grid = torch.linspace(0, timesteps, timesteps + 1) / timesteps
        abar = torch.cos((grid + 0.008) / 1.008 * math.pi / 2) ** 2
        abar = abar / abar[0]
        return (1 - (abar[1:] / abar[:-1])).clamp(0.0001, 0.9999)
    raise ValueError(f"Unknown schedule {schedule}")


class NoiseScheduler:
    def __init__(self, timesteps: int = 1000, schedule: str = "cosine", device: str = "cuda") -> None:
        self.timesteps = timesteps
        self.schedule = schedule
        self.device = device
        betas = _beta_table(timesteps, schedule)  # host, fp32
        alphas = 1.0 - betas
        abar = torch.cumprod(alphas, dim=0)
        abar_prev = torch.cat([torch.ones(1), abar[:-1]])
        host = {
            "betas": betas,
            "alphas": alphas,
            "alphas_cumprod": abar,
            "alphas_cumprod_prev": abar_prev,
            "sqrt_alphas_cumprod": torch.sqrt(abar),
            "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
            "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
            "posterior_variance": betas * (1.0 - abar_prev) / (1.0 - abar),
        }
        self._host = host  # the DDIM coefficient table is derived from these on the host
        for name, tab in host.items():
            setattr(self, name, tab.to(device))

    # ---- helpers kept for interface parity (reference scheduler.py:46-68) ------------------------------------
    @staticmethod
    def _col(v: torch.Tensor) -> torch.Tensor:
        return v.view(-1, 1, 1, 1)

    def _combine(self, x, y, ca, cb, cdiv=None, clamp=False):
        if x.is_cuda:
            from .. import ops
            return ops.ddpm_combine(x, y, ca, cb, cdiv, clamp)
        out = self._col(ca) * x + self._col(cb) * y
        if cdiv is not None:
            out = out / self._col(cdiv)
        return out.clamp(-1, 1) if clamp else out

    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """scheduler.py:46-49."""
        return self._combine(x0, noise, self.sqrt_alphas_cumprod[t], self.sqrt_one_minus_alphas_cumprod[t])

    def predict_x0_from_eps(self, x_t: torch.Tensor, t: torch.Tensor, eps_hat: torch.Tensor, clamp: bool = False) -> torch.Tensor:
        """scheduler.py:51-55: (x_t - b*eps) / a, evaluated as (1*x_t + (-b)*eps) / a — the same roundings."""
        b = self.sqrt_one_minus_alphas_cumprod[t]
        return self._combine(x_t, eps_hat, torch.ones_like(b), -b, self.sqrt_alphas_cumprod[t], clamp)

    def p_mean_variance(self, model, x_t: torch.Tensor, z_clip: torch.Tensor, t: torch.Tensor):
        """scheduler.py:57-68."""
        eps = model(x_t, z_clip, t)
        x0_pred = self.predict_x0_from_eps(x_t, t, eps, clamp=True)
        a_t, abar_t, abar_prev = self.alphas[t], self.alphas_cumprod[t], self.alphas_cumprod_prev[t]
        c0 = (torch.sqrt(abar_prev) * (1 - a_t)) / (1 - abar_t)
        c1 = (torch.sqrt(a_t) * (1 - abar_prev)) / (1 - abar_t)
        mean = self._combine(x0_pred, x_t, c0, c1)
        return mean, self._col(self.posterior_variance[t]), x0_pred
