"""DDIM sampler (interface mirror of the reference's PKG/diffusion/ddim.py:14-46).

`DDIMSampler(scheduler, eta).sample(model, z_clip, shape, steps, cfg_scale, x_T)` keeps the reference's exact — and
non-textbook — update rule: "previous" alpha-bar is alphas_cumprod_prev[t] (t-1, not the
next sub-sampled step), the direction coefficient is sqrt(a_s - sigma^2), the last step uses a_s = 1 and the result is
returned unclamped.  cfg_scale is accepted and ignored, like the reference.

Fast path (model is this package's CLIPCondUNet on CUDA): the whole loop runs inside libclpk — per-step conditioning
tables, one captured CUDA graph per step, fused update kernel, zero host synchronisation.
Generic path (any other callable): the literal loop, with the fused update kernel doing ddim.py:36-45.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import ops
from .._lib import check, ptr, require_cuda, stream_ptr
from ..models.unet import CLIPCondUNet


def ddim_timesteps(T: int, steps: int) -> torch.Tensor:
    """ts = linspace(T-1, 0, steps).long() evaluated on the host in fp32 (ddim.py:25)."""
    return torch.linspace(T - 1, 0, steps).long()


def ddim_coefficients(sched, ts: torch.Tensor, eta: float) -> torch.Tensor:
    """[steps, 5] fp32 host table {sqrt(1-a_t), sqrt(a_t), sqrt(a_s), sqrt(a_s - sigma^2), sigma} per step, built
    from the scheduler's host tables with the same fp32 torch ops, in the same order, as ddim.py:34-43."""
    abar = sched._host["alphas_cumprod"] if hasattr(sched, "_host") else sched.alphas_cumprod.detach().cpu()
    abar_prev = sched._host["alphas_cumprod_prev"] if hasattr(sched, "_host") else sched.alphas_cumprod_prev.detach().cpu()
    steps = ts.numel()
    rows = []
    for i in range(steps):
        t = ts[i]
        a_t = abar[t]
        a_s = abar_prev[t] if i < steps - 1 else torch.tensor(1.0)
        if a_s != 0:
            sigma = eta * torch.sqrt((1 - a_s) / (1 - a_t) * (1 - a_t / a_s))
        else:
            sigma = torch.tensor(0.0)
        sigma = torch.as_tensor(sigma, dtype=torch.float32)
        rows.append(torch.stack([torch.sqrt(1 - a_t), torch.sqrt(a_t), torch.sqrt(a_s),
                                 torch.sqrt(a_s - sigma ** 2), sigma]))
    return torch.stack(rows).to(torch.float32).contiguous()


class DDIMSampler:
    def __init__(self, scheduler, eta: float = 0.0) -> None:
        self.sch = scheduler
        self.eta = eta
        self.use_graph = True      # CUDA-graph replay of the step (set False to launch kernels one by one)
        # Philox key of the in-kernel eta > 0 noise.  None (default): every sample() call draws a fresh 63-bit key from
        # torch's default generator, so the noise follows torch.manual_seed / --seed and differs between calls,
        # micro-batches and ranks like the reference's torch.randn_like draws (ddim.py:44-45).  An int pins the key
        # (reproducible tests).
        self.seed = None

    @torch.no_grad()
    def sample(self, model, z_clip: torch.Tensor, shape: tuple, steps: int = 50, cfg_scale: float = 1.0,
               x_T: Optional[torch.Tensor] = None, *, noise: Optional[torch.Tensor] = None,
               trace: Optional[dict] = None) -> torch.Tensor:
        """Extra keyword-only arguments (not in the reference): `noise` = pre-drawn [steps, *shape] N(0,1) used instead
        of the in-kernel generator when eta > 0; `trace` = dict that receives 'x' and 'eps' [steps, *shape] tensors."""
        require_cuda(z_clip, x_T)
        device = z_clip.device
        ts = ddim_timesteps(self.sch.timesteps, steps)
        coef = ddim_coefficients(self.sch, ts, float(self.eta))
        x = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        if isinstance(model, CLIPCondUNet):
            return self._sample_plan(model, z_clip, x, ts, coef, noise, trace)
        return self._sample_generic(model, z_clip, x, ts, coef, noise, trace)

    # ------------------------------------------------------------------ whole loop inside libclpk
    def _sample_plan(self, net: CLIPCondUNet, z_clip, x, ts, coef, noise, trace):
        b, _, h, w = x.shape
        steps = ts.numel()
        plan = net.plan_for(b, h, w)
        key = (steps, float(self.eta), bool(self.use_graph), tuple(ts.tolist()), coef.numpy().tobytes())
        if plan.ddim_key != key:
            ts_arr = (C.c_int64 * steps)(*ts.tolist())
            cf = coef.reshape(-1).tolist()
            cf_arr = (C.c_float * len(cf))(*cf)
            with torch.cuda.device(x.device):
                check(plan.lib.clpk_plan_prepare_ddim(plan.handle, steps, ts_arr, cf_arr, int(self.use_graph), stream_ptr()),
                      "clpk_plan_prepare_ddim")
            plan.ddim_key = key
        x = x.contiguous().clone()
        z = z_clip.contiguous().float()
        nz = noise.contiguous().float() if noise is not None else None
        eps_tr = x_tr = None
        if trace is not None:
            eps_tr = torch.empty((steps,) + tuple(x.shape), dtype=torch.float32, device=x.device)
            x_tr = torch.empty_like(eps_tr)
        seed = self.seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 63 - 1, (1,), dtype=torch.int64).item()) if float(self.eta) > 0 else 0
        with torch.cuda.device(x.device):
            check(plan.lib.clpk_ddim_sample(plan.handle, ptr(z), ptr(x), ptr(nz), int(seed) & (2 ** 64 - 1), ptr(eps_tr),
                                            ptr(x_tr), stream_ptr()), "clpk_ddim_sample")
        if trace is not None:
            trace["eps"], trace["x"] = eps_tr, x_tr
        return x

    # ------------------------------------------------------------------ arbitrary epsilon model
    def _sample_generic(self, model, z_clip, x, ts, coef, noise, trace):
        steps = ts.numel()
        b = x.shape[0]
        if trace is not None:
            trace["eps"], trace["x"] = [], []
        for i in range(steps):
            t_b = torch.full((b,), int(ts[i]), device=x.device, dtype=torch.long)
            eps = model(x, z_clip, t_b)
            if trace is not None:
                trace["x"].append(x.clone())
                trace["eps"].append(eps.clone())
            c = coef[i].tolist()
            nz = None
            if c[4] > 0:
                nz = noise[i] if noise is not None else torch.randn_like(x)
            x = ops.ddim_step(x, eps, c, nz)
        if trace is not None:
            trace["eps"], trace["x"] = torch.stack(trace["eps"]), torch.stack(trace["x"])
        return x
