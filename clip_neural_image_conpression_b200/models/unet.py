"""CLIPCondUNet: the FiLM-conditioned epsilon predictor, run as a libclpk plan on one B200.

Interface mirror of the reference's PKG/models/unet.py:22-106: `timestep_embedding(t, dim, max_period)`,
`CLIPCondUNet(z_dim, base, ch_mult, time_dim, img_ch)` with the reference's parameter tree and
`forward(x_t, z_clip, t) -> eps`.  The nn.Modules own the fp32 parameters; on the first CUDA call for a given
(batch, H, W) a *plan* is built in C++ (16-bit — fp16 default, bf16 selectable — K-major weight repack, NHWC workspaces, TMA descriptors, launch
sequence) and cached.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch
from torch import nn

from .. import _lib, ops
from .._lib import UnetConfig, check, op_code, ptr, require_cuda, stream_ptr
from .blocks import FiLM, ResBlock  # noqa: F401  (FiLM re-exported like the reference module does)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """Sinusoidal embedding [cos(t f_k) | sin(t f_k)], f_k = exp(-ln(max_period) k / (dim//2)) (reference :22-39)."""
    require_cuda(t)
    return ops.timestep_embedding(t, dim, float(max_period))


class _Plan:
    """Owner of one clpk_plan* (fixed batch / image size)."""

    def __init__(self, net: "CLIPCondUNet", batch: int, height: int, width: int) -> None:
        lib = _lib.load()
        cfg = UnetConfig()
        cfg.z_dim, cfg.base, cfg.n_levels = net.z_dim, net.base, len(net.ch_mult)
        for i, m in enumerate(net.ch_mult):
            cfg.ch_mult[i] = int(m)
        cfg.time_dim, cfg.img_ch, cfg.groups = net.time_dim, net.img_ch, 8
        cfg.op_dtype = op_code(net.operand_dtype)
        sd = {k: v.detach().float().contiguous() for k, v in net.state_dict().items()}
        devs = {v.device for v in sd.values()}
        if len(devs) != 1 or next(iter(devs)).type != "cuda":
            raise _lib.ClpkError("CLIPCondUNet parameters must live on ONE CUDA device")
        self.device = next(iter(devs))
        names = (C.c_char_p * len(sd))(*[k.encode() for k in sd])
        ptrs = (C.c_void_p * len(sd))(*[v.data_ptr() for v in sd.values()])
        numels = (C.c_int64 * len(sd))(*[v.numel() for v in sd.values()])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):   # the plan's memory, tensor maps and graph belong to the weights' GPU
            torch.cuda.current_stream().synchronize()
            check(lib.clpk_plan_create(C.byref(cfg), batch, height, width, len(sd), names, ptrs, numels, C.byref(handle)),
                  "clpk_plan_create")
        self.lib, self.handle = lib, handle
        with torch.cuda.device(self.device):   # timestep-embedding frequencies as the reference's torch-CPU exp gives them
            f = ops.timestep_frequencies(net.time_dim).to(self.device)
            torch.cuda.current_stream().synchronize()
            check(lib.clpk_plan_set_time_freqs(handle, f.data_ptr()), "clpk_plan_set_time_freqs")
        self.batch, self.height, self.width = batch, height, width
        self.ddim_key = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.clpk_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 — interpreter shutdown
            pass

    @property
    def flops_per_forward(self) -> float:
        return float(self.lib.clpk_plan_flops_per_forward(self.handle))

    @property
    def launches_per_forward(self) -> int:
        return int(self.lib.clpk_plan_launches_per_forward(self.handle))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.clpk_plan_device_bytes(self.handle))


class CLIPCondUNet(nn.Module):
    """FiLM-conditioned pixel-space UNet (reference unet.py:42-106), executed by hand-written sm_100a kernels."""

    def __init__(self, z_dim: int = 512, base: int = 128, ch_mult: Tuple[int, ...] = (1, 2, 2), time_dim: int = 256,
                 img_ch: int = 3) -> None:
        super().__init__()
        self.z_dim, self.base, self.ch_mult = z_dim, base, tuple(int(m) for m in ch_mult)
        self.time_dim, self.img_ch = time_dim, img_ch
        # parameter containers, created in the reference's order (unet.py:47-79)
        self.time_proj = nn.Sequential(nn.Linear(time_dim, 4 * time_dim), nn.SiLU(), nn.Linear(4 * time_dim, time_dim))
        self.z_proj = nn.Sequential(nn.Linear(z_dim, time_dim), nn.SiLU())
        self.in_conv = nn.Conv2d(img_ch, base, 3, padding=1)
        width = base
        self.down_chs: List[int] = [width]
        stages = []
        for m in self.ch_mult:
            stages += [ResBlock(width, time_dim), ResBlock(width, time_dim),
                       nn.Conv2d(width, width * m, 3, stride=2, padding=1)]
            width *= m
            self.down_chs.append(width)
        self.down = nn.ModuleList(stages)
        self.mid1 = ResBlock(width, time_dim)
        self.mid2 = ResBlock(width, time_dim)
        stages = []
        for m in reversed(self.ch_mult):
            stages += [ResBlock(width, time_dim), ResBlock(width, time_dim),
                       nn.ConvTranspose2d(width, width // m, 4, stride=2, padding=1)]
            width //= m
        self.up = nn.ModuleList(stages)
        self.out_norm = nn.GroupNorm(8, width)
        self.out = nn.Conv2d(width, img_ch, 3, padding=1)
        # Tensor-core operand format.  fp16 (default) and bf16 run at the same tcgen05 rate with fp32 accumulation;
        # fp16's 3 extra mantissa bits keep the per-step epsilon error ~8x lower (DESIGN.md "Precision").
        self.operand_dtype = torch.float16
        self._plans: dict = {}
        self._plan_version = None

    # ------------------------------------------------------------------ plan cache
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def plan_for(self, batch: int, height: int, width: int) -> _Plan:
        """Returns the (cached) plan for this shape; rebuilt when parameters were replaced or modified in place."""
        ver = (self._weights_version(), self.operand_dtype)
        if ver != self._plan_version:
            self._plans.clear()
            self._plan_version = ver
        key = (batch, height, width)
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan(self, batch, height, width)
            self._plans[key] = plan
        return plan

    def release_plans(self) -> None:
        self._plans.clear()

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x_t: torch.Tensor, z_clip: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        require_cuda(x_t, z_clip, t)
        if next(self.parameters()).device != x_t.device:
            raise _lib.ClpkError("CLIPCondUNet parameters and inputs must live on the same CUDA device")
        b, c, h, w = x_t.shape
        if c != self.img_ch or z_clip.shape != (b, self.z_dim) or t.shape != (b,):
            raise ValueError(f"bad input shapes x{tuple(x_t.shape)} z{tuple(z_clip.shape)} t{tuple(t.shape)}")
        plan = self.plan_for(b, h, w)
        x = x_t.contiguous().float()
        z = z_clip.contiguous().float()
        tt = t.contiguous().to(torch.int64)
        eps = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(plan.lib.clpk_unet_forward(plan.handle, ptr(x), ptr(z), ptr(tt), ptr(eps), stream_ptr()),
                  "clpk_unet_forward")
        return eps
