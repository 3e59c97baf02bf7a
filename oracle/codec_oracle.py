"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (numpy for byte/integer work, torch-CPU fp32 for the floating-point network) of the decode hot path of
lionl1106/Clip-Neural-image-conpression (package clip_feature_codec v0.3.0).  Each function cites the reference
file:line it follows (paths relative to the reference root; PKG = src/clip_feature_codec).

Pinning status: PINNED.  The reference itself ships no golden vectors (tests/test_unet.py:7-13 and
tests/test_blocks.py:5-10 assert shapes only), so the pins are outputs of the UNMODIFIED reference executed in the
build container: tests/golden/*.npz, written by oracle/gen_golden.py (committed), and — whenever /root/reference is
present — a live comparison in tests/test_oracle_vs_reference.py.  Exceptions, stated per SURVEY.md §8(c):
  * zstd frames: arithmetic lives in the un-vendored PyPI wheel `zstandard>=0.22.0` (pyproject.toml:21), absent here.
    Pinned through the reference's own reader/writer running over a ctypes shim of the system libzstd 1.5.5.
  * SSIM: the arithmetic lives in un-vendored scikit-image (`scikit-image>=0.22.0`, pyproject.toml:23; absent here, and
    the reference returns NaN without it, PKG/eval/metrics.py:34-37) -> PARITY UNPINNED.  `ssim` below restates
    skimage 0.22's `structural_similarity` as the reference calls it (metrics.py:45), from its published algorithm and
    defaults, on top of the same scipy.ndimage.uniform_filter skimage itself uses; it is anchored by closed-form
    known answers (identical images -> 1; constant images -> (2ab+C1)/(a^2+b^2+C1)) in tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import math
import struct
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------- zstd
_Z = None


def _zstd():
    global _Z
    if _Z is None:
        z = C.CDLL(ctypes.util.find_library("zstd") or "libzstd.so.1")
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        z.ZSTD_isError.restype = C.c_uint
        z.ZSTD_isError.argtypes = [C.c_size_t]
        _Z = z
    return _Z


def zstd_compress(data: bytes, level: int = 22) -> bytes:
    z = _zstd()
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), level)
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


def zstd_decompress(frame: bytes) -> bytes:
    z = _zstd()
    size = int(z.ZSTD_getFrameContentSize(frame, len(frame)))
    buf = C.create_string_buffer(max(size, 1))
    n = z.ZSTD_decompress(buf, size, frame, len(frame))
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


# ---------------------------------------------------------------------------------------------------------- bitstream
def clp_encode(q_bytes: bytes) -> bytes:
    """PKG/io/bitstream.py:18-23 — b"CLPF" + <I len(frame) + zstd(level 22) frame."""
    frame = zstd_compress(q_bytes, 22)
    return b"CLPF" + struct.pack("<I", len(frame)) + frame


def clp_decode(blob: bytes) -> np.ndarray:
    """PKG/io/bitstream.py:26-34 — magic check (AssertionError 'Bad magic'), length, frame -> uint8[D]."""
    assert blob[:4] == b"CLPF", "Bad magic"
    ln = struct.unpack("<I", blob[4:8])[0]
    return np.frombuffer(zstd_decompress(blob[8:8 + ln]), dtype=np.uint8)


# ---------------------------------------------------------------------------------------------------------- quantiser
def quant_fit(X: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """PKG/codecs/quantizer.py:22-27 — scale = clamp_min(max-min, 1e-8)/255, zero = min (fp32)."""
    X = np.asarray(X, dtype=np.float32)
    xmin, xmax = X.min(axis=0), X.max(axis=0)
    scale = np.maximum((xmax - xmin).astype(np.float32), np.float32(1e-8)) / np.float32(255)
    return scale.astype(np.float32), xmin.astype(np.float32)


def quant_encode(x: np.ndarray, scale: np.ndarray, zero: np.ndarray) -> np.ndarray:
    """PKG/codecs/quantizer.py:29-33 — round half to even ((x-zero)/scale), clamp [0,255], uint8."""
    v = (np.asarray(x, np.float32) - zero.astype(np.float32)) / scale.astype(np.float32)
    return np.clip(np.rint(v.astype(np.float32)), 0, 255).astype(np.uint8)


def dequant(q: np.ndarray, scale: np.ndarray, zero: np.ndarray) -> np.ndarray:
    """PKG/cli/reconstruct_diffusion.py:43 / PKG/codecs/quantizer.py:38-39 — q.astype(f32) * scale + zero (two ops)."""
    return q.astype(np.float32) * scale.astype(np.float32) + zero.astype(np.float32)


def l2_normalize(x: np.ndarray, eps: float = 1e-9) -> np.ndarray:
    """PKG/cli/reconstruct_diffusion.py:21-23 — x / max(||x||_2, eps) row-wise."""
    n = np.linalg.norm(x, axis=-1, keepdims=True)
    return (x / np.maximum(n, eps)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------- scheduler
def scheduler_tables(timesteps: int = 1000, schedule: str = "cosine") -> Dict[str, torch.Tensor]:
    """PKG/diffusion/scheduler.py:21-44 (fp32, CPU)."""
    if schedule == "linear":
        betas = torch.linspace(1e-4, 0.02, timesteps)
    elif schedule == "cosine":
        s = 0.008
        t = torch.linspace(0, timesteps, timesteps + 1) / timesteps
        ac = torch.cos((t + s) / (1 + s) * math.pi / 2) ** 2
        ac = ac / ac[0]
        betas = (1 - (ac[1:] / ac[:-1])).clamp(0.0001, 0.9999)
    else:
        raise ValueError(f"Unknown schedule {schedule}")
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat([torch.tensor([1.0]), ac[:-1]], dim=0)
    return {
        "betas": betas, "alphas": alphas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac), "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
        "posterior_variance": betas * (1.0 - ac_prev) / (1.0 - ac),
    }


# ---------------------------------------------------------------------------------------------------------------- UNet
def q_sample(tabs: Dict[str, torch.Tensor], x0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """PKG/diffusion/scheduler.py:46-49."""
    return (tabs["sqrt_alphas_cumprod"][t].view(-1, 1, 1, 1) * x0 +
            tabs["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1) * noise)


def predict_x0_from_eps(tabs: Dict[str, torch.Tensor], x_t: torch.Tensor, t: torch.Tensor, eps_hat: torch.Tensor) -> torch.Tensor:
    """PKG/diffusion/scheduler.py:51-55."""
    return (x_t - tabs["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1) * eps_hat) / \
        tabs["sqrt_alphas_cumprod"][t].view(-1, 1, 1, 1)


def p_mean_variance(tabs: Dict[str, torch.Tensor], eps: torch.Tensor, x_t: torch.Tensor, t: torch.Tensor):
    """PKG/diffusion/scheduler.py:57-68 given the model output eps = model(x_t, z_clip, t)."""
    x0_pred = predict_x0_from_eps(tabs, x_t, t, eps).clamp(-1, 1)
    al_t, al_bar_t, al_bar_prev = tabs["alphas"][t], tabs["alphas_cumprod"][t], tabs["alphas_cumprod_prev"][t]
    coef1 = (torch.sqrt(al_bar_prev) * (1 - al_t)) / (1 - al_bar_t)
    coef2 = (torch.sqrt(al_t) * (1 - al_bar_prev)) / (1 - al_bar_t)
    mean = coef1.view(-1, 1, 1, 1) * x0_pred + coef2.view(-1, 1, 1, 1) * x_t
    return mean, tabs["posterior_variance"][t].view(-1, 1, 1, 1), x0_pred


def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """PKG/models/unet.py:22-39."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, device=t.device) / half)
    args = t.float().unsqueeze(1) * freqs.unsqueeze(0)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def _resblock(sd, p: str, x: torch.Tensor, h: torch.Tensor, groups: int = 8, taps: Optional[dict] = None) -> torch.Tensor:
    """PKG/models/blocks.py:40-44 (+ FiLM :22-25).  `taps` (tests only) receives the conv1 + FiLM output as "<p>.y"."""
    c = x.shape[1]
    g = min(groups, c)
    y = F.conv2d(F.silu(F.group_norm(x, g, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)),
                 sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    s = F.linear(h, sd[p + ".film.to_scale.weight"], sd[p + ".film.to_scale.bias"])[:, :, None, None]
    b = F.linear(h, sd[p + ".film.to_shift.weight"], sd[p + ".film.to_shift.bias"])[:, :, None, None]
    y = y * (1 + s) + b
    if taps is not None:
        taps[p + ".y"] = y
    y = F.conv2d(F.silu(F.group_norm(y, g, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)),
                 sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    return x + y


def unet_forward(sd: Dict[str, torch.Tensor], ch_mult: Sequence[int], x_t: torch.Tensor, z_clip: torch.Tensor,
                 t: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """PKG/models/unet.py:81-106 as a pure function of the state dict (names: SURVEY.md Appendix B)."""
    time_dim = sd["time_proj.0.weight"].shape[1]
    temb = timestep_embedding(t, time_dim).to(x_t.dtype)
    temb = F.linear(F.silu(F.linear(temb, sd["time_proj.0.weight"], sd["time_proj.0.bias"])),
                    sd["time_proj.2.weight"], sd["time_proj.2.bias"])
    zemb = F.silu(F.linear(z_clip, sd["z_proj.0.weight"], sd["z_proj.0.bias"]))
    h = temb + zemb
    x = F.conv2d(x_t, sd["in_conv.weight"], sd["in_conv.bias"], padding=1)
    skips = []
    n = len(ch_mult)
    for i in range(n):
        x = _resblock(sd, f"down.{3 * i}", x, h, taps=taps)
        x = _resblock(sd, f"down.{3 * i + 1}", x, h, taps=taps)
        skips.append(x)
        x = F.conv2d(x, sd[f"down.{3 * i + 2}.weight"], sd[f"down.{3 * i + 2}.bias"], stride=2, padding=1)
    x = _resblock(sd, "mid1", x, h)
    x = _resblock(sd, "mid2", x, h)
    for i in range(n):
        x = _resblock(sd, f"up.{3 * i}", x, h)
        x = _resblock(sd, f"up.{3 * i + 1}", x, h)
        x = F.conv_transpose2d(x, sd[f"up.{3 * i + 2}.weight"], sd[f"up.{3 * i + 2}.bias"], stride=2, padding=1)
        if skips:
            x = x + skips.pop()
    if taps is not None:
        taps["pre_out"] = x
    x = F.group_norm(x, 8, sd["out_norm.weight"], sd["out_norm.bias"], 1e-5)
    return F.conv2d(x, sd["out.weight"], sd["out.bias"], padding=1)


def param_shapes(z_dim: int, base: int, ch_mult: Sequence[int], time_dim: int = 256, img_ch: int = 3):
    """Names and shapes of the reference state dict (PKG/models/unet.py:45-79), in creation order."""
    out = []

    def lin(p, i, o):
        out.extend([(p + ".weight", (o, i)), (p + ".bias", (o,))])

    def rb(p, c):
        out.extend([(p + ".norm1.weight", (c,)), (p + ".norm1.bias", (c,)),
                    (p + ".conv1.weight", (c, c, 3, 3)), (p + ".conv1.bias", (c,)),
                    (p + ".norm2.weight", (c,)), (p + ".norm2.bias", (c,)),
                    (p + ".conv2.weight", (c, c, 3, 3)), (p + ".conv2.bias", (c,))])
        lin(p + ".film.to_scale", time_dim, c)
        lin(p + ".film.to_shift", time_dim, c)

    lin("time_proj.0", time_dim, 4 * time_dim)
    lin("time_proj.2", 4 * time_dim, time_dim)
    lin("z_proj.0", z_dim, time_dim)
    out.extend([("in_conv.weight", (base, img_ch, 3, 3)), ("in_conv.bias", (base,))])
    ch = base
    for i, m in enumerate(ch_mult):
        rb(f"down.{3 * i}", ch)
        rb(f"down.{3 * i + 1}", ch)
        out.extend([(f"down.{3 * i + 2}.weight", (ch * m, ch, 3, 3)), (f"down.{3 * i + 2}.bias", (ch * m,))])
        ch *= m
    rb("mid1", ch)
    rb("mid2", ch)
    for i, m in enumerate(reversed(list(ch_mult))):
        rb(f"up.{3 * i}", ch)
        rb(f"up.{3 * i + 1}", ch)
        out.extend([(f"up.{3 * i + 2}.weight", (ch, ch // m, 4, 4)), (f"up.{3 * i + 2}.bias", (ch // m,))])
        ch //= m
    out.extend([("out_norm.weight", (ch,)), ("out_norm.bias", (ch,)), ("out.weight", (img_ch, ch, 3, 3)),
                ("out.bias", (img_ch,))])
    return out


def make_state_dict(z_dim: int, base: int, ch_mult: Sequence[int], seed: int = 0, out_gain: float = 1.0,
                    time_dim: int = 256, img_ch: int = 3) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (the harness owns all seeds; the reference has none).  Conv/Linear: uniform
    +-1/sqrt(fan_in) like torch's default init; GroupNorm gamma = 1 + 0.1 N(0,1), beta = 0.1 N(0,1) so the affine
    path is exercised.  out_gain < 1 damps `out.*` (protocol P-gamma of SURVEY.md §0.5: a contractive sampler)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(z_dim, base, ch_mult, time_dim, img_ch):
        if "norm" in name:
            v = torch.randn(shape, generator=g) * 0.1
            if name.endswith(".weight"):
                v = v + 1.0
        else:
            if name.endswith(".weight"):
                fan_in = int(np.prod(shape[1:])) if not (name.startswith("up.") and len(shape) == 4) else shape[1] * 16
                bound = 1.0 / math.sqrt(fan_in)
                last_bound = bound
            else:
                bound = last_bound
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = v.float()
    if out_gain != 1.0:
        sd["out.weight"] = sd["out.weight"] * out_gain
        sd["out.bias"] = sd["out.bias"] * out_gain
    return sd


# ---------------------------------------------------------------------------------------------------------------- DDIM
def ddim_timesteps(T: int, steps: int) -> torch.Tensor:
    """PKG/diffusion/ddim.py:25."""
    return torch.linspace(T - 1, 0, steps).long()


def ddim_update(x, eps, a_t, a_s, eta: float, noise=None):
    """PKG/diffusion/ddim.py:36-45 — one update; a_t, a_s are 0-dim fp32 tensors."""
    x0 = ((x - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)).clamp(-1, 1)
    sigma = eta * torch.sqrt((1 - a_s) / (1 - a_t) * (1 - a_t / a_s)) if a_s != 0 else 0.0
    x_next = torch.sqrt(a_s) * x0 + torch.sqrt(a_s - sigma ** 2) * eps
    if eta > 0 and sigma > 0:
        x_next = x_next + sigma * noise
    return x_next


def ddim_sample(eps_fn: Callable, tables: Dict[str, torch.Tensor], z_clip: torch.Tensor, x_T: torch.Tensor, steps: int,
                eta: float = 0.0, noise: Optional[torch.Tensor] = None, trace: Optional[dict] = None,
                teacher: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PKG/diffusion/ddim.py:21-46.  `noise` [steps, *shape] replaces torch.randn_like (the reference draws from the
    global generator; the harness pre-draws the same sequence).  `teacher` (optional [steps,*shape]) forces x."""
    T = tables["alphas_cumprod"].shape[0]
    ts = ddim_timesteps(T, steps)
    x = x_T
    if trace is not None:
        trace["x"], trace["eps"] = [], []
    for i in range(steps):
        t = ts[i]
        t_b = torch.full((x.shape[0],), int(t.item()), dtype=torch.long, device=x.device)
        if teacher is not None:
            x = teacher[i]
        eps = eps_fn(x, z_clip, t_b)
        if trace is not None:
            trace["x"].append(x.clone())
            trace["eps"].append(eps.clone())
        a_t = tables["alphas_cumprod"][t].to(x.device)
        a_s = (tables["alphas_cumprod_prev"][t] if i < steps - 1 else torch.tensor(1.0)).to(x.device)
        x = ddim_update(x, eps, a_t, a_s, eta, None if noise is None else noise[i])
    if trace is not None:
        trace["x"], trace["eps"] = torch.stack(trace["x"]), torch.stack(trace["eps"])
    return x


# ------------------------------------------------------------------------------------------------------------ metrics
def to_uint8_image(x_chw: np.ndarray) -> np.ndarray:
    """PKG/cli/reconstruct_diffusion.py:55-56 — clamp(-1,1), CHW->HWC, ((img+1)*127.5).astype(uint8) (truncation)."""
    img = np.clip(x_chw, -1, 1).transpose(1, 2, 0)
    return ((img + 1.0) * 127.5).astype(np.uint8)


def metric_uint8(img: np.ndarray) -> np.ndarray:
    """PKG/eval/metrics.py:16-19."""
    return ((img + 1.0) * 127.5).clip(0, 255).astype(np.uint8)


def psnr(img1: np.ndarray, img2: np.ndarray) -> float:
    """PKG/eval/metrics.py:22-29 (float32 mean, like numpy does for float32 input)."""
    x1, x2 = metric_uint8(img1), metric_uint8(img2)
    mse = np.mean((x1.astype(np.float32) - x2.astype(np.float32)) ** 2)
    if mse == 0:
        return float("inf")
    return float(20.0 * np.log10(255.0 / np.sqrt(mse)))


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    """PKG/eval/metrics.py:32-46 with scikit-image present: structural_similarity(x1, x2, data_range=255,
    channel_axis=-1) on the HWC uint8 images.  Restatement of skimage 0.22 `skimage/metrics/_structural_similarity.py`
    for these arguments (PARITY UNPINNED, see the module header): win_size 7, uniform filter, K1 = 0.01, K2 = 0.03,
    use_sample_covariance=True, float64 arithmetic (uint8 input), border of (win_size - 1) // 2 cropped, mean of S per
    channel in float64, then the mean over channels."""
    from scipy.ndimage import uniform_filter

    x1, x2 = metric_uint8(img1), metric_uint8(img2)
    if x1.ndim == 3 and x1.shape[0] in (1, 3):                      # metrics.py:41-43
        x1, x2 = x1.transpose(1, 2, 0), x2.transpose(1, 2, 0)
    multichannel = x1.ndim == 3 and x1.shape[2] > 1                 # metrics.py:44

    def plane(a: np.ndarray, b: np.ndarray) -> float:
        win, k1, k2, r = 7, 0.01, 0.03, 255.0
        if min(a.shape) < win:
            raise ValueError("win_size exceeds image extent.")
        a, b = a.astype(np.float64), b.astype(np.float64)
        npx = win ** a.ndim
        cov_norm = npx / (npx - 1)
        ux, uy = uniform_filter(a, size=win), uniform_filter(b, size=win)
        uxx, uyy, uxy = uniform_filter(a * a, size=win), uniform_filter(b * b, size=win), uniform_filter(a * b, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        c1, c2 = (k1 * r) ** 2, (k2 * r) ** 2
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        pad = (win - 1) // 2
        return float(s[tuple(slice(pad, -pad) for _ in range(a.ndim))].mean(dtype=np.float64))

    if multichannel:
        return float(np.mean([plane(x1[..., c], x2[..., c]) for c in range(x1.shape[-1])]))
    # single channel: skimage filters over every axis of whatever array it is given (HW, or HW1 after the transpose)
    return plane(x1, x2)


# ------------------------------------------------------------------------------------------------ BICUBIC originals
def _pil_bicubic_coeffs(in_size: int, out_size: int):
    """Pillow src/libImaging/Resample.c: precompute_coeffs (bicubic, a = -0.5, full-image box) + normalize_coeffs_8bpc,
    written as the plain loops of the C source.  Third-party arithmetic on the path (PKG/cli/eval.py:66 calls
    `Image.resize((S, S), Image.BICUBIC)`); pinned against Pillow itself, which IS installed here and on the GPU box
    (tests/test_oracle_golden.py::test_bicubic_restatement_matches_pillow)."""
    prec = 32 - 8 - 2

    def filt(x: float) -> float:
        a = -0.5
        x = -x if x < 0.0 else x
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0

    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds, kk = [], []
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = 0 if xmin < 0 else xmin
        xmax = int(center + support + 0.5)
        xmax = in_size if xmax > in_size else xmax
        xmax -= xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        ki = [int(-0.5 + w * (1 << prec)) if w < 0 else int(0.5 + w * (1 << prec)) for w in k]
        bounds.append((xmin, xmax))
        kk.append(ki + [0] * (ksize - xmax))
    return bounds, kk


def bicubic_resize_u8(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`Image.fromarray(img).resize((out_w, out_h), Image.BICUBIC)` for uint8 HWC images (eval.py:66): horizontal pass, then
    vertical pass (a pass is skipped when its size does not change), out = clip8((2^21 + sum k*in) >> 22)."""
    def one_pass(a: np.ndarray, out_size: int, axis: int) -> np.ndarray:
        if a.shape[axis] == out_size:
            return a
        bounds, kk = _pil_bicubic_coeffs(a.shape[axis], out_size)
        src = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((out_size,) + src.shape[1:], np.uint8)
        for o, ((xmin, n), k) in enumerate(zip(bounds, kk)):
            acc = (1 << 21) + np.tensordot(np.asarray(k[:n], np.int64), src[xmin:xmin + n], axes=(0, 0))
            out[o] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)

    return one_pass(one_pass(np.asarray(img_hwc, np.uint8), out_w, 1), out_h, 0)


def original_to_float_chw(img_hwc_u8: np.ndarray) -> np.ndarray:
    """PKG/cli/eval.py:67 — (np.array(img).astype(np.float32) / 127.5 - 1.0).transpose(2, 0, 1)."""
    return (np.asarray(img_hwc_u8).astype(np.float32) / 127.5 - 1.0).transpose(2, 0, 1)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (the north-star's per-step epsilon metric)."""
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b).clamp_min(1e-30))


def psnr_float(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    """PSNR between two [-1,1] tensors after clamp (north-star's final-reconstruction metric, data range 2)."""
    a, b = a.double().clamp(-1, 1), b.double().clamp(-1, 1)
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
