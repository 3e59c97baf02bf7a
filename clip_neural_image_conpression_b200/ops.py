"""Thin torch-tensor wrappers over the leaf entry points of libclpk.so.

torch is used for device memory and streams only; every computation happens in the CUDA kernels of csrc/.
Layout notes: reference-facing tensors are NCHW fp32; the kernels' activations are NHWC (16-bit tensor-core operands —
fp16 by default, bf16 selectable; the residual stream is fp32 or, at the wide levels of an fp16 plan, fp16 only).
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import torch

from . import _lib
from ._lib import (CONV_3X3_S1, CONV_3X3_S2, CONVT_4X4_S2, ConvEpilogue, check, on_tensor_device, op_code, ptr,
                   require_cuda, stream_ptr)

__all__ = [
    "dequant_l2norm", "quant_encode", "quant_fit", "ddim_step", "timestep_embedding", "linear", "film_apply",
    "groupnorm_silu", "groupnorm_affine", "conv_in_affine_supported", "pack_conv_weight", "conv_igemm", "conv_direct", "head_conv", "conv_in", "to_uint8_hwc", "psnr_sqerr_u8", "ssim_u8", "ddpm_combine",
    "CONV_3X3_S1", "CONV_3X3_S2", "CONVT_4X4_S2",
]


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


@on_tensor_device
def dequant_l2norm(q: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor, l2norm: bool = True,
                   return_raw: bool = False):
    """uint8 [B,D] -> fp32 [B,D]: q*scale+zero, then row L2 normalisation (reconstruct_diffusion.py:43-44)."""
    require_cuda(q, scale, zero)
    assert q.dtype == torch.uint8 and q.dim() == 2
    b, d = q.shape
    q = q.contiguous()
    z = torch.empty((b, d), dtype=torch.float32, device=q.device)
    raw = torch.empty_like(z) if return_raw else None
    check(_lib.load().clpk_dequant_l2norm_u8(ptr(q), ptr(_f32c(scale)), ptr(_f32c(zero)), ptr(z), ptr(raw), b, d,
                                             int(l2norm), stream_ptr()), "clpk_dequant_l2norm_u8")
    return (z, raw) if return_raw else z


@on_tensor_device
def quant_encode(x: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor) -> torch.Tensor:
    require_cuda(x, scale, zero)
    x2 = _f32c(x).reshape(-1, x.shape[-1])
    q = torch.empty(x2.shape, dtype=torch.uint8, device=x.device)
    check(_lib.load().clpk_quant_encode_u8(ptr(x2), ptr(_f32c(scale)), ptr(_f32c(zero)), ptr(q), x2.shape[0],
                                           x2.shape[1], stream_ptr()), "clpk_quant_encode_u8")
    return q.reshape(x.shape)


@on_tensor_device
def quant_fit(x: torch.Tensor):
    require_cuda(x)
    x = _f32c(x)
    n, d = x.shape
    scale = torch.empty(d, dtype=torch.float32, device=x.device)
    zero = torch.empty_like(scale)
    check(_lib.load().clpk_quant_fit(ptr(x), ptr(scale), ptr(zero), n, d, stream_ptr()), "clpk_quant_fit")
    return scale, zero


@on_tensor_device
def ddim_step(x: torch.Tensor, eps: torch.Tensor, coef, noise: torch.Tensor | None = None,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """One DDIM update (ddim.py:36-45); coef = (sqrt(1-a_t), sqrt(a_t), sqrt(a_s), sqrt(a_s-sigma^2), sigma)."""
    require_cuda(x, eps, noise)
    x, eps = _f32c(x), _f32c(eps)
    noise = _f32c(noise) if noise is not None else None
    out = torch.empty_like(x) if out is None else out
    c5 = (C.c_float * 5)(*[float(v) for v in coef])
    check(_lib.load().clpk_ddim_step(ptr(x), ptr(eps), ptr(noise), c5, ptr(out), x.numel(), stream_ptr()),
          "clpk_ddim_step")
    return out


@functools.lru_cache(maxsize=32)
def timestep_frequencies(dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """The reference's frequency table (unet.py:33-34), evaluated by torch on the CPU exactly as the reference writes it:
    float32 scalar * arange / half, then torch.exp.  CPU tensor [dim // 2]."""
    half = dim // 2
    return torch.exp(-math.log(max_period) * torch.arange(0, half) / half)


@on_tensor_device
def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0, host_freqs: bool = True) -> torch.Tensor:
    """unet.py:22-39.  host_freqs (default): the frequencies come from the torch-CPU table above, so the embedding equals
    the reference's CPU result to ~1e-6; False: the device's own expf (what the reference computes when t is on CUDA)."""
    require_cuda(t)
    t = t.contiguous().to(torch.int64)
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    if host_freqs:
        f = timestep_frequencies(dim, float(max_period)).to(t.device)
        check(_lib.load().clpk_timestep_embedding_table(ptr(t), ptr(f), ptr(out), t.shape[0], dim, stream_ptr()),
              "clpk_timestep_embedding_table")
    else:
        check(_lib.load().clpk_timestep_embedding(ptr(t), ptr(out), t.shape[0], dim, float(max_period), stream_ptr()),
              "clpk_timestep_embedding")
    return out


@on_tensor_device
def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, act: int = 0,
           add: torch.Tensor | None = None) -> torch.Tensor:
    require_cuda(x, w)
    x, w = _f32c(x), _f32c(w)
    m, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k
    y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    check(_lib.load().clpk_linear(ptr(x), ptr(w), ptr(_f32c(b)) if b is not None else None,
                                  ptr(_f32c(add)) if add is not None else None, ptr(y), m, n, k, act, stream_ptr()),
          "clpk_linear")
    return y


@on_tensor_device
def film_apply(x_nchw: torch.Tensor, scale1p: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    require_cuda(x_nchw, scale1p, shift)
    x = _f32c(x_nchw)
    b, c = x.shape[:2]
    hw = x.numel() // (b * c)
    y = torch.empty_like(x)
    check(_lib.load().clpk_film_apply(ptr(x), ptr(_f32c(scale1p)), ptr(_f32c(shift)), ptr(y), b, c, hw, stream_ptr()),
          "clpk_film_apply")
    return y


@on_tensor_device
def groupnorm_silu(x_nhwc: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float = 1e-5,
                   silu: bool = True, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """fp32 NHWC [B,H,W,C] -> 16-bit (fp16 default / bf16) NHWC GroupNorm(+SiLU) = the conv A operand."""
    require_cuda(x_nhwc, gamma, beta)
    x = _f32c(x_nhwc)
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    lib = _lib.load()
    ws = torch.empty(int(lib.clpk_groupnorm_ws_bytes(b, hw, c, groups)), dtype=torch.uint8, device=x.device)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(lib.clpk_groupnorm_silu(ptr(x), ptr(_f32c(gamma)), ptr(_f32c(beta)), ptr(y), ptr(ws), b, hw, c, groups,
                                  float(eps), int(silu), op_code(dtype), stream_ptr()), "clpk_groupnorm_silu")
    return y


@on_tensor_device
def pack_conv_weight(w: torch.Tensor, kind: int, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Reference-layout fp32 conv weight -> 16-bit K-major GEMM layout (see include/clpk.h)."""
    require_cuda(w)
    w = _f32c(w)
    if kind == CONVT_4X4_S2:
        cin, cout = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[0], w.shape[1]
    lib = _lib.load()
    n = int(lib.clpk_pack_conv_weight(None, None, kind, cin, cout, op_code(dtype), None))
    if n < 0:
        check(1, "clpk_pack_conv_weight")
    out = torch.empty(n, dtype=dtype, device=w.device)
    if lib.clpk_pack_conv_weight(ptr(w), ptr(out), kind, cin, cout, op_code(dtype), stream_ptr()) < 0:
        check(2, "clpk_pack_conv_weight")
    return out


def _conv(fn_name: str, x_nhwc_op, w_packed, kind, cout, bias, film_scale1p, film_shift, resid, out_f32, out_op,
          out_nchw, gn_partial=None, gn_cpg=0, in_affine=None):
    require_cuda(x_nhwc_op, w_packed, bias)
    assert x_nhwc_op.is_contiguous() and x_nhwc_op.dtype == w_packed.dtype, "operands must share one 16-bit dtype"
    b, h, w, cin = x_nhwc_op.shape
    ep = ConvEpilogue()
    keep = [_f32c(bias)]
    ep.bias = ptr(keep[0])
    if film_scale1p is not None:
        fs, fb = _f32c(film_scale1p), _f32c(film_shift)
        keep += [fs, fb]
        ep.film_scale1p, ep.film_shift, ep.film_stride = ptr(fs), ptr(fb), fs.stride(0)
    if resid is not None and resid.dtype in (torch.float16, torch.bfloat16):
        assert resid.dtype == x_nhwc_op.dtype and resid.is_contiguous(), "a 16-bit residual must be in the operand dtype"
        ep.resid_op = ptr(resid)      # residual stream kept in 16 bits (clpk_conv_epilogue.resid_op)
    else:
        ep.resid = ptr(resid)
    ep.out_f32 = ptr(out_f32)
    ep.out_op = ptr(out_op)
    ep.out_nchw = ptr(out_nchw)
    ep.cout_valid = cout
    ep.gn_partial = ptr(gn_partial)
    ep.gn_cpg = int(gn_cpg)
    if in_affine is not None:
        isc, ish, act = in_affine
        isc, ish = _f32c(isc), _f32c(ish)
        assert isc.shape == (b, cin) and ish.shape == (b, cin)
        keep += [isc, ish]
        ep.in_scale, ep.in_shift, ep.in_silu = ptr(isc), ptr(ish), int(bool(act))
    fn = getattr(_lib.load(), fn_name)
    check(fn(ptr(x_nhwc_op), ptr(w_packed), kind, b, h, w, cin, cout, op_code(x_nhwc_op.dtype), C.byref(ep), stream_ptr()),
          fn_name)


def _conv_out_hw(kind: int, h: int, w: int):
    if kind == CONV_3X3_S2:
        return h // 2, w // 2
    if kind == CONVT_4X4_S2:
        return 2 * h, 2 * w
    return h, w


@on_tensor_device
def conv_igemm(x_nhwc_op: torch.Tensor, w_packed: torch.Tensor, kind: int, cout: int, bias: torch.Tensor, *,
               film_scale1p=None, film_shift=None, resid=None, want_f32=True, want_op=False, want_nchw=False,
               gn_groups: int = 0, impl: str = "igemm", in_affine=None, return_partial: bool = False,
               inplace: bool = False):
    """Implicit-GEMM conv on the tensor cores (x and w_packed in the same 16-bit dtype).  Returns a dict of the
    requested outputs: "f32" NHWC fp32, "op" NHWC in the operand dtype, "nchw" fp32 NCHW.  gn_groups > 0 additionally
    returns "gn_stats" [B, gn_groups, 2] = (mean, rstd) of a GroupNorm(gn_groups) over the output, accumulated by the
    conv epilogue (no extra pass over the tensor).  in_affine = (scale [B, Cin], shift [B, Cin], silu) applies
    a <- act(x * scale + shift) to the A operand inside the kernel (clpk_conv_epilogue.in_scale; geometries with
    conv_in_affine_supported(...) only).  return_partial additionally returns the raw per-tile statistics ("gn_partial",
    "gn_slots") for groupnorm_affine.  A `resid` in the operand dtype is added as a 16-bit residual (resid_op: the
    16-bit-only residual stream of the plan); with inplace=True the result overwrites it."""
    b, h, w, _ = x_nhwc_op.shape
    oh, ow = _conv_out_hw(kind, h, w)
    dev = x_nhwc_op.device
    outs = {}
    if want_f32:
        outs["f32"] = torch.empty((b, oh, ow, cout), dtype=torch.float32, device=dev)
    if want_op:
        outs["op"] = torch.empty((b, oh, ow, cout), dtype=x_nhwc_op.dtype, device=dev)
    if want_nchw:
        outs["nchw"] = torch.empty((b, cout, oh, ow), dtype=torch.float32, device=dev)
    resid16 = None
    if resid is not None and resid.dtype in (torch.float16, torch.bfloat16):
        # 16-bit residual: needs a 16-bit-only NHWC output; inplace=True updates the residual tensor itself
        assert want_op and not want_f32 and not want_nchw, "a 16-bit residual needs want_op=True, want_f32=False"
        resid16 = resid.contiguous()
        if inplace:
            outs["op"] = resid16
    partial, slots, cpg = None, 0, 0
    if gn_groups > 0:
        cpg = cout // gn_groups
        slots = int(_lib.load().clpk_conv_gn_slots(kind, h, w, cout, cpg))
        if slots <= 0:
            raise ValueError(f"fused GroupNorm statistics unsupported for cout={cout}, groups={gn_groups}")
        # [b][slots][groups] (tile mean, tile M2) pairs followed by [slots] element counts (see include/clpk.h)
        partial = torch.zeros(2 * b * slots * gn_groups + slots, dtype=torch.float32, device=dev)
    _conv("clpk_conv_igemm" if impl == "igemm" else "clpk_conv_direct", x_nhwc_op, w_packed, kind, cout, bias,
          film_scale1p, film_shift, resid16 if resid16 is not None else (_f32c(resid) if resid is not None else None),
          outs.get("f32"), outs.get("op"),
          outs.get("nchw"), partial, cpg, in_affine)
    if gn_groups > 0 and return_partial:
        outs["gn_partial"], outs["gn_slots"] = partial, slots
    if gn_groups > 0:
        stats = torch.empty((b, gn_groups, 2), dtype=torch.float32, device=dev)
        check(_lib.load().clpk_groupnorm_finalize(ptr(partial), ptr(stats), b, slots, gn_groups, float(oh * ow * cpg), 1e-5,
                                                  stream_ptr()), "clpk_groupnorm_finalize")
        outs["gn_stats"] = stats
    return outs


def conv_in_affine_supported(kind: int, h: int, w: int, cin: int, cout: int) -> bool:
    return bool(_lib.load().clpk_conv_in_affine_supported(kind, h, w, cin, cout))


@on_tensor_device
def groupnorm_affine(partial: torch.Tensor, batch: int, slots: int, gamma: torch.Tensor, beta: torch.Tensor, groups: int,
                     eps: float = 1e-5):
    """Per-tile conv statistics (the flat "gn_partial" buffer of conv_igemm(..., return_partial=True)) -> (scale, shift)
    [B, C] of GroupNorm(groups) in affine form."""
    require_cuda(partial, gamma, beta)
    c = gamma.numel()
    pieces = (partial.numel() - slots) // (2 * batch * slots)
    scale = torch.empty((batch, c), dtype=torch.float32, device=partial.device)
    shift = torch.empty_like(scale)
    check(_lib.load().clpk_groupnorm_affine(ptr(partial), ptr(_f32c(gamma)), ptr(_f32c(beta)), ptr(scale), ptr(shift), batch,
                                            slots, pieces, groups, c, eps, stream_ptr()), "clpk_groupnorm_affine")
    return scale, shift


@on_tensor_device
def groupnorm_apply(x_nhwc: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats: torch.Tensor, groups: int,
                    silu: bool = True, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Second half of GroupNorm(+SiLU) given (mean, rstd) [B, groups, 2] (e.g. from conv_igemm(gn_groups=...))."""
    require_cuda(x_nhwc, gamma, beta, stats)
    x = _f32c(x_nhwc)
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(_lib.load().clpk_groupnorm_apply(ptr(x), ptr(_f32c(gamma)), ptr(_f32c(beta)), ptr(_f32c(stats)), ptr(y), b, hw, c,
                                           groups, int(silu), op_code(dtype), stream_ptr()), "clpk_groupnorm_apply")
    return y


@on_tensor_device
def head_conv(x_nhwc_op: torch.Tensor, in_scale: torch.Tensor, in_shift: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """out(out_norm(x)) of the UNet head as one kernel (see clpk_head_conv): x 16-bit NHWC, (in_scale, in_shift) [B, C] the
    GroupNorm in affine form, w [3, C, 3, 3] / b [3] the reference-layout fp32 conv parameters.  Returns fp32 NCHW."""
    require_cuda(x_nhwc_op, in_scale, in_shift, w, b)
    bsz, h, wd, c = x_nhwc_op.shape
    lib = _lib.load()
    if not lib.clpk_head_conv_supported(h, wd, c, w.shape[0]):
        raise ValueError(f"fused head kernel unsupported for W={wd}, C={c}, Cout={w.shape[0]}")
    wp = torch.empty((32, c), dtype=x_nhwc_op.dtype, device=x_nhwc_op.device)
    check(lib.clpk_pack_head_weight(ptr(_f32c(w)), ptr(wp), c, op_code(x_nhwc_op.dtype), stream_ptr()), "clpk_pack_head_weight")
    out = torch.empty((bsz, 3, h, wd), dtype=torch.float32, device=x_nhwc_op.device)
    keep = [_f32c(in_scale), _f32c(in_shift), _f32c(b), x_nhwc_op.contiguous()]
    check(lib.clpk_head_conv(ptr(keep[3]), ptr(keep[0]), ptr(keep[1]), ptr(wp), ptr(keep[2]), ptr(out), bsz, h, wd, c,
                             op_code(x_nhwc_op.dtype), stream_ptr()), "clpk_head_conv")
    return out


def conv_direct(*args, **kwargs):
    """CUDA-core evaluation of the same contract (on-device cross-check for tests)."""
    return conv_igemm(*args, impl="direct", **kwargs)


@on_tensor_device
def conv_in(x_nchw: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Stem conv (unet.py:55): fp32 NCHW -> fp32 NHWC."""
    require_cuda(x_nchw, w, b)
    x = _f32c(x_nchw)
    bsz, cin, h, wd = x.shape
    cout = w.shape[0]
    y = torch.empty((bsz, h, wd, cout), dtype=torch.float32, device=x.device)
    check(_lib.load().clpk_conv_in(ptr(x), ptr(_f32c(w)), ptr(_f32c(b)), ptr(y), bsz, cin, h, wd, cout, stream_ptr()),
          "clpk_conv_in")
    return y


@on_tensor_device
def to_uint8_hwc(x_nchw: torch.Tensor) -> torch.Tensor:
    """clamp(-1,1) -> ((x+1)*127.5) truncated to uint8, HWC (reconstruct_diffusion.py:55-56)."""
    require_cuda(x_nchw)
    x = _f32c(x_nchw)
    b, c, h, w = x.shape
    out = torch.empty((b, h, w, c), dtype=torch.uint8, device=x.device)
    check(_lib.load().clpk_to_uint8_hwc(ptr(x), ptr(out), b, c, h, w, stream_ptr()), "clpk_to_uint8_hwc")
    return out


@on_tensor_device
def psnr_sqerr_u8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per-image sum of squared uint8-domain differences (metrics.py:16-26) as int64 [B]."""
    require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    assert a.shape == b.shape
    bsz = a.shape[0]
    per = a.numel() // bsz
    out = torch.empty(bsz, dtype=torch.int64, device=a.device)
    check(_lib.load().clpk_psnr_sqerr_u8(ptr(a), ptr(b), ptr(out), bsz, per, stream_ptr()), "clpk_psnr_sqerr_u8")
    return out


@on_tensor_device
def ssim_u8(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per-image SSIM (metrics.py:32-46, scikit-image defaults: 7x7 uniform window, uint8 domain, data_range 255) of two
    fp32 NCHW [B,C,H,W] tensors in [-1,1], as fp64 [B]."""
    require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    assert a.shape == b.shape and a.dim() == 4
    bsz, ch, h, w = a.shape
    lib = _lib.load()
    nbytes = int(lib.clpk_ssim_ws_bytes(bsz, ch, h, w))
    if nbytes < 0:
        raise ValueError("win_size exceeds image extent. Either ensure that your images are at least 7x7; or pass "
                         "win_size explicitly in the function call, with an odd value less than or equal to the smaller "
                         "side of your images.")
    ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=a.device)
    out = torch.empty(bsz, dtype=torch.float64, device=a.device)
    check(lib.clpk_ssim_u8(ptr(a), ptr(b), ptr(out), ptr(ws), bsz, ch, h, w, stream_ptr()), "clpk_ssim_u8")
    return out


@on_tensor_device
def ddpm_combine(x: torch.Tensor, y: torch.Tensor, ca: torch.Tensor, cb: torch.Tensor, cdiv: torch.Tensor | None = None,
                 clamp: bool = False) -> torch.Tensor:
    """out[b] = (ca[b]*x[b] + cb[b]*y[b]) [/ cdiv[b]] [clamp(-1, 1)] with every operation individually rounded — the DDPM
    helpers of the reference scheduler (scheduler.py:46-68) as one fused pass (see clpk.h)."""
    require_cuda(x, y, ca, cb)
    x, y = _f32c(x), _f32c(y)
    assert x.shape == y.shape and x.dim() >= 1
    bsz = x.shape[0]
    ca, cb = _f32c(ca).reshape(bsz), _f32c(cb).reshape(bsz)
    cd = _f32c(cdiv).reshape(bsz) if cdiv is not None else None
    out = torch.empty_like(x)
    check(_lib.load().clpk_ddpm_combine(ptr(x), ptr(y), ptr(ca), ptr(cb), ptr(cd), ptr(out), bsz, x.numel() // bsz,
                                        1 if clamp else 0, stream_ptr()), "clpk_ddpm_combine")
    return out
