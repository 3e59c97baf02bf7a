"""Batched decode pipeline: uint8 codes (host) -> H2D -> dequantise + L2 renorm (device) -> DDIM -> images.

This is the body of the reference's eval loop (PKG/cli/eval.py:56-64) and of reconstruct_diffusion.main
(PKG/cli/reconstruct_diffusion.py:41-56), batched: the reference decodes one image at a time with B = 1.
`write_store` is the encode-side counterpart: the store-writing tail of PKG/cli/encode_images.py:75-87 (quantiser fit,
per-vector encode, .clp files, codec_meta.npz, manifest.json) with the quantiser on the device and a batched writer.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Optional, Sequence

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


def write_store(feats, image_paths: Sequence[str], out_dir, threads: int = 16) -> List[dict]:
    """Writes a feature store exactly as PKG/cli/encode_images.py:75-87 does once it has the CLIP features: fits the
    per-channel affine quantiser on ALL features (device kernel, bit-exact), saves `codec_meta.npz` (scale / zero fp32,
    dim int32), quantises every vector (one device launch for the whole [N, D] matrix instead of N host calls), writes
    one `<image stem>.clp` per vector (batched writer) and `manifest.json` (same schema and formatting).  `feats`: fp32
    [N, D] (numpy or torch, any device); returns the manifest."""
    from .codecs.quantizer import PerChannelAffineQuantizer
    from .io.bitstream import write_bitstreams

    require_cuda()
    x = torch.as_tensor(feats, dtype=torch.float32)
    if x.ndim != 2 or x.shape[0] != len(image_paths):
        raise ValueError(f"feats {tuple(x.shape)} do not match {len(image_paths)} image paths")
    if x.shape[0] == 0:
        raise SystemExit("No images encoded.")                     # encode_images.py:73-74
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    dim = int(x.shape[1])
    qzr = PerChannelAffineQuantizer(8).fit(x.cuda())
    np.savez(out / "codec_meta.npz", scale=qzr.scale.cpu().numpy().astype("float32"),
             zero=qzr.zero.cpu().numpy().astype("float32"), dim=np.int32(dim))
    codes = qzr.encode(x.cuda())                                   # uint8 [N, D], one kernel launch
    clp_paths = [out / (Path(p).stem + ".clp") for p in image_paths]
    write_bitstreams(codes, clp_paths, threads=threads)
    manifest = [{"image": str(p), "bitstream": str(c)} for p, c in zip(image_paths, clp_paths)]
    with open(out / "manifest.json", "w", encoding="utf-8") as f:
        json.dump(manifest, f, ensure_ascii=False, indent=2)
    return manifest
