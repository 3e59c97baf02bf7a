"""python -m clip_neural_image_conpression_b200.cli.reconstruct_diffusion — one .clp bitstream -> one PNG.

Drop-in for the reference CLI (PKG/cli/reconstruct_diffusion.py:26-58): same flags, same hard-coded architecture
(base=128, ch_mult=(1,2,2), cosine schedule, T=1000), same post-processing, same "Saved to <out>" line.  The
dequantise + L2-renorm runs on the device, the DDIM loop inside libclpk.  `--device` must be a CUDA device.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch

from .. import ops
from ..diffusion import DDIMSampler, NoiseScheduler
from ..io.bitstream import read_bitstream
from ..models import CLIPCondUNet
from ..pipeline import decode_codes


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Reconstruct an image from a .clp bitstream via DDIM sampling (B200 path).")
    ap.add_argument("--store_dir", type=str, required=True)
    ap.add_argument("--bitstream", type=str, required=True)
    ap.add_argument("--weights", type=str, required=True)
    ap.add_argument("--out", type=str, default="recon.png")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--eta", type=float, default=0.0)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--device", type=str, default="cuda")
    # opt-in extras (defaults reproduce the reference)
    ap.add_argument("--base", type=int, default=128)
    ap.add_argument("--ch_mult", type=int, nargs="+", default=[1, 2, 2])
    ap.add_argument("--seed", type=int, default=None, help="seed torch's generator before drawing x_T")
    return ap


def load_store_meta(store_dir: Path, device):
    meta = np.load(store_dir / "codec_meta.npz")
    scale = torch.from_numpy(meta["scale"].astype("float32")).to(device)
    zero = torch.from_numpy(meta["zero"].astype("float32")).to(device)
    return scale, zero


def load_net(weights: str, z_dim: int, base: int, ch_mult, device) -> CLIPCondUNet:
    net = CLIPCondUNet(z_dim=z_dim, base=base, ch_mult=tuple(ch_mult), img_ch=3).to(device)
    net.load_state_dict(torch.load(weights, map_location=device), strict=True)
    return net.eval()


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    device = torch.device(args.device)
    if device.type != "cuda":
        raise SystemExit("this decoder runs on CUDA (sm_100a) only; there is no CPU path")
    torch.cuda.set_device(device if device.index is not None else torch.cuda.current_device())   # --device cuda:N: streams, plan memory and kernels all on that GPU
    if args.seed is not None:
        torch.manual_seed(args.seed)
    scale, zero = load_store_meta(Path(args.store_dir), device)
    q = read_bitstream(Path(args.bitstream))
    net = load_net(args.weights, scale.shape[0], args.base, args.ch_mult, device)
    sampler = DDIMSampler(NoiseScheduler(timesteps=1000, schedule="cosine", device=device), eta=args.eta)
    x = decode_codes(net, sampler, q[None, :], scale, zero, args.size, steps=args.steps, batch=1)
    img = ops.to_uint8_hwc(x)[0].cpu().numpy()
    from PIL import Image

    Image.fromarray(img).save(args.out)
    print(f"Saved to {args.out}")


if __name__ == "__main__":
    main()
