"""python -m clip_neural_image_conpression_b200.cli.eval — decode a whole store and report quality metrics.

Drop-in for the reference CLI (PKG/cli/eval.py:33-86): same flags, same printed lines, same JSON schema
({image, psnr, ssim, lpips, clip_sim} per image).  Differences, all opt-in or metric-neutral:
  * images are decoded in micro-batches (`--batch`, default 8) instead of one at a time;
  * under torchrun the manifest is partitioned contiguously over the ranks (one GPU each); PSNR sums are all-reduced
    and rank 0 prints — nothing is exchanged inside the DDIM loop;
  * the originals are decoded on the host (PIL) but BICUBIC-resized and converted on the device, bit-identically to
    Pillow (eval/resample.py);
  * PSNR and SSIM are computed on the device in the uint8 domain (SSIM = scikit-image's structural_similarity defaults,
    see eval/metrics.py); LPIPS / CLIP-similarity are NaN (lpips and open_clip need pretrained networks that are not
    part of this stack — the reference also reports NaN for LPIPS when the package is missing).
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np
import torch

from .. import parallel
from ..diffusion import DDIMSampler, NoiseScheduler
from ..eval.metrics import psnr_batch, ssim_batch
from ..eval.resample import load_original_device
from ..io.bitstream import read_bitstreams
from ..pipeline import decode_codes
from .reconstruct_diffusion import load_net, load_store_meta


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Evaluate reconstruction quality on a store of images (B200 path).")
    ap.add_argument("--store_dir", type=str, required=True)
    ap.add_argument("--weights", type=str, required=True)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--eta", type=float, default=0.0)
    ap.add_argument("--device", type=str, default="cuda")
    ap.add_argument("--out_json", type=str, default=None)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--base", type=int, default=128)
    ap.add_argument("--ch_mult", type=int, nargs="+", default=[1, 2, 2])
    ap.add_argument("--seed", type=int, default=None)
    return ap


def load_original(path: str, size: int) -> np.ndarray:
    """eval.py:66-67 on the host — RGB, BICUBIC resize, [-1,1], CHW (kept as the reference-shaped helper; the CLI itself
    uses eval.resample.load_original_device)."""
    from PIL import Image

    img = Image.open(path).convert("RGB").resize((size, size), Image.BICUBIC)
    return (np.array(img).astype(np.float32) / 127.5 - 1.0).transpose(2, 0, 1)


def _nanmean(vals) -> float:
    vals = [v for v in vals if not np.isnan(v)]
    return float(np.mean(vals)) if vals else float("nan")


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    rank, local, world = parallel.init_from_env()
    device = torch.device(args.device if world == 1 else f"cuda:{local}")
    if device.type != "cuda":
        raise SystemExit("this decoder runs on CUDA (sm_100a) only; there is no CPU path")
    torch.cuda.set_device(device if device.index is not None else torch.cuda.current_device())   # --device cuda:N is honoured (single process); under torchrun: cuda:LOCAL_RANK
    if args.seed is not None:
        torch.manual_seed(args.seed + rank)
    store = Path(args.store_dir)
    manifest = json.loads((store / "manifest.json").read_text(encoding="utf-8"))
    lo, hi = parallel.shard_bounds(len(manifest), rank, world)
    mine = manifest[lo:hi]
    scale, zero = load_store_meta(store, device)
    net = load_net(args.weights, scale.shape[0], args.base, args.ch_mult, device)
    sampler = DDIMSampler(NoiseScheduler(timesteps=1000, schedule="cosine", device=device), eta=args.eta)
    q = read_bitstreams([Path(r["bitstream"]) for r in mine]) if mine else np.zeros((0, scale.shape[0]), np.uint8)
    recon = decode_codes(net, sampler, q, scale, zero, args.size, steps=args.steps, batch=args.batch).clamp_(-1, 1)
    metrics = []
    for i in range(0, len(mine), 64):
        chunk = mine[i:i + 64]
        # eval.py:66-67 with the BICUBIC resize + float conversion on the device (bit-identical to Pillow / numpy)
        orig = torch.stack([load_original_device(r["image"], args.size, device) for r in chunk])
        rec = recon[i:i + len(chunk)]
        for r, p, ss in zip(chunk, psnr_batch(orig, rec), ssim_batch(orig, rec)):   # eval.py:69-70, on the device
            metrics.append({"image": r["image"], "psnr": p, "ssim": ss, "lpips": float("nan"), "clip_sim": float("nan")})
    finite = [m["psnr"] for m in metrics if np.isfinite(m["psnr"])]
    ssims = [m["ssim"] for m in metrics if not np.isnan(m["ssim"])]
    # mean over non-NaN like eval.py:77-79 (inf from identical images is kept out of the all-reduced sum)
    sums = parallel.reduce_sums([sum(finite), len(finite), sum(ssims), len(ssims)], device)
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, metrics)
        metrics = [m for part in gathered for m in part]
    if rank == 0:
        avg = float(sums[0] / sums[1]) if float(sums[1]) > 0 else _nanmean([m["psnr"] for m in metrics])
        print(f"Average PSNR: {avg:.2f} dB")
        avg_ssim = float(sums[2] / sums[3]) if float(sums[3]) > 0 else float("nan")
        print(f"Average SSIM: {avg_ssim:.4f}")
        print(f"Average LPIPS: {_nanmean([m['lpips'] for m in metrics]):.4f}")
        print(f"Average CLIP similarity: {_nanmean([m['clip_sim'] for m in metrics]):.4f}")
        if args.out_json:
            with open(args.out_json, "w", encoding="utf-8") as f:
                json.dump(metrics, f, ensure_ascii=False, indent=2)
    parallel.barrier()


if __name__ == "__main__":
    main()
