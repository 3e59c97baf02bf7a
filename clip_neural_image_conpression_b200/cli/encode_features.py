"""Store writer for pre-computed CLIP features — the tail of the reference's `encode_images` CLI
(PKG/cli/encode_images.py:75-87) without its encoder: the CLIP image tower needs pretrained open_clip weights and is out
of this repo's scope, so the features come from a file.

    python -m clip_neural_image_conpression_b200.cli.encode_features --features feats.npy --image_list images.json \
        --out_dir store

`--features`: .npy fp32 [N, D]; `--image_list`: JSON list of N image paths (or a text file, one path per line).  Writes
`<stem>.clp` per vector, `codec_meta.npz` and `manifest.json` exactly as the reference does (quantiser fit / encode run on
the device, bit-exact) and prints the reference's closing line.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np

from ..pipeline import write_store


def main(argv=None) -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--features", type=str, required=True)
    ap.add_argument("--image_list", type=str, required=True)
    ap.add_argument("--out_dir", type=str, required=True)
    ap.add_argument("--threads", type=int, default=16)
    args = ap.parse_args(argv)
    feats = np.load(args.features).astype(np.float32)
    text = Path(args.image_list).read_text(encoding="utf-8")
    paths = json.loads(text) if text.lstrip().startswith("[") else [ln for ln in text.splitlines() if ln.strip()]
    manifest = write_store(feats, paths, args.out_dir, threads=args.threads)
    print(f"Done. Stored {len(manifest)} vectors in {Path(args.out_dir)}")


if __name__ == "__main__":
    main()
