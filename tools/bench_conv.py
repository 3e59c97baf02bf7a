"""Per-layer microbenchmark of the tcgen05 conv kernel on the default-config layer shapes (CUDA events, L2 flushed
between iterations).  python tools/bench_conv.py [--batch 8] [--iters 10]"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from clip_neural_image_conpression_b200 import ops  # noqa: E402

LAYERS = [  # name, kind, H, W, Cin, Cout, film, resid
    ("rb256 conv1", 0, 256, 256, 128, 128, True, False),
    ("rb256 conv2", 0, 256, 256, 128, 128, False, True),
    ("rb128 conv1", 0, 128, 128, 128, 128, True, False),
    ("rb128 conv2", 0, 128, 128, 128, 128, False, True),
    ("rb64  conv1", 0, 64, 64, 256, 256, True, False),
    ("rb64  conv2", 0, 64, 64, 256, 256, False, True),
    ("rb32  conv1", 0, 32, 32, 512, 512, True, False),
    ("rb32  conv2", 0, 32, 32, 512, 512, False, True),
    ("down 256->128", 1, 256, 256, 128, 128, False, False),
    ("down 64->32", 1, 64, 64, 256, 512, False, False),
    ("up 32->64", 2, 32, 32, 512, 256, False, True),
    ("up 128->256", 2, 128, 128, 128, 128, False, True),
    ("out 128->3", 0, 256, 256, 128, 3, False, False),
    ("stem 32->128", 3, 256, 256, 32, 128, False, False),   # pointwise GEMM over the im2col columns
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", type=str, default="")
    ap.add_argument("--fp32-stream", dest="res16", action="store_false", help="model-like epilogues of the fp32 residual stream")
    ap.add_argument("--model-like", action="store_true", help="epilogues as the plan uses them (16-bit conv1 output, fused GN stats)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"CLPK_IGEMM_DBG={os.environ.get('CLPK_IGEMM_DBG', '0')} batch={args.batch}")
    tot_ms = tot_fl = 0.0
    for name, kind, h, w, cin, cout, film, resid in LAYERS:
        if args.only and args.only not in name:
            continue
        b = args.batch
        x = torch.randn(b, h, w, cin, device=dev).to(torch.float16)
        wt = (torch.randn(cin, cout, 4, 4, device=dev) if kind == 2 else torch.randn(cout, cin, 1, 1, device=dev) if kind == 3
              else torch.randn(cout, cin, 3, 3, device=dev)) * 0.03
        wp = ops.pack_conv_weight(wt, kind)
        bias = torch.randn(cout, device=dev)
        oh, ow = (h // 2, w // 2) if kind == 1 else ((2 * h, 2 * w) if kind == 2 else (h, w))
        kw = {}
        if film:
            kw.update(film_scale1p=torch.ones(b, cout, device=dev), film_shift=torch.zeros(b, cout, device=dev))
        if resid:
            kw["resid"] = torch.randn(b, oh, ow, cout, device=dev)
        nchw = cout % 16 != 0
        if args.model_like and not nchw:
            # what the plan launches: conv1 keeps only the 16-bit copy, every producer emits GroupNorm partials; with the
            # 16-bit residual stream (default) every conv stores 16 bits only and a residual arrives as a 16-bit tile
            only16 = film or args.res16
            if resid and args.res16:
                kw["resid"] = kw["resid"].to(torch.float16)
            run = lambda: ops.conv_igemm(x, wp, kind, cout, bias, want_f32=not only16, want_op=only16, gn_groups=8, **kw)  # noqa: E731
        else:
            run = lambda: ops.conv_igemm(x, wp, kind, cout, bias, want_f32=not nchw, want_nchw=nchw, **kw)  # noqa: E731
        for _ in range(3):
            run()
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        taps = 16 if kind == 2 else 9
        pix = b * (h * w if kind == 2 else oh * ow)
        fl = 2.0 * pix * cin * cout * taps
        tot_ms += med
        tot_fl += fl
        print(f"{name:14s} B{b} {h}x{w} {cin}->{cout}: {med * 1e3:8.1f} us  {fl / med / 1e9:7.1f} TFLOP/s")
    print(f"total {tot_ms:.3f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s aggregate")


if __name__ == "__main__":
    main()
