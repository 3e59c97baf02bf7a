"""Debug-build only (CLPK_NVCC_EXTRA=-DCLPK_IGEMM_DEBUG, CLPK_IGEMM_DBG=64): per-phase clock64 trace of one epilogue
warp and of the MMA issuer of CTA 0 for one conv launch.  python tools/trace_conv.py "rb256 conv2" """
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["CLPK_IGEMM_DBG"] = "64"
from clip_neural_image_conpression_b200 import _lib, ops  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parent))
from bench_conv import LAYERS  # noqa: E402


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "rb256 conv2"
    name, kind, h, w, cin, cout, film, resid = [l for l in LAYERS if l[0] == want][0]
    dev = torch.device("cuda")
    b = 8
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
        kw["resid"] = torch.randn(b, oh, ow, cout, device=dev).to(torch.float16)   # 16-bit residual stream (plan default)
    lib = _lib.load()
    lib.clpk_debug_trace.restype = C.c_int
    buf = (C.c_longlong * 8192)()
    for it in range(3):
        ops.conv_igemm(x, wp, kind, cout, bias, want_f32=False, want_op=True, gn_groups=8, **kw)
        n = lib.clpk_debug_trace(buf, 4096)
    ev = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
    epi = [(t, c) for t, c in ev if t < 200]
    mma = [(t, c) for t, c in ev if t >= 200]
    t0 = min(c for _, c in ev)
    names = {100: "tile_top", 101: "tmem_full", 102: "tmem_ld", 103: "math", 104: "res_full", 105: "sts", 106: "fence",
             107: "lane0_tma", 108: "stats", 109: "tile_end", 200: "mma_top", 201: "tmem_empty"}
    print(f"{want}: {n} trace points")
    prev = None
    import collections
    dur = collections.defaultdict(list)
    for t, c in epi:
        if prev is not None:
            dur[names[t]].append(c - prev)
        prev = c
    for k, v in dur.items():
        v2 = v[len(v) // 4:]  # skip warm-up tiles
        print(f"  epilogue -> {k:10s} n={len(v):4d} mean {sum(v2) / len(v2):8.0f} clk   min {min(v2):6d} max {max(v2):6d}")
    prev = None
    dm = collections.defaultdict(list)
    for t, c in mma:
        if prev is not None:
            dm[names[t]].append(c - prev)
        prev = c
    for k, v in dm.items():
        v2 = v[len(v) // 4:]
        print(f"  mma      -> {k:10s} n={len(v):4d} mean {sum(v2) / len(v2):8.0f} clk   min {min(v2):6d} max {max(v2):6d}")
    tiles = [c for t, c in epi if t == 100]
    if len(tiles) > 4:
        print(f"  epilogue tile period: {(tiles[-1] - tiles[2]) / (len(tiles) - 3):.0f} clk")


if __name__ == "__main__":
    main()
