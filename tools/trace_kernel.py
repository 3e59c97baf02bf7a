"""Debug-build only (CLPK_NVCC_EXTRA=-DCLPK_IGEMM_DEBUG): kernel-level phase marks (globaltimer, ns) of the first and the
last CTA pair of one conv launch, next to the CUDA-event time of the launch.  python tools/trace_kernel.py "rb128 conv1" """
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ["CLPK_IGEMM_DBG"] = "128"
from clip_neural_image_conpression_b200 import _lib, ops  # noqa: E402
from bench_conv import LAYERS  # noqa: E402

NAMES = {300: "entry", 301: "setup_done", 302: "first_operands", 303: "tile_mma_issued", 304: "epilogue_done", 305: "exit"}


def main():
    for want in sys.argv[1:] or ["rb128 conv1"]:
        name, kind, h, w, cin, cout, film, resid = [l for l in LAYERS if l[0] == want][0]
        dev = torch.device("cuda")
        b = 8
        x = torch.randn(b, h, w, cin, device=dev).to(torch.float16)
        wt = (torch.randn(cin, cout, 4, 4, device=dev) if kind == 2 else torch.randn(cout, cin, 3, 3, device=dev)) * 0.03
        wp = ops.pack_conv_weight(wt, kind)
        bias = torch.randn(cout, device=dev)
        oh, ow = (h // 2, w // 2) if kind == 1 else ((2 * h, 2 * w) if kind == 2 else (h, w))
        kw = {}
        if film:
            kw.update(film_scale1p=torch.ones(b, cout, device=dev), film_shift=torch.zeros(b, cout, device=dev))
        if resid:
            kw["resid"] = torch.randn(b, oh, ow, cout, device=dev)
        lib = _lib.load()
        lib.clpk_debug_trace.restype = C.c_int
        buf = (C.c_longlong * 8192)()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for it in range(4):
            flush.zero_()
            torch.cuda.synchronize()
            lib.clpk_debug_trace(buf, 4096)  # reset
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_igemm(x, wp, kind, cout, bias, want_f32=not film, want_op=film, gn_groups=0, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            n = lib.clpk_debug_trace(buf, 4096)
        ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n))
        t0 = ev[0][0]
        print(f"{want}: event time {ms * 1e3:.1f} us, {n} marks (ns since the first mark):")
        first = {}
        last = {}
        for t, tag in ev:
            first.setdefault(tag, t - t0)
            last[tag] = t - t0
        for tag in sorted(NAMES):
            if tag in first:
                print(f"   {NAMES[tag]:16s} first {first[tag] / 1e3:8.2f} us   last {last[tag] / 1e3:8.2f} us")


if __name__ == "__main__":
    main()
