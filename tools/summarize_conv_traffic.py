"""profiles/conv_traffic_<tag>.txt from an `ncu --metrics ... --csv` log over the 36 conv_igemm launches of one DDIM step.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none \
        -k regex:conv_igemm -s 216 -c 36 --csv --log-file gpurun_out/convs_r1.csv \
        python bench.py --steps 1 --warmup 3 --ddim-steps 2 --no-cpu-baseline
    python tools/summarize_conv_traffic.py gpurun_out/convs_r1.csv profiles/ r1
"""
from __future__ import annotations

import collections
import csv
import sys
from pathlib import Path

# launch order of one forward: stem, [conv1, conv2] x2 + down per level, mid x2, [conv1, conv2] x2 + up per level, out
RES_IDX = [1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14, 16, 17, 18, 19, 20, 21, 22, 23, 25, 26, 27, 28, 30, 31, 32, 33]


def main():
    src, outdir, tag = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3]
    rows = list(csv.DictReader([ln for ln in src.read_text().splitlines() if not ln.startswith("==")]))
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"].replace("void clpk::", "")[:28], "grid": r["Grid Size"]})
        v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
        if "bytes" in m:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if "time" in m:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        d[m] = v
    import subprocess
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = ["# ncu --metrics (time, DRAM bytes, L2 bytes, tensor-pipe activity) --clock-control none over the 36 conv launches of ONE DDIM step",
           f"# command: python bench.py --steps 1 --warmup 3 --ddim-steps 2 --no-cpu-baseline --no-extras   (-k regex:conv_igemm|head_conv -s 216 -c 36); kernels of commit {head}",
           "# launch order = stem, [ResBlock conv1, conv2] x2 + down per level, mid x2, [conv1, conv2] x2 + up per level, head (#35: head_conv_kernel = out_norm + out)",
           f"{'#':>2s} {'kernel':28s} {'grid':12s} {'us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'l2_MB':>9s} {'tensor%':>8s}"]
    items = list(by.values())
    tt = 0.0
    for i, d in enumerate(items):
        t = d["gpu__time_duration.sum"]
        tt += t
        out.append(f"{i:2d} {d['name']:28s} {d['grid']:12s} {t:8.1f} {d['dram__bytes_read.sum'] / 1e6:10.1f} "
                   f"{d['dram__bytes_write.sum'] / 1e6:10.1f} {d['lts__t_bytes.sum'] / 1e6:9.1f} "
                   f"{d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:8.1f}")
    rb = [items[i] for i in RES_IDX if i < len(items)]
    rr = sum(d["dram__bytes_read.sum"] for d in rb) / len(rb)
    ww = sum(d["dram__bytes_write.sum"] for d in rb) / len(rb)
    out.append(f"# total {tt:.1f} us;  the 28 ResBlock convs: mean DRAM read {rr / 1e6:.1f} MB + write {ww / 1e6:.1f} MB = "
               f"{(rr + ww) / 1e6:.1f} MB per launch, mean {sum(d['gpu__time_duration.sum'] for d in rb) / len(rb):.1f} us")
    (outdir / f"conv_traffic_{tag}.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
