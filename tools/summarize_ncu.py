"""Turns the ncu artefacts of a gpurun call into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep profiles/ r1
"""
from __future__ import annotations

import collections
import csv
import re
import subprocess
import sys
from pathlib import Path


def launch_list(path: Path):
    lines = [ln for ln in path.read_text().splitlines() if not ln.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("clpk::", "")
        seq.append((name, row["Grid Size"], row["Block Size"], v))
    return seq


def main():
    launches, rep, outdir, tag = Path(sys.argv[1]), Path(sys.argv[2]), Path(sys.argv[3]), sys.argv[4]
    outdir.mkdir(exist_ok=True)
    seq = launch_list(launches)
    agg = collections.OrderedDict()
    for n, g, b, v in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over "
           f"`python bench.py --steps 1 --warmup 3 --ddim-steps 2 --no-cpu-baseline --no-extras` ({len(seq)} launches)",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
           f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{k[:58]:58s} {n:8d} {t:10.1f} {t / tot * 100:6.1f}% {t / n:9.1f}")
    out.append(f"{'TOTAL':58s} {len(seq):8d} {tot:10.1f}")
    idx = [i for i, s in enumerate(seq) if "stem_im2col" in s[0]]
    if len(idx) >= 2:
        out += ["", "# one UNet forward + DDIM update in launch order (default config, batch 8):"]
        for n, g, b, v in seq[idx[0]:idx[1]]:
            out.append(f"{n[:44]:44s} grid={g:14s} block={b:12s} {v:8.1f} us")
        out.append(f"# forward total {sum(s[3] for s in seq[idx[0]:idx[1]]):.1f} us in {idx[1] - idx[0]} launches")
    (outdir / f"launches_{tag}.txt").write_text("\n".join(out) + "\n")

    if str(rep) == "-":   # launch list only (the per-kernel table comes from tools/summarize_full.py)
        print((outdir / f"launches_{tag}.txt").read_text()[:3000])
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
            "launch__cluster_dim_x", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full --clock-control none --import-source on ({tag}); one column per captured launch", ""]
    for w in want:
        if w in col:
            i = col[w]
            vals = [re.sub(r"\(.*", "", r[i]).replace("void clpk::", "")[:24] if w == "Kernel Name" else r[i] for r in data]
            out.append(f"{w[:70]:70s} {units[i][:10]:10s} " + " | ".join(f"{v:>14s}" for v in vals))
    (outdir / f"ncu_full_{tag}.txt").write_text("\n".join(out) + "\n")
    print((outdir / f"launches_{tag}.txt").read_text()[:3000])
    print((outdir / f"ncu_full_{tag}.txt").read_text())


if __name__ == "__main__":
    main()
