"""profiles/ncu_full_<tag>.txt from the raw-page csv of an `ncu --set full` capture of one whole UNet forward
(tools/collect_profiles.sh step 4):  python tools/summarize_full.py gpurun_out/full_r2.csv profiles/ r2"""
from __future__ import annotations

import csv
import re
import sys
from pathlib import Path

COLS = [("us", "gpu__time_duration.sum", 1.0), ("regs", "launch__registers_per_thread", 1.0),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("dram_rd_MB", "dram__bytes_read.sum", 1.0), ("dram_wr_MB", "dram__bytes_write.sum", 1.0),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("L2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("smemLSU%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
        ("ipc", "smsp__inst_executed.avg.per_cycle_active", 1.0)]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "membar", "not_selected", "no_instruction", "branch_resolving", "sleeping", "dispatch_stall"]


def to_unit(v: str, unit: str, want: str) -> float:
    x = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    if want == "us":
        return x / 1000 if unit == "ns" else x * 1000 if unit == "ms" else x
    if want.endswith("MB"):
        return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0) * x
    return x


def main() -> None:
    src, outdir, tag = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(src.read_text().splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full --clock-control none over ONE whole UNet forward + DDIM update ({tag}; tools/collect_profiles.sh),",
           "# command: python bench.py --steps 1 --warmup 3 --ddim-steps 2 --no-cpu-baseline --no-extras; cold-cache, serialised launches.",
           "# stall = the two largest warp-issue stall reasons (warps per issue-active cycle).", "",
           f"{'#':>3s} {'kernel':34s} {'grid':>12s} " + " ".join(f"{n:>10s}" for n, _, _ in COLS) + "  stall"]
    tot = 0.0
    for k, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("clpk::", "")
        vals = []
        for n, key, _ in COLS:
            vals.append(to_unit(r[col[key]], units[col[key]], n) if key in col else float("nan"))
        tot += vals[0]
        st = []
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in col:
                try:
                    st.append((float(r[col[key]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        out.append(f"{k:3d} {name[:34]:34s} {r[col['Grid Size']].replace(' ', ''):>12s} " + " ".join(f"{v:10.1f}" for v in vals) +
                   "  " + ", ".join(f"{s} {v:.1f}" for v, s in st[:2]))
    out.append(f"# total {tot:.1f} us over {len(data)} launches")
    (outdir / f"ncu_full_{tag}.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
