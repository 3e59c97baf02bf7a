#!/usr/bin/env python
"""profiles/sass_opcodes_<tag>.txt: per-kernel counts of the Blackwell-native SASS mnemonics in the in-tree libclpk.so
(`cuobjdump -sass`; runs on the CPU-only build box).   python tools/sass_opcodes.py profiles/ r2"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "clip_neural_image_conpression_b200" / "csrc" / "libclpk.so"
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "F2FP.SATF", "MUFU.TANH", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("void ", "").replace("clpk::", "") for n in out]


def main() -> None:
    outdir, tag = Path(sys.argv[1]), sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["instr"] += 1
            for k in OPS:
                if op == k or op.startswith(k + ".") or (k == "F2FP.SATF" and op.startswith("F2FP.SATFINITE")):
                    counts[cur][k] += 1
    names = demangle(order)
    commit = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    lines = [f"# cuobjdump -sass {LIB.relative_to(ROOT)}: per-kernel counts of Blackwell-native SASS mnemonics (tools/sass_opcodes.py, sources of commit {commit})",
             f"# architectures in the fat binary: {archs};  UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG/UTMAPF = TMA tensor load/store/prefetch,",
             "# LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2/FADD2 = packed fp32x2 math, HMMA = legacy mma.sync (must be 0)",
             "", f"{'kernel':62s}" + "".join(f"{k:>10s}" for k in ["instr"] + OPS)]
    for raw, name in zip(order, names):
        c = counts[raw]
        lines.append(f"{name[:62]:62s}" + "".join(f"{c[k]:10d}" for k in ["instr"] + OPS))
    (outdir / f"sass_opcodes_{tag}.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[-14:]))


if __name__ == "__main__":
    main()
