"""Stages the UNMODIFIED reference package for the CPU baseline arm: /root/reference/src/clip_feature_codec ->
oracle/_ref/clip_feature_codec.zip (one importable archive built from the sources where they lie; git-ignored, travels to
the GPU box with the gpurun snapshot like a built .so).  Run in the build container only (`python oracle/stage_ref.py`,
also called by __graft_entry__.build()); the reference is pure Python, so "building" it is archiving it.  bench.py --impl reference and bench.py's cpu_baseline leg
import it from there (kind = "reference"); when oracle/_ref is absent they fall back to the oracle port (kind = "port").
Nothing under oracle/_ref is ever imported by the product package."""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference/src/clip_feature_codec")
DST = ROOT / "oracle" / "_ref" / "clip_feature_codec.zip"


def stage() -> bool:
    import zipfile

    if not SRC.exists():
        return DST.exists()
    DST.parent.mkdir(parents=True, exist_ok=True)
    tmp = DST.with_suffix(".zip.tmp")
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in sorted(SRC.rglob("*.py")):
            z.write(f, Path("clip_feature_codec") / f.relative_to(SRC))
    tmp.replace(DST)
    return True


if __name__ == "__main__":
    ok = stage()
    print(f"oracle/_ref staged: {ok} ({DST.stat().st_size if DST.exists() else 0} bytes)")
    sys.exit(0 if ok else 1)
