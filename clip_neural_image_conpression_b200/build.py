"""Builds csrc/libclpk.so (hand-written sm_100a CUDA + the C ABI of include/clpk.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the resulting .so travels to the B200 box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libclpk.so"
SOURCES = ["elementwise.cu", "groupnorm.cu", "conv_in.cu", "conv_igemm.cu", "head_conv.cu", "plan.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libclpk.so cannot be built (set NVCC=/path/to/nvcc)")


def _deps() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "clpk.h"]


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps() if p.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compiles the sources when libclpk.so is missing or older than them.  Safe under torchrun: an inter-process file
    lock serialises the ranks (the first one builds, the others find a fresh library), objects go to a per-PID
    directory and the library is replaced atomically."""
    if not force and not is_stale():
        return LIB
    import fcntl
    import tempfile

    with open(CSRC / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():   # another rank built it while this one waited for the lock
                return LIB
            nvcc = find_nvcc()
            (CSRC / "build").mkdir(exist_ok=True)
            with tempfile.TemporaryDirectory(prefix=f"obj{os.getpid()}_", dir=CSRC / "build") as objdir:
                return _compile_and_link(nvcc, Path(objdir), verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _compile_and_link(nvcc: str, objdir: Path, verbose: bool) -> Path:
    def compile_one(src: str) -> Path:
        obj = objdir / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("CLPK_NVCC_EXTRA", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
