""".clp bitstream container (format of the reference's PKG/io/bitstream.py:14-34, unchanged).

Layout: bytes 0-3 b"CLPF" | bytes 4-7 little-endian u32 = length of the zstd frame | one zstd frame holding the D raw
uint8 codes.  (The reference declares VERSION = 1 but never writes it, and ignores `dim`; both are preserved.)

The reference delegates to the un-vendored PyPI package `zstandard>=0.22.0` (pyproject.toml:21).  That wheel is not
part of this image, so the frame codec here is the system libzstd (>= 1.4) bound through ctypes — zstd frames are a
stable, versioned format, so frames written by either side decode to identical bytes on the other.
Entropy decoding of ~0.5 KB frames stays on the host by design (north-star: "zstd-decoded ... on the host"); what
follows it (dequantise, L2 renorm) runs on the device in one batched kernel.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import struct
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Iterable, List

import numpy as np

MAGIC = b"CLPF"
VERSION = 1
_LEVEL = 22  # write_bitstream uses ZstdCompressor(level=22) (bitstream.py:19)

_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        lib = C.CDLL(name)
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_compress.restype = C.c_size_t
        lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.ZSTD_decompress.restype = C.c_size_t
        lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        lib.ZSTD_isError.restype = C.c_uint
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        lib.ZSTD_getErrorName.restype = C.c_char_p
        lib.ZSTD_getErrorName.argtypes = [C.c_size_t]
        _zstd = lib
    return _zstd


def zstd_compress(data: bytes, level: int = _LEVEL) -> bytes:
    lib = _libzstd()
    bound = lib.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(bound)
    n = lib.ZSTD_compress(dst, bound, data, len(data), level)
    if lib.ZSTD_isError(n):
        raise ValueError(f"zstd compress failed: {lib.ZSTD_getErrorName(n).decode()}")
    return dst.raw[:n]


def zstd_decompress(frame: bytes) -> bytes:
    lib = _libzstd()
    size = lib.ZSTD_getFrameContentSize(frame, len(frame))
    if size in (2 ** 64 - 1, 2 ** 64 - 2):  # ZSTD_CONTENTSIZE_UNKNOWN / _ERROR
        raise ValueError("zstd frame without a content size (not produced by write_bitstream)")
    dst = C.create_string_buffer(max(int(size), 1))
    n = lib.ZSTD_decompress(dst, int(size), frame, len(frame))
    if lib.ZSTD_isError(n):
        raise ValueError(f"zstd decompress failed: {lib.ZSTD_getErrorName(n).decode()}")
    return dst.raw[:n]


def write_bitstream(q_bytes: bytes, dim: int, out_path: Path) -> None:
    frame = zstd_compress(bytes(q_bytes))
    with open(out_path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(frame)) + frame)


def parse_bitstream(blob: bytes) -> np.ndarray:
    assert blob[:4] == MAGIC, "Bad magic"
    (ln,) = struct.unpack("<I", blob[4:8])
    return np.frombuffer(zstd_decompress(blob[8:8 + ln]), dtype=np.uint8)


def read_bitstream(in_path: Path) -> np.ndarray:
    with open(in_path, "rb") as f:
        return parse_bitstream(f.read())


def read_bitstreams(paths: Iterable[Path], threads: int = 16) -> np.ndarray:
    """Batched reader: N .clp files -> uint8 [N, D] (one pinned-host-friendly array for a single H2D copy).
    File IO and zstd both release the GIL, so a small thread pool overlaps them."""
    paths = list(paths)
    if not paths:
        return np.zeros((0, 0), dtype=np.uint8)
    with ThreadPoolExecutor(max_workers=max(1, min(threads, len(paths)))) as ex:
        rows: List[np.ndarray] = list(ex.map(read_bitstream, paths))
    d = rows[0].shape[0]
    for p, r in zip(paths, rows):
        if r.shape[0] != d:
            raise ValueError(f"{p}: {r.shape[0]} codes, expected {d}")
    return np.stack(rows)


def write_bitstreams(codes: np.ndarray, paths: Iterable[Path], threads: int = 16) -> None:
    """Batched writer: uint8 [N, D] -> N .clp files, byte-identical to N write_bitstream calls (the per-vector loop of
    PKG/cli/encode_images.py:79-83).  zstd and file IO release the GIL, so a small thread pool overlaps them."""
    paths = list(paths)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    if codes.ndim != 2 or codes.shape[0] != len(paths):
        raise ValueError(f"codes {codes.shape} do not match {len(paths)} paths")
    if not paths:
        return
    dim = codes.shape[1]
    with ThreadPoolExecutor(max_workers=max(1, min(threads, len(paths)))) as ex:
        list(ex.map(lambda ip: write_bitstream(codes[ip[0]].tobytes(), dim, ip[1]), enumerate(paths)))
