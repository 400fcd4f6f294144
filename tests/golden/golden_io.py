"""Flat binary container for golden vectors that both numpy and a plain Julia script (no NPZ.jl) can read and write.

Layout `SBPGOLD1` (little endian): 8 magic bytes, Int64 count, then per array:
Int64 name length, name (UTF-8), Int64 ndim, Int64 dims[ndim] (numpy C-order shape), Float64 data in C order
(a Julia `Matrix(N, B)` holding one instance per column is written column by column, i.e. as the numpy array (B, N)).
"""
import struct

import numpy as np

MAGIC = b"SBPGOLD1"


def write_bin(path, arrays: dict):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<q", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a, dtype=np.float64)
            nb = name.encode()
            f.write(struct.pack("<q", len(nb)))
            f.write(nb)
            f.write(struct.pack("<q", a.ndim))
            f.write(struct.pack(f"<{a.ndim}q", *a.shape))
            f.write(a.tobytes())


def read_bin(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not an SBPGOLD1 file")
        (count,) = struct.unpack("<q", f.read(8))
        for _ in range(count):
            (ln,) = struct.unpack("<q", f.read(8))
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<q", f.read(8))
            shape = struct.unpack(f"<{nd}q", f.read(8 * nd)) if nd else ()
            n = int(np.prod(shape)) if nd else 1
            out[name] = np.frombuffer(f.read(8 * n), dtype="<f8").reshape(shape).copy()
    return out
