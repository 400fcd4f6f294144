"""Portable operator container: replaces the Julia `.jls` files of the reference's examples.

The reference builds its RBF (M)FSBP operators with an optimiser on the CPU and moves them between scripts with
`Serialization.serialize` (`examples/RBF_MFSBP.jl:51-52` writes `D_*.jls`, `examples/RBF_MFSBP_advection.jl:9` reads
it).  `.jls` is Julia-only, so an operator built on a Julia machine could not be uploaded through this package.  This
module defines a flat little-endian binary file that both sides can write with a few `write` calls
(`julia/SBPB200.jl: export_operator`) and read without any dependency:

    offset  type        content
    0       8 bytes     magic  b"SBPOP001"
    8       int64[6]    kind (1: 1D MatrixDerivativeOperator, 2: MultidimensionalMatrixDerivativeOperator), N, dim,
                        Nb (boundary entries; 0 for kind 1), accuracy_order, flags (bit 0: operators stored as CSR,
                        bit 1: explicit corner positions follow the operators)
    56      float64[2]  xmin, xmax (kind 1; zeros otherwise)
    72      float64[N*dim]   nodes, node-major (x1, y1, x2, y2, ...) = a Julia Vector{SVector{dim}} as it lies in memory
            float64[N]       weights (diagonal of P)
            int64[Nb]        boundary_indices, 1-BASED as in Julia
            float64[Nb*dim]  normals, entry-major
            float64[Nb]      weights_boundary
            dim operators (1 for kind 1), each either
              dense: float64[N*N] in COLUMN-major order (the memory of a Julia Matrix), or
              CSR:   int64 nnz, int64[N+1] rowptr (0-based offsets), int64[nnz] columns (1-based), float64[nnz] values
            flags bit 1 (operators whose corners are NOT the duplicates that `find_corners` detects, src/utils/corners.jl:6-18,
            e.g. upstream's TensorProductOperator ordering): int64 ncx, int64[ncx] corners_x, int64 ncy, int64[ncy]
            corners_y -- 1-based positions into boundary_indices

Everything on the path is Float64, so the container stores nothing else.
"""
from __future__ import annotations

import struct

import numpy as np

from ._lib import ArgumentError
from .operators import MatrixDerivativeOperator, MultidimensionalMatrixDerivativeOperator

MAGIC = b"SBPOP001"
KIND_1D, KIND_MD = 1, 2


def _csr(D):
    D = np.asarray(D)
    rows, cols = np.nonzero(D)
    rowptr = np.zeros(D.shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), cols.astype(np.int64) + 1, D[rows, cols].astype(np.float64)


def save_operator(path, D, storage: str = "auto") -> None:
    """Write `D` (a `MatrixDerivativeOperator` or a `MultidimensionalMatrixDerivativeOperator`) to `path`.
    storage: "dense", "csr" or "auto" (CSR when less than 25 % of the entries are non-zero)."""
    if isinstance(D, MultidimensionalMatrixDerivativeOperator):
        kind, dim, mats = KIND_MD, len(D.Ds), list(D.Ds)
        nodes = np.ascontiguousarray(D.grid, dtype=np.float64).reshape(D.N, dim)
        bidx = np.asarray(D.boundary_indices, dtype=np.int64) + 1
        normals = np.ascontiguousarray(D.normals, dtype=np.float64).reshape(len(bidx), dim)
        wb = np.asarray(D.weights_boundary, dtype=np.float64)
        xmin = xmax = 0.0
    elif isinstance(D, MatrixDerivativeOperator):
        kind, dim, mats = KIND_1D, 1, [D.D]
        nodes = np.asarray(D.grid, dtype=np.float64).reshape(D.N, 1)
        bidx, normals, wb = np.zeros(0, np.int64), np.zeros((0, 1)), np.zeros(0)
        xmin, xmax = D.xmin, D.xmax
    else:
        raise ArgumentError("save_operator: unsupported operator type " + type(D).__name__)
    N = D.N
    if storage not in ("auto", "dense", "csr"):
        raise ArgumentError("storage must be 'auto', 'dense' or 'csr'")
    sparse = storage == "csr" or (storage == "auto" and sum(np.count_nonzero(M) for M in mats) < 0.25 * len(mats) * N * N)
    with open(path, "wb") as f:
        f.write(MAGIC)
        corners = getattr(D, "corners", None) if kind == KIND_MD else None
        flags = (1 if sparse else 0) | (2 if corners is not None else 0)
        f.write(struct.pack("<6q", kind, N, dim, len(bidx), int(D.accuracy_order or 0), flags))
        f.write(struct.pack("<2d", float(xmin), float(xmax)))
        for a in (nodes, np.asarray(D.weights(), dtype=np.float64), bidx, normals, wb):
            f.write(np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
        for M in mats:
            M = np.asarray(M, dtype=np.float64)
            if sparse:
                rowptr, col, val = _csr(M)
                f.write(struct.pack("<q", len(val)))
                f.write(rowptr.astype("<i8").tobytes())
                f.write(col.astype("<i8").tobytes())
                f.write(val.astype("<f8").tobytes())
            else:
                f.write(np.asfortranarray(M).astype("<f8").tobytes(order="F"))
        if corners is not None:
            for c in corners:
                f.write(struct.pack("<q", len(c)))
                f.write((np.asarray(c, dtype=np.int64) + 1).astype("<i8").tobytes())


def load_operator(path, storage: str = "auto"):
    """Read an operator written by `save_operator` or by the Julia exporter.  Returns a `MatrixDerivativeOperator`
    (kind 1) or a `MultidimensionalMatrixDerivativeOperator` (kind 2; boundary indices converted to 0-based).
    `storage` is the device storage hint of the returned operator ("auto" | "dense" | "sparse")."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ArgumentError(f"{path}: not an SBPOP001 operator container")
    kind, N, dim, Nb, order, flags = struct.unpack_from("<6q", buf, 8)
    sparse, has_corners = flags & 1, (flags >> 1) & 1
    xmin, xmax = struct.unpack_from("<2d", buf, 56)
    if kind not in (KIND_1D, KIND_MD) or N <= 0 or dim <= 0 or Nb < 0 or flags not in (0, 1, 2, 3):
        raise ArgumentError(f"{path}: corrupt header")
    off = 72

    def take(dtype, count):
        nonlocal off
        nbytes = np.dtype(dtype).itemsize * count
        if off + nbytes > len(buf):
            raise ArgumentError(f"{path}: truncated file")
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
        off += nbytes
        return a

    nodes = take("<f8", N * dim).reshape(N, dim)
    w = take("<f8", N)
    bidx = take("<i8", Nb)
    normals = take("<f8", Nb * dim).reshape(Nb, dim)
    wb = take("<f8", Nb)
    mats = []
    for _ in range(dim if kind == KIND_MD else 1):
        if sparse:
            nnz = int(take("<i8", 1)[0])
            rowptr, col, val = take("<i8", N + 1), take("<i8", nnz), take("<f8", nnz)
            if rowptr[0] != 0 or rowptr[-1] != nnz or np.any(np.diff(rowptr) < 0) or (nnz and (col.min() < 1 or col.max() > N)):
                raise ArgumentError(f"{path}: corrupt CSR operator")
            M = np.zeros((N, N))
            rows = np.repeat(np.arange(N), np.diff(rowptr))
            np.add.at(M, (rows, col - 1), val)
        else:
            M = take("<f8", N * N).reshape(N, N, order="F")
        mats.append(np.ascontiguousarray(M))
    corners = None
    if has_corners:
        corners = []
        for _ in range(2):
            n = int(take("<i8", 1)[0])
            if n < 0 or n > Nb:
                raise ArgumentError(f"{path}: corrupt corner list")
            corners.append([int(c) - 1 for c in take("<i8", n)])
        corners = tuple(corners)
    if off != len(buf):
        raise ArgumentError(f"{path}: {len(buf) - off} trailing bytes")
    if kind == KIND_1D:
        x = nodes[:, 0]
        return MatrixDerivativeOperator(xmin, xmax, x, w, mats[0], order, "container:" + str(path), storage=storage)
    if Nb and (bidx.min() < 1 or bidx.max() > N):
        raise ArgumentError(f"{path}: boundary index out of range")
    return MultidimensionalMatrixDerivativeOperator(nodes, bidx - 1, normals, w, wb, mats, order,
                                                    "container:" + str(path), storage=storage, corners=corners)
