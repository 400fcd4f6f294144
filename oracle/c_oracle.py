"""ctypes wrapper of oracle/oracle_c.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (see the C file's header).

The shared object is compiled with the host's gcc (-O3 -march=native -fopenmp) into oracle/_build/, keyed by the
host CPU's flag set so that a library built in the CPU-only container is never run on a different CPU on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _cpu_key() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()[:12]
    except OSError:
        pass
    return "generic"


def build() -> str:
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, f"liboracle_c_{_cpu_key()}.so")
    src = os.path.join(_HERE, "oracle_c.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-o", so, src]
        subprocess.run(cmd, check=True, capture_output=True)
    return so


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        dp, ip, L, I, D = C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_double
        lib.oracle_max_threads.restype = I
        lib.oracle_apply_dense.argtypes = [dp, I, dp, dp, L, D, D, I]
        lib.oracle_apply_dense_blas.argtypes = [dp, dp, I, dp, dp, L, D, D, I]
        lib.oracle_apply_dense_blas.restype = None
        lib.oracle_apply_subcell.argtypes = [dp, dp, dp, I, I, ip, dp, dp, dp, L, D, D, I]
        lib.oracle_apply_csr.argtypes = [ip, ip, dp, I, dp, dp, L, D, D, I]
        lib.oracle_rhs_advection2d.argtypes = [dp, ip, ip, dp, I, dp, ip, dp, L, dp, dp, L, I, I]
        lib.oracle_energy.argtypes = [dp, I, I, dp, dp, L, I]
        lib.oracle_solve_ssprk53_1d.argtypes = [dp, I, dp, D, D, dp, L, dp, dp, dp]
        lib.oracle_solve_ssprk53_1d.restype = None
        for f in ("oracle_apply_dense", "oracle_apply_subcell", "oracle_apply_csr", "oracle_rhs_advection2d",
                  "oracle_energy"):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(load().oracle_max_threads())


def apply_dense(Dmat, U, alpha=1.0, beta=0.0, Y=None, nthreads=1):
    """per-instance dgemv loop; Dmat is the N x N apply matrix, U is (B, N)."""
    D = np.asfortranarray(Dmat, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    Y = np.empty_like(U) if Y is None else Y
    load().oracle_apply_dense(_p(D), N, _p(U), _p(Y), B, alpha, beta, nthreads)
    return Y


def openblas_dgemv_address() -> int:
    """Address of the Fortran-interface `dgemv_` of the OpenBLAS bundled with scipy (scipy.linalg.cython_blas exports the
    BLAS entry points as PyCapsules) -- the routine Julia's `mul!(dest, D::Matrix, u, alpha, beta)` ends up in."""
    from scipy.linalg import cython_blas

    cap = cython_blas.__pyx_capi__["dgemv"]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


def apply_dense_blas(Dmat, U, alpha=1.0, beta=0.0, Y=None, nthreads=1):
    """per-instance OpenBLAS dgemv loop (the reference's own BLAS call per vector), instances spread over OpenMP threads."""
    D = np.asfortranarray(Dmat, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    Y = np.empty_like(U) if Y is None else Y
    load().oracle_apply_dense_blas(C.c_void_p(openblas_dgemv_address()), _p(D), N, _p(U), _p(Y), B, alpha, beta, nthreads)
    return Y


def apply_subcell(Q_left, Q_right, w, U, op_index=None, alpha=1.0, beta=0.0, energy=False, nthreads=1):
    QL = np.ascontiguousarray(np.stack([np.asfortranarray(q).ravel(order="F") for q in Q_left]))
    QR = np.ascontiguousarray(np.stack([np.asfortranarray(q).ravel(order="F") for q in Q_right]))
    W = np.ascontiguousarray(np.stack(w), dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    G = W.shape[0]
    Y = np.empty_like(U)
    E = np.empty(B) if energy else None
    idx = None if op_index is None else np.ascontiguousarray(op_index, dtype=np.int32)
    load().oracle_apply_subcell(_p(QL), _p(QR), _p(W), N, G, _p(idx), _p(U), _p(Y), _p(E), B, alpha, beta, nthreads)
    return (Y, E) if energy else Y


def dense_to_csr(Dmat):
    Dm = np.asarray(Dmat)
    N = Dm.shape[0]
    rows, cols = np.nonzero(Dm)
    rowptr = np.zeros(N + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, cols.astype(np.int32), np.ascontiguousarray(Dm[rows, cols], dtype=np.float64)


def apply_csr(rowptr, col, val, U, alpha=1.0, beta=0.0, Y=None, nthreads=1):
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    Y = np.empty_like(U) if Y is None else Y
    load().oracle_apply_csr(_p(rowptr), _p(col), _p(val), N, _p(U), _p(Y), B, alpha, beta, nthreads)
    return Y


def rhs_advection2d(A, bdiag, mode, g, U, faithful=False, nthreads=1):
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    dU = np.empty_like(U)
    g = np.ascontiguousarray(g, dtype=np.float64)
    stride = 0 if g.ndim == 1 else N
    b = np.ascontiguousarray(bdiag, dtype=np.float64)
    m = np.ascontiguousarray(mode, dtype=np.int32)
    if faithful:
        Af = np.asfortranarray(A, dtype=np.float64)
        load().oracle_rhs_advection2d(_p(Af), None, None, None, N, _p(b), _p(m), _p(g), stride, _p(U), _p(dU), B, 1,
                                      nthreads)
    else:
        rp, ci, va = dense_to_csr(A)
        load().oracle_rhs_advection2d(None, _p(rp), _p(ci), _p(va), N, _p(b), _p(m), _p(g), stride, _p(U), _p(dU), B, 0,
                                      nthreads)
    return dU


def energy(w, U, nthreads=1):
    W = np.ascontiguousarray(np.atleast_2d(w), dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    B, N = U.shape
    E = np.empty(B)
    load().oracle_energy(_p(W), N, W.shape[0], _p(U), _p(E), B, nthreads)
    return E


def solve_ssprk53_1d(Dmat, a, w0, wN, u0, hsteps, bc_table):
    """Whole fixed-step SSPRK53 solve of the upstream 1D semidiscretization for one instance (the CPU path of
    examples/RBF_FSBP_advection.jl: dense dgemv per stage); bc_table: (5 * nsteps, 2) [left, right] at the stage times."""
    D = np.asfortranarray(Dmat, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    u = np.array(u0, dtype=np.float64, copy=True)
    hs = np.ascontiguousarray(hsteps, dtype=np.float64)
    bc = np.ascontiguousarray(bc_table, dtype=np.float64)
    N = len(u)
    work = np.empty(5 * N)
    load().oracle_solve_ssprk53_1d(_p(D), N, _p(a), float(w0), float(wN), _p(u), len(hs), _p(hs), _p(bc), _p(work))
    return u
