"""ctypes binding of libsbp_b200.so (C ABI declared in include/sbp_b200.h).

There is no CPU fallback: if the shared library is missing `load()` raises, and if no CUDA device is
present `Context()` raises `CudaError`.  Error codes map onto the reference's exception types:
SBP_EDIM -> DimensionMismatch, SBP_EINVAL -> ArgumentError (src/polynomialbases_operators.jl:143-146,
src/subcell_operators.jl:195-196).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBP_B200_LIB") or os.path.join(_HERE, "libsbp_b200.so")  # override: experiment builds
CSRC = os.path.join(_HERE, "csrc")

SBP_OK, SBP_EINVAL, SBP_EDIM, SBP_ECUDA, SBP_ENOMEM, SBP_ENCCL, SBP_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
OP_DENSE, OP_SPARSE, OP_TABLE, OP_ADVECTION2D, OP_ADVECTION1D = 1, 2, 3, 4, 5
DENSE_AUTO, DENSE_FORCE_GEMM, DENSE_NO_SYMMETRY, DENSE_SYMMETRIZE, DENSE_NO_SKIP, DENSE_SYM_BITWISE = 0, 1, 2, 4, 8, 16
KERNEL_NAMES = {1: "dmma_smem", 2: "dmma_smem_sym", 3: "dgemm_tiled", 4: "csr_smem", 5: "table_generic", 6: "bsr_dmma"}
NCCL_ID_BYTES = 128


class SbpError(RuntimeError):
    pass


class ArgumentError(ValueError, SbpError):
    """Julia `ArgumentError` (ArgCheck.@argcheck failures in the reference)."""


class DimensionMismatch(ValueError, SbpError):
    """Julia `DimensionMismatch`."""


class CudaError(SbpError):
    pass


class NcclError(SbpError):
    pass


class UnsupportedError(SbpError):
    pass


_ERR = {SBP_EINVAL: ArgumentError, SBP_EDIM: DimensionMismatch, SBP_ECUDA: CudaError, SBP_ENOMEM: MemoryError,
        SBP_ENCCL: NcclError, SBP_EUNSUPPORTED: UnsupportedError}

_lib = None

c_i32, c_i64, c_dbl, c_void_p, c_size_t = C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_size_t
P = C.POINTER

# name -> (restype, argtypes); mirrors include/sbp_b200.h one to one
SIGNATURES = {
    "sbp_version": (c_i32, []),
    "sbp_last_error": (C.c_char_p, []),
    "sbp_device_count": (c_i32, [P(c_i32)]),
    "sbp_ctx_create": (c_i32, [c_i32, P(c_void_p)]),
    "sbp_ctx_set_stream": (c_i32, [c_void_p, c_void_p]),
    "sbp_ctx_get_stream": (c_i32, [c_void_p, P(c_void_p)]),
    "sbp_ctx_sync": (c_i32, [c_void_p]),
    "sbp_ctx_destroy": (c_i32, [c_void_p]),
    "sbp_ctx_device_info": (c_i32, [c_void_p, P(c_i32), P(c_i64), P(c_i64), P(c_i64)]),
    "sbp_ctx_launch_count": (c_i64, [c_void_p]),
    "sbp_ctx_set_option": (c_i32, [c_void_p, C.c_char_p, c_i32]),
    "sbp_ctx_get_option": (c_i32, [c_void_p, C.c_char_p, P(c_i32)]),
    "sbp_malloc": (c_i32, [c_void_p, c_size_t, P(c_void_p)]),
    "sbp_free": (c_i32, [c_void_p, c_void_p]),
    "sbp_ctx_pci_bus_id": (c_i32, [c_void_p, C.c_char_p, c_i32]),
    "sbp_host_alloc": (c_i32, [c_size_t, P(c_void_p)]),
    "sbp_host_free": (c_i32, [c_void_p]),
    "sbp_memcpy_h2d": (c_i32, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "sbp_memcpy_d2h": (c_i32, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "sbp_memset": (c_i32, [c_void_p, c_void_p, c_i32, c_size_t]),
    "sbp_op_dense_create": (c_i32, [c_void_p, c_i32, c_void_p, c_i32, c_void_p, c_dbl, c_i32, P(c_void_p)]),
    "sbp_op_sparse_create": (c_i32, [c_void_p, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_dbl, P(c_void_p)]),
    "sbp_op_sparse_from_dense": (c_i32, [c_void_p, c_i32, c_void_p, c_i32, c_void_p, c_dbl, P(c_void_p)]),
    "sbp_op_table_create": (c_i32, [c_void_p, c_i32, c_i32, c_void_p, c_void_p, P(c_void_p)]),
    "sbp_op_advection2d_create": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p,
                                          c_void_p, P(c_void_p)]),
    "sbp_op_advection1d_create": (c_i32, [c_void_p, c_void_p, c_void_p, P(c_void_p)]),
    "sbp_op_destroy": (c_i32, [c_void_p]),
    "sbp_op_info": (c_i32, [c_void_p, P(c_i32), P(c_i32), P(c_i32), P(c_i64), P(c_i32)]),
    "sbp_apply": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_dbl]),
    "sbp_apply_table": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_dbl, c_dbl, c_void_p,
                                c_void_p]),
    "sbp_integrate": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_void_p, c_void_p]),
    "sbp_scale_by_mass_matrix": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_i32]),
    "sbp_rhs_advection2d": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "sbp_rhs_advection1d": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "sbp_rhs_advection2d_stage": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_dbl, c_dbl,
                                          c_void_p, c_dbl]),
    "sbp_rhs_advection1d_stage": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_dbl, c_dbl,
                                          c_void_p, c_dbl]),
    "sbp_axpbypcz": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_dbl, c_dbl]),
    "sbp_lincomb": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_dbl, c_dbl, c_dbl]),
    "sbp_comm_unique_id": (c_i32, [c_void_p]),
    "sbp_comm_init": (c_i32, [c_void_p, c_i32, c_i32, c_void_p]),
    "sbp_allreduce_sum": (c_i32, [c_void_p, c_void_p, c_i32]),
    "sbp_comm_destroy": (c_i32, [c_void_p]),
    "sbp_apply_host": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_dbl, c_i64]),
    "sbp_rhs_advection2d_host": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64]),
    "sbp_solve_ssprk53": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_void_p]),
    "sbp_solve_ssprk53_host": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_void_p]),
    "sbp_pcie_probe": (c_i32, [c_void_p, c_size_t, P(c_dbl), P(c_dbl), P(c_dbl)]),
    "sbp_solve_3sp_fsal": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_dbl, c_dbl, c_dbl, c_i32, c_dbl, c_dbl,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_i64, c_void_p]),
    "sbp_error_norm": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_dbl, c_dbl, P(c_dbl)]),
    "sbp_objective_create": (c_i32, [c_void_p, c_i32, c_i32, c_i32, c_i32, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_dbl, P(c_void_p)]),
    "sbp_objective_eval": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "sbp_objective_destroy": (c_i32, [c_void_p]),
}


class Tableau3Sp(C.Structure):
    """`sbp_tableau_3sp` of include/sbp_b200.h."""
    _fields_ = [("stages", c_i32), ("order_for_controller", c_i32), ("gamma1", c_void_p), ("gamma2", c_void_p),
                ("gamma3", c_void_p), ("delta", c_void_p), ("c", c_void_p), ("beta", c_void_p), ("bhat", c_void_p),
                ("bhat_fsal", c_dbl)]


BC_FN = C.CFUNCTYPE(None, c_dbl, P(c_dbl), c_void_p)
STOP_FN = C.CFUNCTYPE(None, c_dbl, c_void_p)


def build(verbose: bool = False) -> str:
    """Compile libsbp_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libsbp_b200.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    if verbose:
        print(proc.stdout[-2000:])
    return LIB_PATH


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` or `make -C {CSRC}`. "
            "This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc == SBP_OK:
        return
    msg = load().sbp_last_error().decode("utf-8", "replace")
    raise _ERR.get(rc, SbpError)(f"[sbp {rc}] {msg}")
