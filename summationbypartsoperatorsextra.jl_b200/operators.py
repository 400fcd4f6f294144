"""Host-side mirror of the reference's operator interface for the hot path.

Names and argument meaning follow the reference (Julia `f!` is spelled `f_` here):

    mul_(dest, D, u, alpha=True, beta=False)        LinearAlgebra.mul!          (src/polynomialbases_operators.jl:139-149,
                                                                                src/subcell_operators.jl:270-280, upstream)
    D * u, D[dim] * u, Matrix(D), size(D), grid(D)  upstream generic fallbacks
    integrate(func, u, D), mass_matrix(D), mass_matrix_boundary(D[, dim]),
    scale_by_mass_matrix_(u, D, factor), scale_by_inverse_mass_matrix_(u, D, factor), get_weight(D, i), ...

Operator CONSTRUCTION stays on the CPU (numpy) exactly as in the reference (north_star: "uploaded
once"); every operator lazily uploads its arrays through the C ABI the first time it is applied to a
`DeviceBatch` of some `Context`.  All batch arithmetic runs in libsbp_b200.so -- nothing here computes
`D*u` on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, check
from .device import Context, DeviceBatch, RawBuffer

# ------------------------------------------------------------------------------------------------
# functions accepted by integrate(func, u, D) on the device
# ------------------------------------------------------------------------------------------------


def identity(x):
    return x


def abs2(x):
    return x * x


class power:
    """`x -> x^k` for an integer 0 <= k < 64 as an `integrate` integrand (evaluated on the device by binary powering)."""

    def __init__(self, k: int):
        if int(k) != k or not 0 <= int(k) < 64:
            raise ArgumentError("power(k): integer 0 <= k < 64")
        self.k = int(k)

    def __call__(self, x):
        return x ** self.k


_INTEGRATE_MODES = {identity: 0, abs2: 1, np.square: 1, None: 0, abs: 3, np.abs: 3, np.absolute: 3}


def _ptr(x):
    return C.c_void_p(x) if x else None


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _DeviceOp:
    """Handle of an uploaded `sbp_op` (immutable on the device)."""

    def __init__(self, ctx: Context, handle: C.c_void_p, keep=()):
        self.ctx, self.handle, self._keep = ctx, handle, keep

    def info(self):
        kind, N, G, nnz, var = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32()
        check(self.ctx._lib.sbp_op_info(self.handle, C.byref(kind), C.byref(N), C.byref(G), C.byref(nnz),
                                        C.byref(var)))
        return {"kind": kind.value, "N": N.value, "G": G.value, "nnz": nnz.value,
                "kernel": _lib.KERNEL_NAMES.get(var.value, str(var.value))}

    def destroy(self):
        if self.handle:
            self.ctx._lib.sbp_op_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            if self.ctx._h:
                self.destroy()
        except Exception:
            pass


def upload_dense(ctx: Context, Dmat, weights=None, scale=1.0, flags=_lib.DENSE_AUTO) -> _DeviceOp:
    D = np.asfortranarray(Dmat, dtype=np.float64)  # Julia Matrix: column-major
    if D.ndim != 2 or D.shape[0] != D.shape[1]:
        raise DimensionMismatch("operator matrix must be square")
    N = D.shape[0]
    w = None if weights is None else _as_f64(weights)
    if w is not None and w.shape != (N,):
        raise DimensionMismatch("weights must have length N")
    h = C.c_void_p()
    check(ctx._lib.sbp_op_dense_create(ctx.handle, N, D.ctypes.data_as(C.c_void_p), N,
                                       w.ctypes.data_as(C.c_void_p) if w is not None else None, float(scale),
                                       int(flags), C.byref(h)))
    return _DeviceOp(ctx, h)


def upload_sparse(ctx: Context, Dmat, weights=None, scale=1.0) -> _DeviceOp:
    """Upload only the structural non-zeros of a dense-stored operator (CSR -> SELL-32 on the device)."""
    D = np.asfortranarray(Dmat, dtype=np.float64)
    if D.ndim != 2 or D.shape[0] != D.shape[1]:
        raise DimensionMismatch("operator matrix must be square")
    N = D.shape[0]
    w = None if weights is None else _as_f64(weights)
    h = C.c_void_p()
    check(ctx._lib.sbp_op_sparse_from_dense(ctx.handle, N, D.ctypes.data_as(C.c_void_p), N,
                                            w.ctypes.data_as(C.c_void_p) if w is not None else None, float(scale),
                                            C.byref(h)))
    return _DeviceOp(ctx, h)


def upload_csr(ctx: Context, N, rowptr, col, val, weights=None, scale=1.0) -> _DeviceOp:
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(col, dtype=np.int32)
    va = _as_f64(val)
    if rp.shape != (N + 1,) or ci.shape != va.shape:
        raise DimensionMismatch("bad CSR arrays")
    w = None if weights is None else _as_f64(weights)
    h = C.c_void_p()
    check(ctx._lib.sbp_op_sparse_create(ctx.handle, int(N), rp.ctypes.data_as(C.c_void_p),
                                        ci.ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p),
                                        w.ctypes.data_as(C.c_void_p) if w is not None else None, float(scale),
                                        C.byref(h)))
    return _DeviceOp(ctx, h)


def upload_table(ctx: Context, Dmats, weights) -> _DeviceOp:
    Ds = np.stack([np.asfortranarray(D, dtype=np.float64).ravel(order="F") for D in Dmats])  # G x N*N (col-major)
    W = _as_f64(np.stack([_as_f64(w) for w in weights]))
    G = Ds.shape[0]
    N = int(round(np.sqrt(Ds.shape[1])))
    if N * N != Ds.shape[1] or W.shape != (G, N):
        raise DimensionMismatch("table needs G matrices N x N and G weight vectors of length N")
    Ds = np.ascontiguousarray(Ds)
    h = C.c_void_p()
    check(ctx._lib.sbp_op_table_create(ctx.handle, N, G, Ds.ctypes.data_as(C.c_void_p), W.ctypes.data_as(C.c_void_p),
                                       C.byref(h)))
    return _DeviceOp(ctx, h)


# ------------------------------------------------------------------------------------------------
# generic operator behaviour (the reference's AbstractNonperiodicDerivativeOperator interface)
# ------------------------------------------------------------------------------------------------


class DerivativeOperator:
    """Common host-side behaviour.  Subclasses provide `_dense()` (N x N apply matrix WITHOUT `_scale()`),
    `_scale()` (constant folded into every apply), `weights()` and `grid`."""

    storage = "auto"  # "auto" | "dense" | "sparse"
    # sbp_op_dense_create flags: 0 = use the even/odd kernel when D is centro-antisymmetric up to 256 eps (the
    # antisymmetric part (D - JDJ)/2 is applied); _lib.DENSE_NO_SYMMETRY never; _lib.DENSE_SYM_BITWISE only when bitwise
    dense_flags = _lib.DENSE_AUTO

    def __init__(self):
        self._dev: dict[int, _DeviceOp] = {}

    # -- to be provided -----------------------------------------------------------------------------
    def _dense(self) -> np.ndarray:  # pragma: no cover
        raise NotImplementedError

    def _scale(self) -> float:
        return 1.0

    def weights(self) -> np.ndarray:  # pragma: no cover
        raise NotImplementedError

    # -- generic interface ----------------------------------------------------------------------------
    def size(self, dim: int | None = None):
        N = len(self.grid)
        return (N, N) if dim is None else N

    @property
    def N(self) -> int:
        return len(self.grid)

    def eltype(self):
        return np.float64

    def Matrix(self) -> np.ndarray:
        """`Matrix(D)`: the N x N matrix that `mul_` applies (host copy, for inspection/tests)."""
        return self._scale() * self._dense()

    def mass_matrix(self) -> np.ndarray:
        return np.diag(self.weights())

    def _want_sparse(self, D: np.ndarray) -> bool:
        if self.storage == "dense":
            return False
        if self.storage == "sparse":
            return True
        N = D.shape[0]
        if N <= _lib_small_n():
            # smem-resident DMMA path: HBM/FP64-rate bound for full matrices.  Banded operators (FD-SBP, the banded RBF
            # FSBP operators of examples/RBF_FSBP_advection.jl) from N ~ 48 are faster on the block-sparse kernel
            # (measured, B = 2^19: N = 40: 341 dense vs 293 sparse GDOF/s; 65: 179 vs 314; 101: 168 vs 294; 104: 321 vs
            # 391; 128: 328 vs 337)
            return N >= 48 and np.count_nonzero(D) < 0.25 * N * N
        return np.count_nonzero(D) < 0.15 * N * N

    def device_op(self, ctx: Context) -> _DeviceOp:
        key = id(ctx)
        op = self._dev.get(key)
        if op is None or op.handle is None:
            D = self._dense()
            if self._want_sparse(D):
                op = upload_sparse(ctx, D, self.weights(), self._scale())
            else:
                op = upload_dense(ctx, D, self.weights(), self._scale(), self.dense_flags)
            self._dev[key] = op
        return op

    def __mul__(self, u: DeviceBatch) -> DeviceBatch:
        """`D * u`: dest = similar(u); mul!(dest, D, u)."""
        dest = u.similar()
        return mul_(dest, self, u)

    __matmul__ = __mul__


def _lib_small_n() -> int:
    return 128


def _check_batch(D, dest: DeviceBatch | None, u: DeviceBatch):
    N = D.size(1)
    if u.N != N or (dest is not None and dest.N != N):
        # @argcheck N == length(u); @argcheck N == length(dest)  -> ArgumentError
        raise ArgumentError(f"N == length(u) must hold (operator has {N} nodes, vectors have {u.N})")
    if dest is not None and dest.B != u.B:
        raise DimensionMismatch("dest and u hold different numbers of instances")


def mul_(dest: DeviceBatch, D: DerivativeOperator, u: DeviceBatch, alpha=True, beta=False) -> DeviceBatch:
    """`mul!(dest, D, u, alpha, beta)` for every instance of the batch: dest = alpha*D*u + beta*dest."""
    _check_batch(D, dest, u)
    ctx = u.ctx
    op = D.device_op(ctx)
    check(ctx._lib.sbp_apply(ctx.handle, op.handle, _ptr(dest.ptr), _ptr(u.ptr), u.B, float(alpha), float(beta)))
    return dest


def grid(D):
    return D.grid


def mass_matrix(D):
    return D.mass_matrix()


def mass_matrix_boundary(D, dim: int | None = None):
    return D.mass_matrix_boundary() if dim is None else D.mass_matrix_boundary(dim)


def Matrix(D):
    return D.Matrix()


def integrate(func, u: DeviceBatch, D: DerivativeOperator, v: DeviceBatch | None = None, total: bool = False):
    """`integrate(func, u, D)` = sum_i func(u_i) w_i per instance (host array of B values; a float for B == 1).
    func is `identity`, `abs2`/`np.square`, `abs`, `power(k)`, or None with `v` given (weighted inner product
    sum_i w_i u_i v_i).
    `total=True` additionally returns the sum over the batch computed on the device."""
    if u.N != D.size(1):
        raise DimensionMismatch("sizes of input vector and operator do not match")
    ctx = u.ctx
    if v is not None:
        mode = 2
        if v.shape != u.shape:
            raise DimensionMismatch("u and v differ in shape")
    else:
        if isinstance(func, power):
            mode = 16 + func.k
        elif func in _INTEGRATE_MODES:
            mode = _INTEGRATE_MODES[func]
        else:
            raise ArgumentError("device integrate supports identity, abs2, abs, power(k) and the weighted inner product "
                                "(the reference accepts any func, src/subcell_operators.jl:92-94)")
    op = D.device_op(ctx)
    out = RawBuffer(ctx, max(u.B, 1) * 8 + 8)
    check(ctx._lib.sbp_integrate(ctx.handle, op.handle, _ptr(u.ptr), _ptr(v.ptr) if v is not None else None, u.B,
                                 mode, _ptr(out.ptr), _ptr(out.ptr + max(u.B, 1) * 8)))
    vals = out.to_host(np.float64, max(u.B, 1) + 1)
    out.free()
    per = vals[: u.B]
    res = float(per[0]) if u.B == 1 else per.copy()
    return (res, float(vals[-1])) if total else res


def scale_by_mass_matrix_(u: DeviceBatch, D: DerivativeOperator, factor=True) -> DeviceBatch:
    if u.N != D.size(2):
        raise DimensionMismatch("sizes of input vector and operator do not match")
    ctx = u.ctx
    check(ctx._lib.sbp_scale_by_mass_matrix(ctx.handle, D.device_op(ctx).handle, _ptr(u.ptr), u.B, float(factor), 0))
    return u


def scale_by_inverse_mass_matrix_(u: DeviceBatch, D: DerivativeOperator, factor=True) -> DeviceBatch:
    if u.N != D.size(2):
        raise DimensionMismatch("sizes of input vector and operator do not match")
    ctx = u.ctx
    check(ctx._lib.sbp_scale_by_mass_matrix(ctx.handle, D.device_op(ctx).handle, _ptr(u.ptr), u.B, float(factor), 1))
    return u


# ------------------------------------------------------------------------------------------------
# PolynomialBases.jl subset used by the reference (nodes / weights / D of nodal bases on [-1, 1])
# ------------------------------------------------------------------------------------------------


class NodalBasis:
    """`basis.nodes`, `basis.weights`, `basis.D`, `basis.baryweights` of a PolynomialBases.NodalBasis{Line}."""

    name = "NodalBasis"

    def __init__(self, nodes, weights):
        self.nodes = _as_f64(nodes)
        self.weights = _as_f64(weights)
        x = self.nodes
        diff = x[:, None] - x[None, :]
        np.fill_diagonal(diff, 1.0)
        self.baryweights = 1.0 / np.prod(diff, axis=1)
        lam = self.baryweights
        D = (lam[None, :] / lam[:, None]) / diff
        np.fill_diagonal(D, 0.0)
        np.fill_diagonal(D, -D.sum(axis=1))
        # (kept exactly as the barycentric formula yields it, like PolynomialBases.jl: for symmetric node sets it is
        # centro-antisymmetric only up to rounding; the library detects that at upload, sbp_b200.h SBP_DENSE_*)
        self.D = D

    def interpolation_matrix(self, dest) -> np.ndarray:
        return interpolation_matrix(dest, self.nodes, self.baryweights)


def interpolation_matrix(dest, nodes, baryweights=None) -> np.ndarray:
    nodes = _as_f64(nodes)
    if baryweights is None:
        diff = nodes[:, None] - nodes[None, :]
        np.fill_diagonal(diff, 1.0)
        baryweights = 1.0 / np.prod(diff, axis=1)
    dest = np.atleast_1d(_as_f64(dest))
    d = dest[:, None] - nodes[None, :]
    tol = 2 * np.finfo(float).eps * np.maximum(1.0, np.abs(dest))[:, None]
    exact = np.abs(d) <= tol
    with np.errstate(divide="ignore", invalid="ignore"):
        t = baryweights[None, :] / d
        M = t / t.sum(axis=1, keepdims=True)
    hit = exact.any(axis=1)
    M[hit] = exact[hit].astype(np.float64)
    return M


def _newton_legendre(x, f_and_df, iters=60):
    x = np.asarray(x, dtype=np.longdouble)
    for _ in range(iters):
        f, df = f_and_df(x)
        step = f / df
        x = x - step
        if np.max(np.abs(step)) < 1e-19:
            break
    return x


def _legendre(n, x):
    """(P_n, P_n') at x in extended precision."""
    x = np.asarray(x, dtype=np.longdouble)
    pm, p = np.ones_like(x), x.copy()
    if n == 0:
        return pm, np.zeros_like(x)
    for k in range(1, n):
        pm, p = p, ((2 * k + 1) * x * p - k * pm) / (k + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = n * (pm - x * p) / (1 - x * x)
    return p, dp


class LobattoLegendre(NodalBasis):
    name = "LobattoLegendre"

    def __init__(self, p: int, T=np.float64):
        N = p + 1
        if p == 1:
            x, w = np.array([-1.0, 1.0]), np.array([1.0, 1.0])
        else:
            # interior nodes: roots of P_p'; start from Gauss-Chebyshev-Lobatto points
            xi = -np.cos(np.pi * np.arange(1, p) / p)

            def f(x):
                P, dP = _legendre(p, x)
                # d/dx of P_p' from the Legendre ODE: (1-x^2) P'' = 2x P' - p(p+1) P
                return dP, (2 * x * dP - p * (p + 1) * P) / (1 - x * x)

            xi = _newton_legendre(xi, f)
            x = np.concatenate([[-1.0], np.asarray(xi, dtype=np.longdouble), [1.0]])
            x = 0.5 * (x - x[::-1])
            P, _ = _legendre(p, x)
            w = 2.0 / (p * (p + 1) * P * P)
        super().__init__(np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64))
        assert len(self.nodes) == N


class GaussLegendre(NodalBasis):
    name = "GaussLegendre"

    def __init__(self, p: int, T=np.float64):
        N = p + 1
        x0, _ = np.polynomial.legendre.leggauss(N)
        x = _newton_legendre(x0, lambda x: _legendre(N, x))
        x = 0.5 * (x - x[::-1])
        _, dP = _legendre(N, x)
        w = 2.0 / ((1 - x * x) * dP * dP)
        super().__init__(np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64))


class GaussRadauLeft(NodalBasis):
    name = "GaussRadauLeft"

    def __init__(self, p: int, T=np.float64):
        N = p + 1
        # free nodes: roots of (P_{N-1} + P_N)/(1 + x), initial guess from numpy's companion matrix, Newton polish
        c = np.zeros(N + 1)
        c[N - 1], c[N] = 1.0, 1.0  # P_{N-1} + P_N
        r = np.sort(np.real(np.polynomial.legendre.legroots(c)))
        x0 = r[1:]  # drop the root at -1

        def f(x):
            Pn, dPn = _legendre(N, x)
            Pm, dPm = _legendre(N - 1, x)
            q = (Pm + Pn) / (1 + x)
            dq = (dPm + dPn) / (1 + x) - (Pm + Pn) / (1 + x) ** 2
            return q, dq

        xi = _newton_legendre(x0, f)
        Pm, _ = _legendre(N - 1, xi)
        wi = (1 - xi) / (N * N * Pm * Pm)
        x = np.concatenate([[np.longdouble(-1.0)], xi])
        w = np.concatenate([[np.longdouble(2.0) / (N * N)], wi])
        super().__init__(np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64))


class GaussRadauRight(NodalBasis):
    name = "GaussRadauRight"

    def __init__(self, p: int, T=np.float64):
        left = GaussRadauLeft(p)
        super().__init__(-left.nodes[::-1].copy(), left.weights[::-1].copy())


# ------------------------------------------------------------------------------------------------
# PolynomialBasesDerivativeOperator  (src/polynomialbases_operators.jl)
# ------------------------------------------------------------------------------------------------


class PolynomialBasesDerivativeOperator(DerivativeOperator):
    """src/polynomialbases_operators.jl:7-31.  `PolynomialBasesDerivativeOperator(basis_type, xmin, xmax, N)`
    or `PolynomialBasesDerivativeOperator(xmin, xmax, basis)`."""

    def __init__(self, *args):
        super().__init__()
        if len(args) == 4:
            basis_type, xmin, xmax, N = args
            if N < 2:
                raise ArgumentError("N >= 2 must hold")  # @argcheck N >= 2 (:42)
            basis = basis_type(N - 1, np.float64)
        elif len(args) == 3:
            xmin, xmax, basis = args
        else:
            raise ArgumentError("PolynomialBasesDerivativeOperator(basis_type, xmin, xmax, N) or (xmin, xmax, basis)")
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.basis = basis
        self.jac = 2.0 / (self.xmax - self.xmin)  # :26
        self.dx = 1.0 / self.jac  # Δx (:27)
        self.grid = (basis.nodes + 1) / 2 * (self.xmax - self.xmin) + self.xmin  # map_from_canonical (:25)

    def _dense(self):
        return self.basis.D

    def _scale(self):
        return self.jac  # mul!(dest, basis.D, u, α*jac, β) (:148)

    def weights(self):
        return self.dx * self.basis.weights  # :70-72

    def mass_matrix_boundary(self):
        R = self.basis.interpolation_matrix([-1.0, 1.0])  # mass_matrix_boundary(D.basis) (:74-76)
        return np.outer(R[1], R[1]) - np.outer(R[0], R[0])

    def get_weight(self, i: int):
        if not 1 <= i <= self.N:
            raise ArgumentError("1 <= i <= N must hold")  # :120-122 (1-based like the reference)
        return self.dx * self.basis.weights[i - 1]

    derivative_order = 1
    issymmetric = False

    @property
    def accuracy_order(self):
        return self.N - 1  # :159-161

    lower_bandwidth = upper_bandwidth = property(lambda self: self.N - 1)
    left_boundary_weight = property(lambda self: self.dx * self.basis.weights[0])
    right_boundary_weight = property(lambda self: self.dx * self.basis.weights[-1])

    def __repr__(self):
        return (f"First derivative operator {{T=Float64}} on {self.N} {self.basis.name} nodes in "
                f"[{self.xmin}, {self.xmax}]")


def polynomialbases_derivative_operator(basis_type, xmin=None, xmax=None, N=None):
    """src/polynomialbases_operators.jl:48-61 (positional or keyword form)."""
    if xmin is None or xmax is None or N is None:
        raise ArgumentError("xmin, xmax and N are required")
    return PolynomialBasesDerivativeOperator(basis_type, float(xmin), float(xmax), int(N))


# ------------------------------------------------------------------------------------------------
# upstream MatrixDerivativeOperator (what function_space_operator returns, ext/function_space_operators.jl:13-14)
# ------------------------------------------------------------------------------------------------


class MatrixDerivativeOperator(DerivativeOperator):
    """`MatrixDerivativeOperator(xmin, xmax, nodes, weights, D, accuracy_order, source)`; the dense matrix is
    rescaled by jac = (last(nodes)-first(nodes))/(xmax-xmin) as upstream does (== 1 for the reference's call)."""

    def __init__(self, xmin, xmax, nodes, weights, D, accuracy_order=0, source=None, storage="auto"):
        super().__init__()
        nodes, weights, D = _as_f64(nodes), _as_f64(weights), _as_f64(D)
        N = len(nodes)
        if weights.shape != (N,) or D.shape != (N, N):
            raise ArgumentError("nodes, weights and D have inconsistent sizes")
        jac = (nodes[-1] - nodes[0]) / (float(xmax) - float(xmin))
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.grid = (nodes - nodes[0]) / jac + self.xmin if jac != 1.0 else nodes.copy()
        self._w = weights / jac
        self.D = jac * D
        self.accuracy_order = accuracy_order
        self.source = source
        self.storage = storage

    def _dense(self):
        return self.D

    def weights(self):
        return self._w

    def mass_matrix_boundary(self):
        B = np.zeros((self.N, self.N))
        B[0, 0], B[-1, -1] = -1.0, 1.0
        return B

    derivative_order = 1
    left_boundary_weight = property(lambda self: self._w[0])
    right_boundary_weight = property(lambda self: self._w[-1])

    def get_weight(self, i: int):
        if not 1 <= i <= self.N:
            raise ArgumentError("1 <= i <= N must hold")
        return self._w[i - 1]


def legendre_derivative_operator(xmin, xmax, N) -> MatrixDerivativeOperator:
    """Upstream LGL operator on [xmin, xmax] (used in test/test_subcell_operators.jl:132-133)."""
    b = LobattoLegendre(N - 1)
    jac = 2.0 / (float(xmax) - float(xmin))
    nodes = (b.nodes + 1) / 2 * (float(xmax) - float(xmin)) + float(xmin)
    op = MatrixDerivativeOperator(nodes[0], nodes[-1], nodes, b.weights / jac, jac * b.D, N - 1)
    op.xmin, op.xmax = float(xmin), float(xmax)
    op.grid[0], op.grid[-1] = float(xmin), float(xmax)
    return op


def mattsson_nordstrom_2004(accuracy_order: int, xmin: float, xmax: float, N: int) -> MatrixDerivativeOperator:
    """Upstream `derivative_operator(MattssonNordström2004(), 1, accuracy_order, xmin, xmax, N)` for orders 2, 4."""
    h = (xmax - xmin) / (N - 1)
    x = np.linspace(xmin, xmax, N)
    if accuracy_order == 2:
        D = (np.diag(np.full(N - 1, 0.5), 1) - np.diag(np.full(N - 1, 0.5), -1))
        D[0, :2] = [-1.0, 1.0]
        D[-1, -2:] = [-1.0, 1.0]
        w = np.ones(N)
        w[[0, -1]] = 0.5
    elif accuracy_order == 4:
        if N < 8:
            raise ArgumentError("need N >= 8 for the 4th-order operator")
        st = np.array([1 / 12, -2 / 3, 0.0, 2 / 3, -1 / 12])
        D = sum(np.diag(np.full(N - abs(o), c), o) for o, c in zip(range(-2, 3), st))
        blk = np.array([[-24 / 17, 59 / 34, -4 / 17, -3 / 34, 0, 0],
                        [-1 / 2, 0, 1 / 2, 0, 0, 0],
                        [4 / 43, -59 / 86, 0, 59 / 86, -4 / 43, 0],
                        [3 / 98, 0, -59 / 98, 0, 32 / 49, -4 / 49]])
        D[:4, :] = 0.0
        D[-4:, :] = 0.0
        D[:4, :6] = blk
        D[-4:, -6:] = -blk[::-1, ::-1]
        w = np.ones(N)
        w[:4] = [17 / 48, 59 / 48, 43 / 48, 49 / 48]
        w[-4:] = w[:4][::-1]
    else:
        raise ArgumentError("only accuracy orders 2 and 4 are available")
    return MatrixDerivativeOperator(xmin, xmax, x, h * w, D / h, accuracy_order, "MattssonNordström2004")


# ------------------------------------------------------------------------------------------------
# upstream MultidimensionalMatrixDerivativeOperator + corner-aware boundary operator (src/utils/corners.jl)
# ------------------------------------------------------------------------------------------------


def find_corners(boundary_indices):
    """src/utils/corners.jl:6-18 -> (corners_x, corners_y): positions (0-based) into `boundary_indices` of the
    first / second occurrence of every node index that appears exactly twice."""
    bi = np.asarray(boundary_indices)
    vals, counts = np.unique(bi, return_counts=True)
    cx, cy = [], []
    for v in vals[counts > 1]:  # sorted, like sort(findall(>(1), countmap(...)))
        pos = np.flatnonzero(bi == v)
        if len(pos) == 2:
            cx.append(int(pos[0]))
            cy.append(int(pos[1]))
    return cx, cy


class _DirectionOperator(DerivativeOperator):
    """`D[dim]` of a multidimensional operator: applies Ds[dim]; shares weights/grid with its parent."""

    def __init__(self, parent, dim):
        super().__init__()
        self.parent, self.dim = parent, dim
        self.grid = parent.grid
        self.storage = parent.storage

    def _dense(self):
        return self.parent.Ds[self.dim]

    def weights(self):
        return self.parent._w


class MultidimensionalMatrixDerivativeOperator:
    """`MultidimensionalMatrixDerivativeOperator(nodes, boundary_indices, normals, weights, weights_boundary, Ds,
    accuracy_order, source)` (built at ext/multidimensional_function_space_operators.jl:20-22).
    `boundary_indices` are 0-based here.  `D[dim]` (0-based dim) is an operator usable with `mul_`/`*`."""

    def __init__(self, nodes, boundary_indices, normals, weights, weights_boundary, Ds, accuracy_order=0, source=None,
                 storage="auto", corners=None):
        self.grid = _as_f64(nodes)
        self.boundary_indices = np.asarray(boundary_indices, dtype=np.int64)
        self.normals = _as_f64(normals)
        self._w = _as_f64(weights)
        self.weights_boundary = _as_f64(weights_boundary)
        self.Ds = tuple(_as_f64(D) for D in Ds)
        self.accuracy_order, self.source, self.storage = accuracy_order, source, storage
        self.corners = corners  # optional explicit (corners_x, corners_y); default: find_corners
        N = len(self._w)
        if self.grid.shape[0] != N or any(D.shape != (N, N) for D in self.Ds):
            raise ArgumentError("inconsistent operator sizes")
        if not (len(self.boundary_indices) == len(self.normals) == len(self.weights_boundary)):
            raise ArgumentError("boundary lists differ in length")
        self._dirs = tuple(_DirectionOperator(self, d) for d in range(len(self.Ds)))

    def __getitem__(self, dim: int) -> _DirectionOperator:
        return self._dirs[dim]

    def size(self, dim=None):
        N = len(self._w)
        return (N, N) if dim is None else N

    @property
    def N(self):
        return len(self._w)

    def weights(self):
        return self._w

    def mass_matrix(self):
        return np.diag(self._w)

    def restrict_boundary(self, x):
        return np.asarray(x)[self.boundary_indices]

    def weights_boundary_scaled(self, dim: int):
        return self.weights_boundary * self.normals[:, dim]

    def mass_matrix_boundary_diag(self, dim: int) -> np.ndarray:
        """src/utils/corners.jl:29-39 (diagonal of the returned Diagonal)."""
        bw = self.weights_boundary_scaled(dim)
        bi = self.boundary_indices
        corners = self.corners if self.corners is not None else find_corners(bi)
        keep = np.ones(len(bi), dtype=bool)
        keep[list(corners[dim])] = False  # deleteat!
        b = np.zeros(self.N)
        for idx, val in zip(bi[keep], bw[keep]):  # b[boundary_indices] .= weights: later entries overwrite
            b[idx] = val
        return b

    def mass_matrix_boundary(self, dim: int) -> np.ndarray:
        return np.diag(self.mass_matrix_boundary_diag(dim))


def tensor_product_operator_2D(D1: MatrixDerivativeOperator, D2: MatrixDerivativeOperator | None = None, storage="auto"):
    """Upstream `tensor_product_operator_2D`: nodes y-fastest; boundary entries [left | bottom | top | right] with the
    corner positions of test/test_MFSBP.jl:111-119 attached as `corners`."""
    D2 = D1 if D2 is None else D2
    Nx, Ny = D1.N, D2.N
    X, Y = np.meshgrid(D1.grid, D2.grid, indexing="ij")
    nodes = np.column_stack([X.ravel(), Y.ravel()])
    Dx = np.kron(D1._dense() * D1._scale(), np.eye(Ny))
    Dy = np.kron(np.eye(Nx), D2._dense() * D2._scale())
    w = np.kron(D1.weights(), D2.weights())
    jj, ii = np.arange(Ny), np.arange(Nx)
    bidx = np.concatenate([jj, ii * Ny, ii * Ny + Ny - 1, (Nx - 1) * Ny + jj])
    normals = np.concatenate([np.tile([-1.0, 0.0], (Ny, 1)), np.tile([0.0, -1.0], (Nx, 1)),
                              np.tile([0.0, 1.0], (Nx, 1)), np.tile([1.0, 0.0], (Ny, 1))])
    wb = np.concatenate([D2.weights(), D1.weights(), D1.weights(), D2.weights()])
    corners = ([Ny, Nx + Ny - 1, Nx + Ny, 2 * Nx + Ny - 1], [0, Ny - 1, 2 * Nx + Ny, 2 * (Nx + Ny) - 1])
    return MultidimensionalMatrixDerivativeOperator(nodes, bidx, normals, w, wb, (Dx, Dy),
                                                    min(D1.accuracy_order, D2.accuracy_order), "tensor_product",
                                                    storage=storage, corners=corners)


def neighborhood_sparsity_pattern(nodes, lengths) -> np.ndarray:
    """src/utils/sparsity_patterns.jl:48-60 (strict upper triangle, boolean)."""
    x = _as_f64(nodes)
    d = (x[:, None, :] - x[None, :, :]) / _as_f64(lengths)[None, None, :]
    return np.triu(np.sqrt(np.sum(d * d, axis=-1)) < 1.0, 1)


# ------------------------------------------------------------------------------------------------
# SubcellOperator (src/subcell_operators.jl)
# ------------------------------------------------------------------------------------------------


class SubcellOperator(DerivativeOperator):
    """src/subcell_operators.jl:36-85.  The apply matrix D = inv(Diagonal(w_L (+) w_R)) * (Q_left + Q_right)
    (:183-184) is formed ONCE at upload instead of on every mul! (:272)."""

    def __init__(self, nodes, x_M, weights_left, weights_right, Q_left, Q_right, B_left, B_right, e_L, e_M_L, e_M_R,
                 e_R, accuracy_order, source=None):
        super().__init__()
        nodes = _as_f64(nodes)
        if len(nodes) != len(weights_left) + len(weights_right):
            raise ArgumentError("If `x_M` is not in `nodes`, then the length of `nodes` must be equal to the sum of "
                                "the lengths of `weights_left` and `weights_right`.")  # :72-74
        self.grid, self.x_M = nodes, float(x_M)
        self.weights_left, self.weights_right = _as_f64(weights_left), _as_f64(weights_right)
        self.Q_left, self.Q_right = _as_f64(Q_left), _as_f64(Q_right)
        self.B_left, self.B_right = _as_f64(B_left), _as_f64(B_right)
        self.e_L, self.e_M_L, self.e_M_R, self.e_R = map(_as_f64, (e_L, e_M_L, e_M_R, e_R))
        self.accuracy_order, self.source = accuracy_order, source

    N_L = property(lambda self: len(self.weights_left))
    N_R = property(lambda self: len(self.weights_right))

    def get_weight_left(self, i: int):  # 1-based (:222-235)
        if not 1 <= i <= self.N:
            raise ArgumentError("1 <= i <= N must hold")
        return self.weights_left[i - 1] if i <= self.N_L else 0.0

    def get_weight_right(self, i: int):  # :237-250
        if not 1 <= i <= self.N:
            raise ArgumentError("1 <= i <= N must hold")
        return self.weights_right[i - 1 - (self.N - self.N_R)] if i > self.N - self.N_R else 0.0

    def get_weight(self, i: int):
        return self.get_weight_left(i) + self.get_weight_right(i)

    def weights(self):  # :121-123
        wl = np.concatenate([self.weights_left, np.zeros(self.N - self.N_L)])
        wr = np.concatenate([np.zeros(self.N - self.N_R), self.weights_right])
        return wl + wr

    def mass_matrix_left(self):
        return np.diag(np.concatenate([self.weights_left, np.zeros(self.N - self.N_L)]))

    def mass_matrix_right(self):
        return np.diag(np.concatenate([np.zeros(self.N - self.N_R), self.weights_right]))

    def mass_matrix_boundary_left(self):
        return self.B_left

    def mass_matrix_boundary_right(self):
        return self.B_right

    def mass_matrix_boundary(self):
        return self.B_left + self.B_right  # :151

    def grid_left(self):
        return self.grid[: self.N_L]

    def grid_right(self):
        return self.grid[self.N - self.N_R:]

    def derivative_matrix(self):
        return (1.0 / self.weights())[:, None] * (self.Q_left + self.Q_right)  # :183-184

    def _dense(self):
        return self.derivative_matrix()

    derivative_order = 1
    issymmetric = False
    lower_bandwidth = upper_bandwidth = property(lambda self: self.N - 1)
    left_boundary_weight = property(lambda self: self.weights_left[0])
    right_boundary_weight = property(lambda self: self.weights_right[-1])
    left_projection_left = property(lambda self: self.e_L)
    left_projection_right = property(lambda self: self.e_M_L)
    right_projection_left = property(lambda self: self.e_M_R)
    right_projection_right = property(lambda self: self.e_R)

    def __repr__(self):
        x = self.grid
        return (f"Sub-cell operator {{T=Float64}} on {len(x)} nodes in [{x[0]}, {x[-1]}] with {self.N_L} nodes in left "
                f"sub-cell [{x[0]}, {self.x_M}] and {self.N_R} nodes in right sub-cell [{self.x_M}, {x[-1]}]")


def couple_subcell(D_left, D_right, x_M, coupling="no") -> SubcellOperator:
    """src/subcell_operators.jl:328-400; coupling in {"no", "central", "plus", "minus"} (Val(:no) ... in Julia)."""
    gl, gr = _as_f64(D_left.grid), _as_f64(D_right.grid)
    N_L, N_R = len(gl), len(gr)
    tol = 1e-11
    if x_M < gl[-1] - tol:
        raise ArgumentError("Left sub-cell must be to the left of x_M.")
    if x_M > gr[0] + tol:
        raise ArgumentError("Right sub-cell must be to the right of x_M.")
    if coupling not in ("no", "central", "plus", "minus"):
        raise ArgumentError(f"Unknown coupling type {coupling}.")
    wl, wr = _as_f64(D_left.weights()), _as_f64(D_right.weights())
    Rl = interpolation_matrix([D_left.xmin, D_left.xmax], gl)
    Rr = interpolation_matrix([D_right.xmin, D_right.xmax], gr)
    e_L_l, e_R_l, e_L_r, e_R_r = Rl[0], Rl[1], Rr[0], Rr[1]
    Ql = wl[:, None] * D_left.Matrix()
    Qr = wr[:, None] * D_right.Matrix()
    UR, LL = np.zeros((N_L, N_R)), np.zeros((N_R, N_L))
    if coupling == "central":
        Ql = Ql - 0.5 * np.outer(e_R_l, e_R_l)
        UR = 0.5 * np.outer(e_R_l, e_L_r)
        Qr = Qr + 0.5 * np.outer(e_L_r, e_L_r)
        LL = -0.5 * np.outer(e_L_r, e_R_l)
    elif coupling == "plus":
        Ql = Ql - np.outer(e_R_l, e_R_l)
        UR = np.outer(e_R_l, e_L_r)
    elif coupling == "minus":
        Qr = Qr + np.outer(e_L_r, e_L_r)
        LL = -np.outer(e_L_r, e_R_l)
    Q_left = np.block([[Ql, UR], [np.zeros((N_R, N_L)), np.zeros((N_R, N_R))]])
    Q_right = np.block([[np.zeros((N_L, N_L)), np.zeros((N_L, N_R))], [LL, Qr]])
    zR, zL = np.zeros(N_R), np.zeros(N_L)
    return SubcellOperator(np.concatenate([gl, gr]), x_M, wl, wr, Q_left, Q_right, Q_left + Q_left.T,
                           Q_right + Q_right.T, np.concatenate([e_L_l, zR]), np.concatenate([e_R_l, zR]),
                           np.concatenate([zL, e_L_r]), np.concatenate([zL, e_R_r]),
                           min(D_left.accuracy_order, D_right.accuracy_order), "GlaubitzLampertWintersNordström2025")


class OperatorTable:
    """A family of G same-size operators (e.g. SubcellOperators with varying split point) applied to one batch:
    instance b uses operator `op_index[b]` (device int32) or `b % G` when no index is given.  Optionally fuses the
    discrete energy `integrate(abs2, u, D_g)` per instance and its sum over the batch."""

    def __init__(self, operators: Sequence[DerivativeOperator]):
        if not operators:
            raise ArgumentError("empty operator table")
        N = operators[0].size(1)
        if any(op.size(1) != N for op in operators):
            raise DimensionMismatch("all operators of a table must have the same size")
        self.operators = list(operators)
        self.N, self.G = N, len(operators)
        self._dev: dict[int, _DeviceOp] = {}

    def size(self, dim=None):
        return (self.N, self.N) if dim is None else self.N

    def device_op(self, ctx: Context) -> _DeviceOp:
        op = self._dev.get(id(ctx))
        if op is None or op.handle is None:
            op = upload_table(ctx, [o.Matrix() for o in self.operators], [o.weights() for o in self.operators])
            self._dev[id(ctx)] = op
        return op

    def mul_(self, dest: DeviceBatch, u: DeviceBatch, alpha=True, beta=False, op_index: int | None = None,
             energy: bool = False, energy_sum: bool = False):
        """Returns dest, or (dest, energies[B] host array or None, total or None) when energies are requested."""
        _check_batch(self, dest, u)
        ctx = u.ctx
        op = self.device_op(ctx)
        ebuf = RawBuffer(ctx, u.B * 8 + 16) if (energy or energy_sum) else None
        e_ptr = ebuf.ptr if energy else 0
        s_ptr = (ebuf.ptr + u.B * 8) if energy_sum else 0
        check(ctx._lib.sbp_apply_table(ctx.handle, op.handle, _ptr(dest.ptr), _ptr(u.ptr), u.B, _ptr(op_index),
                                       float(alpha), float(beta), _ptr(e_ptr), _ptr(s_ptr)))
        if ebuf is None:
            return dest
        vals = ebuf.to_host(np.float64, u.B + 1)
        ebuf.free()
        return dest, (vals[: u.B].copy() if energy else None), (float(vals[u.B]) if energy_sum else None)
