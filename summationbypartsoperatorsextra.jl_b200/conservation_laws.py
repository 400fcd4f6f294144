"""Host-side mirror of src/conservation_laws/ (semidiscretizations + analysis callback) on device batches.

    semi = MultidimensionalLinearAdvectionNonperiodicSemidiscretization(D, a, bc)
    semi(du, u, p, t)                           # in-place RHS for every instance of the batch
    analyze_quantities(semi, du, u, p, t)       # (B, 7) host array in the reference's order
    cb = AnalysisCallback(semi; dt=0.01) / interval=...;  tstops(cb), quantities(cb)
    semidiscretize(u0func, semi, tspan)

The constructors do the reference's one-time host assembly (multidimensional_linear_advection.jl:22-41:
A = a1*Dx + a2*Dy, B = inv(P)*(a1*B_x + a2*B_y)) and resolve the sequential `set_bc!` loop (:60-71) into a
per-node mode (0 interior / 1 inflow / 2 outflow; later boundary entries overwrite earlier ones, i.e. corners
keep the LAST entry).  Each RHS call is ONE fused device pass (operator apply + SAT); `bc(x, t)` is a host
callable, evaluated at the inflow nodes only, once per distinct t, and uploaded as N doubles.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, check
from .device import Context, DeviceBatch, RawBuffer
from .operators import (DerivativeOperator, MultidimensionalMatrixDerivativeOperator, _DeviceOp, _as_f64, _ptr,
                        upload_dense, upload_sparse)


class AbstractSemidiscretization:
    pass


class MultidimensionalLinearAdvectionNonperiodicSemidiscretization(AbstractSemidiscretization):
    """u_t + a . grad u = 0 with boundary data bc(x, t)
    (src/conservation_laws/multidimensional_linear_advection.jl:10-42)."""

    def __init__(self, derivative: MultidimensionalMatrixDerivativeOperator, a, bc, storage: str | None = None,
                 bc_vectorized: bool = False):
        """`bc_vectorized=True` declares that `bc(x, t)` accepts coordinate ARRAYS (x = (xs, ys), one entry per inflow
        node) and returns the array of boundary values; the first evaluation is checked against the node-by-node call at
        every inflow node.  By default `bc` is called node by node, as the reference's `set_bc!` does (:65-66)."""
        if len(a) != 2 or len(derivative.Ds) != 2:
            raise ArgumentError("Only 2D is supported")  # "Only support 2D for now" (:30)
        D = derivative
        self.derivative, self.a, self.bc = D, (float(a[0]), float(a[1])), bc
        self.storage = storage or D.storage
        A = self.a[0] * D.Ds[0] + self.a[1] * D.Ds[1]  # :33
        invP = 1.0 / D.weights()  # :34
        B_x = D.mass_matrix_boundary_diag(0)  # :35
        B_y = D.mass_matrix_boundary_diag(1)  # :36
        B = invP * (self.a[0] * B_x + self.a[1] * B_y)  # :37
        mode = np.zeros(D.N, dtype=np.int32)
        an = D.normals @ np.asarray(self.a)
        for i, j in enumerate(D.boundary_indices):  # set_bc! loop order (:62-70)
            mode[j] = 1 if an[i] < 0 else 2
        self.cache = {"A": A, "B": B, "tmp1": None}
        self.mode = mode
        self.inflow_nodes = np.flatnonzero(mode == 1)
        self.bnd_tau = D.weights_boundary * an  # tau_i of analyze_quantities (:101-103)
        self._dev: dict[int, tuple] = {}
        self._g_cache: dict[int, tuple] = {}
        self._bc_vectorized, self._bc_checked = bool(bc_vectorized), False

    # -- device bundle ----------------------------------------------------------------------------------
    def _want_sparse(self, A):
        if self.storage == "dense":
            return False
        if self.storage == "sparse":
            return True
        return A.shape[0] > 128 and np.count_nonzero(A) < 0.15 * A.size

    def device_op(self, ctx: Context) -> _DeviceOp:
        ent = self._dev.get(id(ctx))
        if ent is None:
            D = self.derivative
            A = self.cache["A"]
            inner = upload_sparse(ctx, A, D.weights()) if self._want_sparse(A) else upload_dense(ctx, A, D.weights())
            b, w = _as_f64(self.cache["B"]), _as_f64(D.weights())
            node = np.ascontiguousarray(D.boundary_indices, dtype=np.int32)
            tau = _as_f64(self.bnd_tau)
            h = C.c_void_p()
            check(ctx._lib.sbp_op_advection2d_create(ctx.handle, inner.handle, b.ctypes.data_as(C.c_void_p),
                                                     w.ctypes.data_as(C.c_void_p),
                                                     self.mode.ctypes.data_as(C.c_void_p), len(node),
                                                     node.ctypes.data_as(C.c_void_p), tau.ctypes.data_as(C.c_void_p),
                                                     C.byref(h)))
            ent = (_DeviceOp(ctx, h), inner)
            self._dev[id(ctx)] = ent
        return ent[0]

    def boundary_values(self, t: float) -> np.ndarray:
        """g_j = bc(x_j, t) at inflow nodes (0 elsewhere) -- the data `set_bc!` writes into tmp1 (:65-66).  `bc` is called
        node by node unless the semidiscretization was built with `bc_vectorized=True` (one call with coordinate arrays,
        validated against the scalar calls at every inflow node the first time; a mismatch raises)."""
        g = np.zeros(self.derivative.N)
        nodes = self.derivative.grid
        idx = np.asarray(self.inflow_nodes, dtype=np.int64)
        if idx.size == 0:
            return g
        if self._bc_vectorized:
            vals = np.asarray(self.bc(nodes[idx].T, t), dtype=np.float64)
            if vals.shape != (idx.size,):
                raise ArgumentError(f"bc_vectorized: bc returned shape {vals.shape}, expected ({idx.size},)")
            if not self._bc_checked:
                ref = np.array([float(self.bc(nodes[j], t)) for j in idx])
                if not np.allclose(vals, ref, rtol=1e-13, atol=1e-300):
                    raise ArgumentError("bc_vectorized: the array call of bc disagrees with its node-by-node evaluation")
                self._bc_checked = True
        else:
            vals = [self.bc(nodes[j], t) for j in idx]
        g[idx] = vals
        return g

    def boundary_table(self, ctx: Context, times) -> RawBuffer:
        """Boundary data of several evaluation times in one upload: row k = boundary_values(times[k]) (N doubles)."""
        tab = np.stack([self.boundary_values(float(t)) for t in times]) if len(times) else np.zeros((0, self.derivative.N))
        return ctx.upload_array(tab)

    def _device_g(self, ctx: Context, t: float):
        ent = self._g_cache.get(id(ctx))
        if ent is None or ent[0] != t:
            if ent is not None:
                ent[1].free()
            ent = (t, ctx.upload_array(self.boundary_values(t)))
            self._g_cache[id(ctx)] = ent
        return ent[1]

    def __call__(self, du: DeviceBatch, u: DeviceBatch, p=None, t: float = 0.0, g: DeviceBatch | None = None,
                 quantities: RawBuffer | None = None):
        """RHS functor (:73-81).  `g` optionally supplies per-instance boundary data as a (B, N) device batch
        (only inflow entries are read); by default bc(x, t) is evaluated on the host and shared by the batch."""
        N = self.derivative.N
        if u.N != N or du.N != N or du.B != u.B:
            raise DimensionMismatch("du, u and the operator differ in size")
        ctx = u.ctx
        op = self.device_op(ctx)
        if g is None:
            gptr, gstride = self._device_g(ctx, float(t)).ptr, 0
        else:
            if g.shape != u.shape:
                raise DimensionMismatch("per-instance boundary data must have the shape of u")
            gptr, gstride = g.ptr, N
        check(ctx._lib.sbp_rhs_advection2d(ctx.handle, op.handle, _ptr(du.ptr), _ptr(u.ptr), u.B, _ptr(gptr),
                                           gstride, _ptr(quantities.ptr) if quantities is not None else None))
        return None

    def rhs_host(self, ctx: Context, hdu: np.ndarray, hu: np.ndarray, t: float = 0.0, chunk_instances: int = 0):
        """`semi(du, u, p, t)` for HOST matrices (B, N): chunks travel H2D -> fused RHS kernel -> D2H through the library's
        pinned-staging pipeline (`sbp_rhs_advection2d_host`).  Use `ctx.pinned_empty` arrays for full copy speed."""
        N = self.derivative.N
        if hu.ndim != 2 or hu.shape[1] != N or hdu.shape != hu.shape:
            raise DimensionMismatch("du, u and the operator differ in size")
        if not (hu.flags.c_contiguous and hdu.flags.c_contiguous and hu.dtype == np.float64 and hdu.dtype == np.float64):
            raise ArgumentError("host batches must be C-contiguous float64 arrays")
        g = _as_f64(self.boundary_values(float(t)))
        check(ctx._lib.sbp_rhs_advection2d_host(ctx.handle, self.device_op(ctx).handle, hdu.ctypes.data_as(C.c_void_p),
                                                hu.ctypes.data_as(C.c_void_p), hu.shape[0], g.ctypes.data_as(C.c_void_p),
                                                0, int(chunk_instances)))
        return None

    def stage(self, unext: DeviceBatch, u: DeviceBatch, t: float, ca: float, ch: float, cb: float = 0.0,
              extra: DeviceBatch | None = None, g: DeviceBatch | None = None, gptr: int | None = None):
        """Runge-Kutta stage update riding on the RHS pass: unext = ca*u + cb*extra + ch*semi(u, t)
        (sbp_rhs_advection2d_stage; `extra` may be `unext` itself, `unext` must not be `u`).  `gptr`: device address
        of boundary data shared by the batch that is already resident (a row of `boundary_table`)."""
        N = self.derivative.N
        if u.N != N or unext.shape != u.shape or (extra is not None and extra.shape != u.shape):
            raise DimensionMismatch("unext, u, extra and the operator differ in size")
        ctx = u.ctx
        op = self.device_op(ctx)
        if gptr is not None:
            gstride = 0
        elif g is None:
            gptr, gstride = self._device_g(ctx, float(t)).ptr, 0
        else:
            if g.shape != u.shape:
                raise DimensionMismatch("per-instance boundary data must have the shape of u")
            gptr, gstride = g.ptr, N
        check(ctx._lib.sbp_rhs_advection2d_stage(ctx.handle, op.handle, _ptr(unext.ptr), _ptr(u.ptr), u.B, _ptr(gptr),
                                                 gstride, float(ca), float(cb),
                                                 _ptr(extra.ptr) if extra is not None else None, float(ch)))
        return unext

    def __repr__(self):
        return ("Semidiscretization of the linear advection equation\n"
                "  $ \\partial_t u(x, t) + a\\cdot \\nabla u(x, t) = 0 $\nwith nonperiodic boundaries")


class VariableLinearAdvectionNonperiodicSemidiscretization(AbstractSemidiscretization):
    """Upstream 1D semidiscretization used by examples/RBF_FSBP_advection.jl:37-38:
    `VariableLinearAdvectionNonperiodicSemidiscretization(D, nothing, afunc, Val(false), left_bc, right_bc)`;
    du = -D*(a.*u) with Godunov SATs at both ends."""

    def __init__(self, derivative: DerivativeOperator, dissipation, afunc, split_form, left_bc, right_bc):
        if dissipation is not None:
            raise ArgumentError("artificial dissipation is not supported on the device path")
        if split_form not in (False, None):
            raise ArgumentError("only the non-split form (Val(false)) is supported")
        self.derivative = derivative
        self.a = np.array([afunc(x) for x in derivative.grid], dtype=np.float64)
        self.left_bc, self.right_bc = left_bc, right_bc
        self._dev: dict[int, tuple] = {}
        self._bc_cache: dict[int, tuple] = {}

    def device_op(self, ctx: Context) -> _DeviceOp:
        ent = self._dev.get(id(ctx))
        if ent is None:
            inner = self.derivative.device_op(ctx)
            a = _as_f64(self.a)
            h = C.c_void_p()
            check(ctx._lib.sbp_op_advection1d_create(ctx.handle, inner.handle, a.ctypes.data_as(C.c_void_p),
                                                     C.byref(h)))
            ent = (_DeviceOp(ctx, h), inner)
            self._dev[id(ctx)] = ent
        return ent[0]

    def _device_bc(self, ctx, t):
        ent = self._bc_cache.get(id(ctx))
        if ent is None or ent[0] != t:
            if ent is not None:
                ent[1].free()
            ent = (t, ctx.upload_array(np.array([self.left_bc(t), self.right_bc(t)], dtype=np.float64)))
            self._bc_cache[id(ctx)] = ent
        return ent[1]

    def __call__(self, du: DeviceBatch, u: DeviceBatch, p=None, t: float = 0.0, bc: DeviceBatch | None = None,
                 quantities: RawBuffer | None = None):
        N = self.derivative.N
        if u.N != N or du.N != N or du.B != u.B:
            raise DimensionMismatch("du, u and the operator differ in size")
        ctx = u.ctx
        op = self.device_op(ctx)
        if bc is None:
            bptr, bstride = self._device_bc(ctx, float(t)).ptr, 0
        else:
            if bc.shape != (u.B, 2):
                raise DimensionMismatch("per-instance boundary data must be (B, 2)")
            bptr, bstride = bc.ptr, 2
        check(ctx._lib.sbp_rhs_advection1d(ctx.handle, op.handle, _ptr(du.ptr), _ptr(u.ptr), u.B, _ptr(bptr), bstride,
                                           _ptr(quantities.ptr) if quantities is not None else None))
        return None

    def boundary_table(self, ctx: Context, times) -> RawBuffer:
        """Boundary data of several evaluation times in one upload: row k = [left_bc(t_k), right_bc(t_k)]."""
        tab = np.array([[self.left_bc(float(t)), self.right_bc(float(t))] for t in times], dtype=np.float64).reshape(-1, 2)
        return ctx.upload_array(tab)

    def stage(self, unext: DeviceBatch, u: DeviceBatch, t: float, ca: float, ch: float, cb: float = 0.0,
              extra: DeviceBatch | None = None, bc: DeviceBatch | None = None, gptr: int | None = None):
        """Runge-Kutta stage update riding on the RHS pass: unext = ca*u + cb*extra + ch*semi(u, t)
        (sbp_rhs_advection1d_stage; `extra` may be `unext` itself, `unext` must not be `u`).  `gptr`: device address
        of resident boundary data [left, right] shared by the batch (a row of `boundary_table`)."""
        N = self.derivative.N
        if u.N != N or unext.shape != u.shape or (extra is not None and extra.shape != u.shape):
            raise DimensionMismatch("unext, u, extra and the operator differ in size")
        ctx = u.ctx
        op = self.device_op(ctx)
        if gptr is not None:
            bptr, bstride = gptr, 0
        elif bc is None:
            bptr, bstride = self._device_bc(ctx, float(t)).ptr, 0
        else:
            if bc.shape != (u.B, 2):
                raise DimensionMismatch("per-instance boundary data must be (B, 2)")
            bptr, bstride = bc.ptr, 2
        check(ctx._lib.sbp_rhs_advection1d_stage(ctx.handle, op.handle, _ptr(unext.ptr), _ptr(u.ptr), u.B, _ptr(bptr),
                                                 bstride, float(ca), float(cb),
                                                 _ptr(extra.ptr) if extra is not None else None, float(ch)))
        return unext


def analyze_quantities(semi, du: DeviceBatch, u: DeviceBatch, p=None, t: float = 0.0, **kw) -> np.ndarray:
    """`analyze_quantities(semi, du, u, p, t)` (multidimensional_linear_advection.jl:83-108,
    analysis_callback.jl:136-173): re-evaluates the RHS into `du` (as the callback does, analysis_callback.jl:123)
    and returns the 7 quantities per instance, shape (B, 7) ((7,) for B == 1)."""
    if not isinstance(semi, (MultidimensionalLinearAdvectionNonperiodicSemidiscretization,
                             VariableLinearAdvectionNonperiodicSemidiscretization)):
        return np.zeros(0)  # fallback method (analysis_callback.jl:134)
    q = RawBuffer(u.ctx, max(u.B, 1) * 7 * 8)
    semi(du, u, p, t, quantities=q, **kw)
    out = q.to_host(np.float64, u.B * 7).reshape(u.B, 7)
    q.free()
    return out[0] if u.B == 1 else out


class ODEProblem:
    def __init__(self, f, u0, tspan, p=None):
        self.f, self.u0, self.tspan, self.p = f, u0, tspan, p


def semidiscretize(u0func, semi, tspan, ctx: Context | None = None, batch: int = 1) -> ODEProblem:
    """`semidiscretize(u0func, semi, tspan)` = ODEProblem{true}(semi, u0func.(grid), tspan); the initial state is
    uploaded as a batch of `batch` identical instances when a context is given (host array otherwise)."""
    nodes = semi.derivative.grid
    u0 = np.array([u0func(x) for x in nodes], dtype=np.float64)
    if ctx is not None:
        u0 = ctx.upload(np.tile(u0, (batch, 1)))
    return ODEProblem(semi, u0, tspan)


class AnalysisCallback:
    """src/conservation_laws/analysis_callback.jl:20-113: records `analyze_quantities` every `interval` accepted
    steps or every `dt` of simulated time.  Works with `solve_ssprk53` below."""

    def __init__(self, semi, interval: int = 0, dt: float | None = None):
        if interval > 0 and dt is not None:
            raise ArgumentError("Setting both `interval` and `dt` is not supported.")  # :78-80
        self.semi, self.interval, self.dt = semi, int(interval), dt
        self._tstops: list[float] = []
        self._quantities: list[np.ndarray] = []
        self._t0, self._k = None, 0

    def __call__(self, du, u, p, t):
        self._tstops.append(float(t))
        self._quantities.append(analyze_quantities(self.semi, du, u, p, t))

    def should_fire(self, step: int, t: float, final: bool) -> bool:
        """interval mode: a DiscreteCallback without `initialize` -- checked after every accepted step, so the first record
        comes after `interval` steps, not at step 0 (analysis_callback.jl:85-105).  dt mode: a PeriodicCallback with
        initial_affect and final_affect: records at t0, at every t0 + k dt (the driver clips its steps to land on them,
        `stop_times`) and at the final time."""
        if self.dt is not None:
            if self._t0 is None:
                self._t0, self._k = t, 0  # initial_affect = true
                return True
            hit = False
            while self._t0 + (self._k + 1) * self.dt <= t + 1e-12 * max(1.0, abs(t)):
                self._k += 1
                hit = True
            return hit or final
        return self.interval > 0 and step > 0 and (step % self.interval == 0 or final)

    def stop_times(self, t0: float, tend: float) -> list[float]:
        """The tstops a PeriodicCallback adds (multiples of dt after t0, inside the time span); empty in interval mode."""
        if self.dt is None:
            return []
        out, k = [], 1
        while t0 + k * self.dt < tend - 1e-12 * max(1.0, abs(tend)):
            out.append(t0 + k * self.dt)
            k += 1
        return out

    def __repr__(self):
        return f"AnalysisCallback(interval={self.interval}, dt={self.dt})"


def tstops(cb: AnalysisCallback):
    return cb._tstops


def quantities(cb: AnalysisCallback):
    return cb._quantities


_SSPRK53 = dict(
    a30=0.355909775063327, a32=0.644090224936674, a40=0.367933791638137, a43=0.632066208361863,
    a52=0.237593836598569, a54=0.762406163401431, b10=0.377268915331368, b21=0.377268915331368,
    b32=0.242995220537396, b43=0.238458932846290, b54=0.287632146308408,
    c1=0.377268915331368, c2=0.754537830662736, c3=0.728985661612188, c4=0.699226135931670)


def step_schedule(t0: float, tend: float, dt: float, stops=()) -> list[tuple[float, float]]:
    """(t, h) of every step of a fixed-step integration from t0 to tend: steps of size dt, clipped so that the integrator
    lands exactly on each stop time (the tstops a PeriodicCallback adds) and on tend, as OrdinaryDiffEq does."""
    marks = sorted(float(x) for x in stops if t0 < x < tend) + [float(tend)]
    out, t, mi = [], float(t0), 0
    while mi < len(marks):
        nxt = marks[mi]
        if nxt - t <= 1e-14 * max(1.0, abs(nxt)):
            mi += 1
            continue
        h = dt
        if t + h >= nxt - 1e-12 * max(1.0, abs(nxt)):  # the step would reach or pass the mark: land on it exactly
            out.append((t, nxt - t))
            t = nxt
            mi += 1
        else:
            out.append((t, h))
            t += h
    return out


def solve_ssprk53(prob: ODEProblem, dt: float, callback: AnalysisCallback | None = None) -> DeviceBatch:
    """Fixed-step SSPRK53 (`solve(ode, SSPRK53(); dt, adaptive=false)`, examples/RBF_FSBP_advection.jl:45-51), device
    resident: the time loop runs inside the library (`sbp_solve_ssprk53`), every stage is ONE launch -- the stage update
    rides on the RHS pass (sbp_rhs_advection*_stage: 104 B/DOF per step on the block-sparse kernel instead of 80 + 144 with
    separate RHS and update passes).  Semidiscretizations without device boundary tables take the RHS + `lincomb_` route.
    Returns the final state."""
    k = _SSPRK53
    u: DeviceBatch = prob.u0
    semi, p = prob.f, prob.p
    t0, tend = float(prob.tspan[0]), float(prob.tspan[1])
    u = u.copy()
    steps = step_schedule(t0, tend, dt, callback.stop_times(t0, tend) if callback is not None else ())
    cs = (0.0, k["c1"], k["c2"], k["c3"], k["c4"])
    if hasattr(semi, "boundary_table"):
        # device-resident: the library runs whole runs of steps (sbp_solve_ssprk53); the host only evaluates the boundary
        # functions (64 steps per upload) and comes back where the analysis callback fires
        ctx = u.ctx
        op = semi.device_op(ctx)
        du = u.similar() if callback is not None else None
        if callback is not None and callback.should_fire(0, t0, False):
            callback(du, u, p, t0)
        fires = [callback is not None and callback.should_fire(si + 1, steps[si + 1][0] if si + 1 < len(steps) else tend,
                                                               si + 1 == len(steps)) for si in range(len(steps))]
        si = 0
        while si < len(steps):
            sj = si
            while sj < len(steps) and sj - si < 64:
                sj += 1
                if fires[sj - 1]:
                    break
            seg = steps[si:sj]
            table = semi.boundary_table(ctx, [tt + c * hh for (tt, hh) in seg for c in cs])
            row = table.nbytes // (5 * len(seg)) // 8
            hs = np.ascontiguousarray([hh for (_, hh) in seg], dtype=np.float64)
            check(ctx._lib.sbp_solve_ssprk53(ctx.handle, op.handle, _ptr(u.ptr), u.B, _ptr(table.ptr), row, len(seg),
                                             hs.ctypes.data_as(C.c_void_p), None))
            ctx.synchronize()
            table.free()
            if fires[sj - 1]:
                callback(du, u, p, steps[sj][0] if sj < len(steps) else tend)
            si = sj
        return u
    u1, u2, u3, du = u.similar(), u.similar(), u.similar(), u.similar()
    if callback is not None and callback.should_fire(0, t0, False):
        callback(du, u, p, t0)
    for si, (t, h) in enumerate(steps):
        semi(du, u, p, t)
        u1.lincomb_(1.0, u, k["b10"] * h, du)
        semi(du, u1, p, t + k["c1"] * h)
        u2.lincomb_(1.0, u1, k["b21"] * h, du)
        semi(du, u2, p, t + k["c2"] * h)
        u3.lincomb_(k["a30"], u, k["a32"], u2, k["b32"] * h, du)
        semi(du, u3, p, t + k["c3"] * h)
        u1.lincomb_(k["a40"], u, k["a43"], u3, k["b43"] * h, du)
        semi(du, u1, p, t + k["c4"] * h)
        u.lincomb_(k["a52"], u2, k["a54"], u1, k["b54"] * h, du)
        tn = steps[si + 1][0] if si + 1 < len(steps) else tend
        if callback is not None and callback.should_fire(si + 1, tn, si + 1 == len(steps)):
            callback(du, u, p, tn)
    return u


# ------------------------------------------------------------------------------------------------------------------
# low-storage 3S*+ schemes with embedded error estimator and FSAL (the family of RDPK3SpFSAL49)
# ------------------------------------------------------------------------------------------------------------------
class LowStorage3SpTableau:
    """Coefficients of a 3S*+ FSAL scheme in OrdinaryDiffEqLowStorageRK's layout (`LowStorageRK3SpFSALConstantCache`):
    gamma1/gamma2/gamma3/delta/c for stages 2..s, beta and bhat for stages 1..s, bhat_fsal; `order` and `embedded_order`
    set the exponent divisor k = min(order, embedded_order) + 1 of the PID controller.

    RDPK3SpFSAL49 itself (examples/RBF_MFSBP_advection.jl:24) is NOT tabulated here: its 9-stage coefficients are the result
    of a numerical optimisation and live in OrdinaryDiffEqLowStorageRK, which is neither under /root/reference nor in this
    image.  `SBPB200.export_tableau(path, RDPK3SpFSAL49())` writes them on a Julia machine; `LowStorage3SpTableau.load(path)`
    reads that file.  The classical schemes below are written in the same form and are what the tests integrate with."""

    def __init__(self, gamma1, gamma2, gamma3, delta, c, beta, bhat, bhat_fsal, order, embedded_order, name=""):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.gamma1, self.gamma2, self.gamma3, self.delta, self.c = f(gamma1), f(gamma2), f(gamma3), f(delta), f(c)
        self.beta, self.bhat, self.bhat_fsal = f(beta), f(bhat), float(bhat_fsal)
        self.order, self.embedded_order, self.name = int(order), int(embedded_order), name
        s = len(self.beta)
        if len(self.bhat) != s or any(len(a) != s - 1 for a in (self.gamma1, self.gamma2, self.gamma3, self.delta, self.c)):
            raise ArgumentError("inconsistent tableau lengths")

    stages = property(lambda self: len(self.beta))

    @classmethod
    def load(cls, path):
        """Flat binary tableau written by `SBPB200.export_tableau` (SBPGOLD1 container, tests/golden/golden_io.py layout)."""
        import struct

        out = {}
        with open(path, "rb") as fh:
            if fh.read(8) != b"SBPGOLD1":
                raise ArgumentError(f"{path}: not a tableau file")
            (count,) = struct.unpack("<q", fh.read(8))
            for _ in range(count):
                (ln,) = struct.unpack("<q", fh.read(8))
                name = fh.read(ln).decode()
                (nd,) = struct.unpack("<q", fh.read(8))
                shape = struct.unpack(f"<{nd}q", fh.read(8 * nd)) if nd else ()
                n = int(np.prod(shape)) if nd else 1
                out[name] = np.frombuffer(fh.read(8 * n), dtype="<f8").copy()
        return cls(out["gamma1"], out["gamma2"], out["gamma3"], out["delta"], out["c"], out["beta"], out["bhat"],
                   float(out["bhat_fsal"][0]), int(out["order"][0]), int(out["embedded_order"][0]), name=path)

    @classmethod
    def ssprk33_heun(cls):
        """SSPRK33 (Shu-Osher) in 3S* form (gamma2 = delta = 0: S3 = uprev carries the convex combinations) with the embedded
        second-order weights (1/2, 1/2, 0): error weights b - bhat = (-1/3, -1/3, 2/3)."""
        return cls([0.25, 2.0 / 3.0], [0.0, 0.0], [0.75, 1.0 / 3.0], [0.0, 0.0], [1.0, 0.5], [1.0, 0.25, 2.0 / 3.0],
                   [-1.0 / 3.0, -1.0 / 3.0, 2.0 / 3.0], 0.0, 3, 2, name="SSPRK33 / Heun pair in 3S*+ form")

    @classmethod
    def heun_euler(cls):
        """Heun's method with the embedded Euler step (error weights (-1/2, 1/2)), in 3S* form."""
        return cls([0.5], [0.0], [0.5], [0.0], [1.0], [1.0, 0.5], [-0.5, 0.5], 0.0, 2, 1, name="Heun / Euler pair in 3S*+ form")


def solve_lowstorage_3sp(prob: ODEProblem, tableau: LowStorage3SpTableau, dt: float, adaptive: bool = True,
                         abstol: float = 1e-6, reltol: float = 1e-3, pid=(0.38, 0.18, 0.01),
                         callback: AnalysisCallback | None = None, max_steps: int = 1_000_000):
    """`solve(ode, alg; dt, abstol, reltol, callback)` for a 3S*+ FSAL low-storage scheme (alg = RDPK3SpFSAL49() in
    examples/RBF_MFSBP_advection.jl:21-27), device resident: the time loop, the five-vector stage updates, the embedded error
    norm and OrdinaryDiffEq's PID step size controller run inside the library (`sbp_solve_3sp_fsal`); the host evaluates the
    boundary function when the library asks for it and records the analysis callback at its stop times.
    Returns (u(t_end), {"naccept", "nreject", "dt_last"}).  `pid` defaults to the controller OrdinaryDiffEq attaches to
    RDPK3SpFSAL49; defaults of abstol / reltol are OrdinaryDiffEq's."""
    semi, p = prob.f, prob.p
    u: DeviceBatch = prob.u0.copy()
    ctx = u.ctx
    t0, tend = float(prob.tspan[0]), float(prob.tspan[1])
    op = semi.device_op(ctx)
    two_d = isinstance(semi, MultidimensionalLinearAdvectionNonperiodicSemidiscretization)
    row = semi.derivative.N if two_d else 2

    def bc_cb(t, h_row, _user):
        vals = semi.boundary_values(t) if two_d else (semi.left_bc(t), semi.right_bc(t))
        for i in range(row):
            h_row[i] = vals[i]

    if two_d:  # bulk copy instead of the element loop
        def bc_cb(t, h_row, _user):  # noqa: F811
            C.memmove(h_row, _as_f64(semi.boundary_values(t)).ctypes.data, 8 * row)

    du = u.similar() if callback is not None else None
    stops = np.ascontiguousarray(callback.stop_times(t0, tend) if callback is not None else [], dtype=np.float64)

    def stop_cb(t, _user):
        if callback is not None and callback.should_fire(-1, t, False):
            callback(du, u, p, t)

    tab = _lib.Tableau3Sp(tableau.stages, min(tableau.order, tableau.embedded_order) + 1,
                          tableau.gamma1.ctypes.data, tableau.gamma2.ctypes.data, tableau.gamma3.ctypes.data,
                          tableau.delta.ctypes.data, tableau.c.ctypes.data, tableau.beta.ctypes.data,
                          tableau.bhat.ctypes.data, tableau.bhat_fsal)
    pid_a = np.ascontiguousarray(pid, dtype=np.float64)
    stats = np.zeros(3)
    bc_f, stop_f = _lib.BC_FN(bc_cb), _lib.STOP_FN(stop_cb)
    if callback is not None and callback.should_fire(0, t0, False):
        callback(du, u, p, t0)
    check(ctx._lib.sbp_solve_3sp_fsal(ctx.handle, op.handle, _ptr(u.ptr), u.B, C.byref(tab), t0, tend, float(dt),
                                      1 if adaptive else 0, float(abstol), float(reltol), pid_a.ctypes.data_as(C.c_void_p),
                                      C.cast(bc_f, C.c_void_p), None, stops.ctypes.data_as(C.c_void_p), len(stops),
                                      C.cast(stop_f, C.c_void_p), None, int(max_steps), stats.ctypes.data_as(C.c_void_p)))
    ctx.synchronize()
    if callback is not None and callback.should_fire(-1, tend, True):
        callback(du, u, p, tend)
    return u, {"naccept": int(stats[0]), "nreject": int(stats[1]), "dt_last": float(stats[2])}
