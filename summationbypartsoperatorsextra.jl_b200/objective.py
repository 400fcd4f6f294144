"""batched evaluation of the operator-construction objective on the device.

The reference minimises, on the CPU and one parameter vector at a time (Optim + ForwardDiff),
    f(x) = sum |S(sigma) V - P(rho) V_x + R|^2                                    (ext/function_space_operators_optim.jl:89-105)
    f(x) = sum_d |S_d V - P V_xd + B_d(phi) V / 2|^2 + |V' B_d(phi) V - M_d|^2    (ext/multidimensional_function_space_operators.jl:124-161)
Construction stays on the CPU (north_star); what the device offers is the evaluation of f for a whole population of candidates
in one launch (`sbp_objective_eval`), e.g. for multi-start / evolutionary strategies or the trial points of a line search.

The skew-symmetric S is described by an INDEX MAP: an N x N integer matrix with +-(l + 1) where S[i, j] = +-sigma[l] and 0 where
S is structurally zero -- exactly what `set_S!` (ext/utils.jl:83-148) writes for sigma = (1, 2, ..., L).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ArgumentError, DimensionMismatch, check
from .device import Context, DeviceBatch, RawBuffer


def get_nsigma(N, bandwidth=None, size_boundary=None, different_values=True, sparsity_pattern=None) -> int:
    """src/utils/optimization.jl:15-43."""
    b = N - 1 if bandwidth is None else bandwidth
    c = 2 * b if size_boundary is None else size_boundary
    if sparsity_pattern is not None:
        return int(np.count_nonzero(np.triu(np.asarray(sparsity_pattern, dtype=bool), 1)))
    if b == N - 1:
        return N * (N - 1) // 2
    if different_values:
        return N * b + b * (b + 1) // 2 + c * c - 2 * b * c - c
    return c * (c - 1) // 2 + b


def skew_index_map(N, bandwidth=None, size_boundary=None, different_values=True, sparsity_pattern=None) -> np.ndarray:
    """N x N int64 matrix M with M[i, j] = +-(l + 1) <=> set_S! writes S[i, j] = +-sigma[l] (0: structural zero)."""
    b = N - 1 if bandwidth is None else int(bandwidth)
    c = 2 * b if size_boundary is None else int(size_boundary)
    M = np.zeros((N, N), dtype=np.int64)
    if sparsity_pattern is not None:
        pat = np.triu(np.asarray(sparsity_pattern, dtype=bool), 1)
        ii, jj = np.nonzero(pat)  # row-major order == the i, j loops of set_S_sparsity_pattern!
        ids = np.arange(1, len(ii) + 1)
        M[ii, jj], M[jj, ii] = ids, -ids
        return M
    if b == N - 1:
        ii, jj = np.triu_indices(N, 1)
        ids = np.arange(1, len(ii) + 1)
        M[ii, jj], M[jj, ii] = ids, -ids
        return M
    if N < 2 * c + 1 or b < 1:
        raise ArgumentError("block-banded structure needs N > 2 * size_boundary and bandwidth >= 1")
    k = 1

    def skew(block, k):
        n = block.shape[0]
        for i in range(n):
            for j in range(i + 1, n):
                block[i, j], block[j, i] = k, -k
                k += 1
        return k

    M1 = M[:c, :c]
    k = skew(M1, k)
    M2 = M[N - c:, N - c:]
    if different_values:
        k = skew(M2, k)
    else:
        M2[:] = -M1[::-1, ::-1]
    Dm = M[c:N - c, c:N - c]
    n = Dm.shape[0]
    k0 = k
    for i in range(n):
        for j in range(i + 1, n):
            if j - i <= b:
                if different_values:
                    l, k = k, k + 1
                else:
                    l = k0 + (j - i) - 1
                Dm[i, j], Dm[j, i] = l, -l

    def triangular(Cb, k, init_k):
        n = Cb.shape[0]
        start = (n - b) if different_values else (c - b)  # 0-based first row
        for i in range(start, n):
            for j in range(i - start + 1):
                l = k if different_values else init_k - 1 + b + (j + 1) - (i - start + 1)
                Cb[i, j] = l
                k += 1
        return k

    C1 = M[:c, c:N - c]
    k = triangular(C1, k, k)
    M[c:N - c, :c] = -C1.T
    C2 = M[c:N - c, N - c:]
    if different_values:
        k = triangular(C2, k, k)
        M[N - c:, c:N - c] = -C2.T
    else:
        C1b = C1[::-1, ::-1]
        C2[:] = C1b.T
        M[N - c:, c:N - c] = -C1b
    return M


def _csr_from_index_maps(maps, offsets):
    """CSR rows {column, index into x, sign} of the D index maps (direction d's sigma starts at offsets[d] in x)."""
    N = maps[0].shape[0]
    rowptr, col, xidx, sign = [], [], [], []
    for d, M in enumerate(maps):
        rowptr.append(len(col))
        for i in range(N):
            js = np.flatnonzero(M[i])
            col.extend(js.tolist())
            xidx.extend((np.abs(M[i, js]) - 1 + offsets[d]).tolist())
            sign.extend(np.sign(M[i, js]).astype(np.float64).tolist())
            rowptr.append(len(col))
    return (np.asarray(rowptr, dtype=np.int32), np.asarray(col, dtype=np.int32), np.asarray(xidx, dtype=np.int32),
            np.asarray(sign, dtype=np.float64))


class ConstructionObjective:
    """The construction objective of a (multidimensional) function-space operator, uploaded once; `evaluate(X)` returns f for
    every row of X (host (C, xlen) array or a DeviceBatch) in one launch.

    1D (function_space_operators_optim.jl): ConstructionObjective(ctx, V, [V_x], x_length, [index_map], R=B V / 2)
    multi-D: ConstructionObjective(ctx, V, V_xis, vol, index_maps, moments=Ms, normals=..., boundary_indices=..., corners=...)"""

    def __init__(self, ctx: Context, V, V_xis, vol, index_maps, R=None, moments=None, normals=None, boundary_indices=None,
                 corners=None):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        V = f64(V)
        N, K = V.shape
        Vx = f64(np.stack([f64(v) for v in V_xis]))
        D = Vx.shape[0]
        maps = [np.asarray(m, dtype=np.int64) for m in index_maps]
        if Vx.shape != (D, N, K) or len(maps) != D or any(m.shape != (N, N) for m in maps):
            raise DimensionMismatch("V, V_xis and the index maps have inconsistent sizes")
        Ls = [int(np.max(np.abs(m))) for m in maps]
        offsets = np.concatenate([[0], np.cumsum(Ls)]).astype(np.int64)
        rowptr, col, xidx, sign = _csr_from_index_maps(maps, offsets)
        self.N, self.K, self.D, self.Ltot = N, K, D, int(offsets[-1])
        Rm = None if R is None else f64(np.reshape(R, (D, N, K)))
        Mm, bidx, bn, Nb = None, None, None, 0
        if moments is not None:
            if normals is None or boundary_indices is None:
                raise ArgumentError("the multidimensional objective needs normals and boundary_indices")
            Mm = f64(np.stack([f64(m) for m in moments]))
            normals, bi = f64(normals), np.asarray(boundary_indices, dtype=np.int64)
            Nb = len(bi)
            corners = corners if corners is not None else tuple([] for _ in range(D))
            bidx = np.full((D, N), -1, dtype=np.int32)
            bn = np.zeros((D, N))
            for d in range(D):  # set_B! (ext/utils.jl:179-190): later entries overwrite, corner entries of this direction are skipped
                skip = set(int(j) for j in corners[d])
                for j, k in enumerate(bi):
                    if j not in skip:
                        bidx[d, k], bn[d, k] = j, normals[j, d]
        self.Nb, self.xlen = Nb, self.Ltot + N + Nb
        self.ctx = ctx
        h = C.c_void_p()
        pp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        check(ctx._lib.sbp_objective_create(ctx.handle, N, K, D, Nb, self.Ltot, pp(rowptr), pp(col), pp(xidx), pp(sign), pp(V),
                                            pp(Vx), pp(Rm), pp(Mm), pp(bidx), pp(bn), float(vol), C.byref(h)))
        self._h = h

    def evaluate(self, X) -> np.ndarray:
        ctx = self.ctx
        if isinstance(X, DeviceBatch):
            x = X
        else:
            Xh = np.atleast_2d(np.ascontiguousarray(X, dtype=np.float64))
            x = ctx.upload(Xh)
        if x.N != self.xlen:
            raise DimensionMismatch(f"candidates have {x.N} entries, the objective takes {self.xlen}")
        out = RawBuffer(ctx, max(x.B, 1) * 8)
        check(ctx._lib.sbp_objective_eval(ctx.handle, self._h, C.c_void_p(x.ptr), x.B, C.c_void_p(out.ptr)))
        vals = out.to_host(np.float64, x.B)
        out.free()
        return vals

    def close(self):
        if self._h:
            self.ctx._lib.sbp_objective_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass
