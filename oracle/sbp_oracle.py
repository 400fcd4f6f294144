"""CPU oracle for the SBP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module.  The product (`summationbypartsoperatorsextra.jl_b200`) never
does; it fails loudly when its CUDA library is missing.

This is a plain numpy *restatement* of the reference algorithm (the reference is Julia and
cannot run in this image: no `julia` binary, no network).  Every function cites the
reference `file:line` it follows (paths relative to /root/reference).  Arithmetic that the
reference delegates to un-vendored dependencies is restated from their published algorithms:

* SummationByPartsOperators.jl (compat "0.5.81", Project.toml:58; no Manifest => unpinned):
  `MatrixDerivativeOperator`/`MultidimensionalMatrixDerivativeOperator` `mul!` (= dense
  gemv on the stored matrix), `mass_matrix`, `weights_boundary_scaled`,
  `restrict_boundary`, `tensor_product_operator_2D`, Mattsson-Nordstroem 2004 coefficients,
  `VariableLinearAdvectionNonperiodicSemidiscretization` RHS + Godunov flux.
* PolynomialBases.jl (compat "0.4.25", Project.toml:48): Lobatto/Gauss/Radau nodes and
  weights, barycentric weights, derivative and interpolation matrices.
* Julia LinearAlgebra `mul!` -> OpenBLAS dgemv (numpy here links OpenBLAS 0.3.30 too).

Pinning status
--------------
* apply (`mul!`) on PolynomialBases / subcell / tensor-product operators: PINNED through the
  reference's own known-answer tests, re-run on this oracle in tests/test_oracle.py
  (test/test_polynomialbases_operators.jl:29-59, test/test_subcell_operators.jl:127-285,
  test/test_unit.jl:44-67, test/test_MFSBP.jl:111-119).
* 2D advection RHS, `set_bc!`, `analyze_quantities`: **parity unpinned** -- the reference has
  no test asserting any value (test/test_unit.jl:108-118 only constructs and shows).  The
  formulas in src/conservation_laws/multidimensional_linear_advection.jl:60-108 are the sole
  specification and are restated literally below.
* 1D advection RHS: **parity unpinned** (upstream source absent); restated from the upstream
  algorithm and cross-checked with the conservation identity the reference itself encodes at
  src/conservation_laws/analysis_callback.jl:143-148.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# PolynomialBases.jl restatement (nodes / weights / barycentric machinery)
# --------------------------------------------------------------------------------------


def _legendre_and_derivative(p: int, x: np.ndarray):
    """P_p(x) and P_p'(x) by the three-term recurrence (extended precision inside)."""
    x = np.asarray(x, dtype=np.longdouble)
    p0 = np.ones_like(x)
    if p == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for n in range(1, p):
        p0, p1 = p1, ((2 * n + 1) * x * p1 - n * p0) / (n + 1)
    # (1-x^2) P_p' = p (P_{p-1} - x P_p)
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = p * (p0 - x * p1) / (1 - x * x)
    return p1, dp


def lobatto_legendre(N: int):
    """LobattoLegendre(p=N-1) nodes and weights on [-1,1] (PolynomialBases semantics used at
    src/polynomialbases_operators.jl:19-31).  Nodes = roots of (1-x^2) P'_p, w = 2/(p(p+1)P_p^2)."""
    p = N - 1
    assert p >= 1
    if p == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    # Chebyshev-Gauss-Lobatto initial guess, Newton on q(x) = P_{p+1} - P_{p-1} ~ (1-x^2)P_p'
    x = -np.cos(np.pi * np.arange(N) / p).astype(np.longdouble)
    for _ in range(100):
        Pp, _ = _legendre_and_derivative(p, x)
        Pm, _ = _legendre_and_derivative(p - 1, x)
        # f = x P_p - P_{p-1} (zero at LGL nodes incl. +-1), f' = (p+1) P_p
        f = x * Pp - Pm
        dx = f / ((p + 1) * Pp)
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce exact symmetry
    Pp, _ = _legendre_and_derivative(p, x)
    w = 2.0 / (p * (p + 1) * Pp * Pp)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def gauss_legendre(N: int):
    """GaussLegendre(p=N-1) nodes/weights: roots of P_N, w = 2/((1-x^2) P_N'^2)."""
    k = np.arange(1, N + 1)
    x = -np.cos(np.pi * (k - 0.25) / (N + 0.5)).astype(np.longdouble)
    for _ in range(100):
        P, dP = _legendre_and_derivative(N, x)
        dx = P / dP
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    x = 0.5 * (x - x[::-1])
    _, dP = _legendre_and_derivative(N, x)
    w = 2.0 / ((1 - x * x) * dP * dP)
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def gauss_radau_left(N: int):
    """GaussRadauLeft(p=N-1): node -1 fixed, others roots of (P_{N-1}+P_N)/(1+x);
    w_0 = 2/N^2, w_j = (1-x_j)/(N^2 P_{N-1}(x_j)^2)."""
    assert N >= 2
    n = N - 1  # number of free nodes
    # initial guess: Chebyshev-like points of the Radau family
    k = np.arange(1, n + 1)
    x = -np.cos(2 * np.pi * k / (2 * n + 1)).astype(np.longdouble)
    for _ in range(200):
        Pn, dPn = _legendre_and_derivative(N, x)
        Pm, dPm = _legendre_and_derivative(N - 1, x)
        f = (Pm + Pn) / (1 + x)
        df = (dPm + dPn) / (1 + x) - (Pm + Pn) / (1 + x) ** 2
        dx = f / df
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    x = np.sort(x)
    Pm, _ = _legendre_and_derivative(N - 1, x)
    w = (1 - x) / (N * N * Pm * Pm)
    nodes = np.concatenate([[np.longdouble(-1.0)], x])
    weights = np.concatenate([[np.longdouble(2.0) / (N * N)], w])
    return np.asarray(nodes, dtype=np.float64), np.asarray(weights, dtype=np.float64)


def gauss_radau_right(N: int):
    """GaussRadauRight(p=N-1): mirror image of the left rule."""
    x, w = gauss_radau_left(N)
    return -x[::-1].copy(), w[::-1].copy()


BASES = {
    "LobattoLegendre": lobatto_legendre,
    "GaussLegendre": gauss_legendre,
    "GaussRadauLeft": gauss_radau_left,
    "GaussRadauRight": gauss_radau_right,
}


def barycentric_weights(nodes):
    """PolynomialBases.barycentric_weights: lambda_j = 1/prod_{k!=j}(x_j-x_k)
    (used by src/utils/interpolation.jl:5)."""
    x = np.asarray(nodes, dtype=np.float64)
    N = len(x)
    lam = np.ones(N)
    for j in range(N):
        for k in range(N):
            if k != j:
                lam[j] *= x[j] - x[k]
    return 1.0 / lam


def derivative_matrix_nodal(nodes):
    """PolynomialBases.derivative_matrix: D[j,k] = (lambda_k/lambda_j)/(x_j-x_k), D[j,j] = -sum_{k!=j} D[j,k]
    (the `basis.D` multiplied at src/polynomialbases_operators.jl:148)."""
    x = np.asarray(nodes, dtype=np.float64)
    lam = barycentric_weights(x)
    N = len(x)
    D = np.zeros((N, N))
    for j in range(N):
        for k in range(N):
            if j != k:
                D[j, k] = (lam[k] / lam[j]) / (x[j] - x[k])
                D[j, j] -= D[j, k]
    return D


def interpolation_matrix(dest, nodes, baryweights=None):
    """PolynomialBases.interpolation_matrix(dest, nodes, baryweights): barycentric Lagrange
    interpolation rows; exact node hits give unit rows (src/utils/interpolation.jl:2-7)."""
    x = np.asarray(nodes, dtype=np.float64)
    lam = barycentric_weights(x) if baryweights is None else np.asarray(baryweights)
    dest = np.atleast_1d(np.asarray(dest, dtype=np.float64))
    M = np.zeros((len(dest), len(x)))
    for r, y in enumerate(dest):
        hit = np.where(np.isclose(y, x, rtol=0.0, atol=2 * np.finfo(float).eps * max(1.0, abs(y))))[0]
        if len(hit):
            M[r, hit[0]] = 1.0
            continue
        t = lam / (y - x)
        M[r, :] = t / t.sum()
    return M


# --------------------------------------------------------------------------------------
# PolynomialBasesDerivativeOperator  (src/polynomialbases_operators.jl)
# --------------------------------------------------------------------------------------


class PolynomialBasesOperator:
    """src/polynomialbases_operators.jl:7-31: jac = 2/(xmax-xmin), dx = 1/jac, grid mapped from
    canonical nodes; apply = (alpha*jac) * basis.D * u + beta*dest (:139-149)."""

    def __init__(self, basis_type: str, xmin: float, xmax: float, N: int):
        assert N >= 2  # @argcheck N >= 2  (:42)
        self.basis_type = basis_type
        self.nodes, self.basis_weights = BASES[basis_type](N)
        self.basis_D = derivative_matrix_nodal(self.nodes)
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.jac = 2.0 / (self.xmax - self.xmin)  # :26
        self.dx = 1.0 / self.jac  # :27
        # map_from_canonical: (x+1)/2*(xmax-xmin)+xmin  (:25)
        self.grid = (self.nodes + 1) / 2 * (self.xmax - self.xmin) + self.xmin
        self.N = N

    def weights(self):
        return self.dx * self.basis_weights  # :70-72

    def mass_matrix(self):
        return np.diag(self.weights())

    def mass_matrix_boundary(self):
        # mass_matrix_boundary(basis) (:74-76): e_R e_R' - e_L e_L' with the endpoint interpolation vectors
        R = interpolation_matrix([-1.0, 1.0], self.nodes)
        return np.outer(R[1], R[1]) - np.outer(R[0], R[0])

    def matrix(self):
        """Matrix(D): upstream fallback multiplies unit vectors -> jac*basis.D."""
        return self.jac * self.basis_D

    def mul(self, u, alpha=1.0, beta=0.0, dest=None):
        if len(u) != self.N or (dest is not None and len(dest) != self.N):
            raise ValueError("ArgumentError: N == length(u)")  # :143-146
        y = (alpha * self.jac) * (self.basis_D @ u)
        if beta != 0.0:
            y = y + beta * dest
        return y

    def integrate(self, func, u):
        return self.dx * float(np.sum(func(np.asarray(u)) * self.basis_weights))  # :66-68

    def scale_by_mass_matrix(self, u, factor=1.0):
        if len(u) != self.N:
            raise ValueError("DimensionMismatch")  # :86-89
        return factor * np.asarray(u) * (self.dx * self.basis_weights)  # :92-94

    def scale_by_inverse_mass_matrix(self, u, factor=1.0):
        if len(u) != self.N:
            raise ValueError("DimensionMismatch")
        return factor * np.asarray(u) / (self.dx * self.basis_weights)  # :110-112


# --------------------------------------------------------------------------------------
# Upstream finite-difference operators (Mattsson, Nordstroem 2004; diagonal norm)
# --------------------------------------------------------------------------------------


class MatrixOperator1D:
    """Upstream `MatrixDerivativeOperator` (built at ext/function_space_operators.jl:13-14):
    fields grid, weights, D; `mul!(dest, Dop, u, a, b) = mul!(dest, Dop.D, u, a, b)` (dense gemv)."""

    def __init__(self, grid, weights, D, accuracy_order=0):
        self.grid = np.asarray(grid, dtype=np.float64)
        self.w = np.asarray(weights, dtype=np.float64)
        self.D = np.asarray(D, dtype=np.float64)
        self.N = len(self.grid)
        self.accuracy_order = accuracy_order
        self.xmin, self.xmax = float(self.grid[0]), float(self.grid[-1])

    def weights(self):
        return self.w

    def matrix(self):
        return self.D

    def mass_matrix(self):
        return np.diag(self.w)

    def mass_matrix_boundary(self):
        B = np.zeros((self.N, self.N))
        B[0, 0], B[-1, -1] = -1.0, 1.0
        return B

    def mul(self, u, alpha=1.0, beta=0.0, dest=None):
        if len(u) != self.N:
            raise ValueError("DimensionMismatch")
        y = alpha * (self.D @ u)
        if beta != 0.0:
            y = y + beta * dest
        return y

    def integrate(self, func, u):
        return float(np.sum(func(np.asarray(u)) * self.w))


def mattsson_nordstrom_2004(accuracy_order: int, xmin: float, xmax: float, N: int) -> MatrixOperator1D:
    """First-derivative diagonal-norm SBP operators of Mattsson & Nordstroem (2004), orders 2 and 4
    (upstream `derivative_operator(MattssonNordström2004(), 1, p, xmin, xmax, N)`, used at
    test/test_unit.jl:45, test/test_MFSBP.jl:93-98)."""
    h = (xmax - xmin) / (N - 1)
    grid = np.linspace(xmin, xmax, N)
    D = np.zeros((N, N))
    w = np.full(N, 1.0)
    if accuracy_order == 2:
        for i in range(1, N - 1):
            D[i, i - 1], D[i, i + 1] = -0.5, 0.5
        D[0, 0], D[0, 1] = -1.0, 1.0
        D[-1, -2], D[-1, -1] = -1.0, 1.0
        w[0] = w[-1] = 0.5
    elif accuracy_order == 4:
        assert N >= 8
        for i in range(4, N - 4):
            D[i, i - 2 : i + 3] = [1 / 12, -2 / 3, 0.0, 2 / 3, -1 / 12]
        left = [
            [-24 / 17, 59 / 34, -4 / 17, -3 / 34, 0, 0],
            [-1 / 2, 0, 1 / 2, 0, 0, 0],
            [4 / 43, -59 / 86, 0, 59 / 86, -4 / 43, 0],
            [3 / 98, 0, -59 / 98, 0, 32 / 49, -4 / 49],
        ]
        for r, row in enumerate(left):
            D[r, :6] = row
            D[N - 1 - r, N - 6 :] = -np.asarray(row)[::-1]
        w[:4] = [17 / 48, 59 / 48, 43 / 48, 49 / 48]
        w[-4:] = w[:4][::-1]
    else:
        raise ValueError("only accuracy orders 2 and 4 are restated")
    return MatrixOperator1D(grid, h * w, D / h, accuracy_order)


def legendre_derivative_operator(xmin, xmax, N) -> MatrixOperator1D:
    """Upstream `legendre_derivative_operator(xmin, xmax, N)` (LGL nodal operator on [xmin,xmax]);
    used by test/test_subcell_operators.jl:132-133."""
    op = PolynomialBasesOperator("LobattoLegendre", xmin, xmax, N)
    m = MatrixOperator1D(op.grid, op.weights(), op.matrix(), N - 1)
    m.xmin, m.xmax = float(xmin), float(xmax)
    return m


# --------------------------------------------------------------------------------------
# Multidimensional operators (upstream MDMDO fields; tensor product; corners)
# --------------------------------------------------------------------------------------


class MultidimensionalOperator:
    """Upstream `MultidimensionalMatrixDerivativeOperator{2}` (built at
    ext/multidimensional_function_space_operators.jl:20-22): nodes, boundary_indices (0-based
    here), normals, weights, weights_boundary, Ds=(Dx,Dy)."""

    def __init__(self, nodes, boundary_indices, normals, weights, weights_boundary, Ds, accuracy_order=0):
        self.nodes = np.asarray(nodes, dtype=np.float64)  # (N, 2)
        self.boundary_indices = np.asarray(boundary_indices, dtype=np.int64)
        self.normals = np.asarray(normals, dtype=np.float64)  # (Nb, 2)
        self.w = np.asarray(weights, dtype=np.float64)
        self.weights_boundary = np.asarray(weights_boundary, dtype=np.float64)
        self.Ds = tuple(np.asarray(D, dtype=np.float64) for D in Ds)
        self.N = len(self.w)
        self.accuracy_order = accuracy_order

    def weights(self):
        return self.w

    def mul(self, u, dim, alpha=1.0, beta=0.0, dest=None):
        y = alpha * (self.Ds[dim] @ u)
        if beta != 0.0:
            y = y + beta * dest
        return y


def tensor_product_operator_2D(D1: MatrixOperator1D, D2: MatrixOperator1D | None = None) -> MultidimensionalOperator:
    """Upstream `tensor_product_operator_2D(D_1, D_2)`; node ordering y-fastest and boundary
    ordering [left edge | (bottom_i, top_i) interleaved over x | right edge] as implied by the
    corner positions listed at test/test_MFSBP.jl:111-119 (1-based: x-dir corners N_y+1, N_x+N_y,
    N_x+N_y+1, 2N_x+N_y; y-dir corners 1, N_y, 2N_x+N_y+1, 2(N_x+N_y)).

    NOTE: the x-direction corner list of the reference is only consistent with a
    *block* ordering bottom edge (N_x entries) then top edge (N_x entries):
    N_y+1 = first bottom entry (lower left), N_x+N_y = last bottom entry (lower right),
    N_x+N_y+1 = first top entry (upper left), 2N_x+N_y = last top entry (upper right).
    That block ordering is what is implemented."""
    if D2 is None:
        D2 = D1
    Nx, Ny = D1.N, D2.N
    X, Y = np.meshgrid(D1.grid, D2.grid, indexing="ij")  # node (i,j) -> i*Ny + j
    nodes = np.stack([X.ravel(), Y.ravel()], axis=1)
    Dx = np.kron(D1.D, np.eye(Ny))
    Dy = np.kron(np.eye(Nx), D2.D)
    weights = np.kron(D1.w, D2.w)
    bidx, normals, wb = [], [], []
    for j in range(Ny):  # left edge x = xmin
        bidx.append(0 * Ny + j), normals.append((-1.0, 0.0)), wb.append(D2.w[j])
    for i in range(Nx):  # bottom edge y = ymin
        bidx.append(i * Ny + 0), normals.append((0.0, -1.0)), wb.append(D1.w[i])
    for i in range(Nx):  # top edge y = ymax
        bidx.append(i * Ny + Ny - 1), normals.append((0.0, 1.0)), wb.append(D1.w[i])
    for j in range(Ny):  # right edge x = xmax
        bidx.append((Nx - 1) * Ny + j), normals.append((1.0, 0.0)), wb.append(D2.w[j])
    return MultidimensionalOperator(nodes, bidx, normals, weights, wb, (Dx, Dy), min(D1.accuracy_order, D2.accuracy_order))


def tensor_product_corners(Nx, Ny):
    """0-based positions (into boundary_indices) of the corner entries, test/test_MFSBP.jl:111-119."""
    corners_x_dir = [Ny, Nx + Ny - 1, Nx + Ny, 2 * Nx + Ny - 1]
    corners_y_dir = [0, Ny - 1, 2 * Nx + Ny, 2 * (Nx + Ny) - 1]
    return corners_x_dir, corners_y_dir


def find_corners(boundary_indices):
    """src/utils/corners.jl:6-18: duplicated node indices are corners; first occurrence goes to
    `corners_x`, second to `corners_y` (positions into boundary_indices, 0-based here)."""
    bi = list(boundary_indices)
    counts = {}
    for v in bi:
        counts[v] = counts.get(v, 0) + 1
    all_corners = sorted(v for v, c in counts.items() if c > 1)
    cx, cy = [], []
    for corner in all_corners:
        inds = [p for p, v in enumerate(bi) if v == corner]
        if len(inds) == 2:
            cx.append(inds[0])
            cy.append(inds[1])
    return cx, cy


def weights_boundary_scaled(D: MultidimensionalOperator, dim: int):
    """Upstream: weights_boundary[i] * normals[i][dim] (called at src/utils/corners.jl:31)."""
    return D.weights_boundary * D.normals[:, dim]


def mass_matrix_boundary_2d(D: MultidimensionalOperator, dim: int, corners=None):
    """src/utils/corners.jl:29-39: scaled boundary weights, entries `corners[dim]` deleted,
    b[boundary_indices] .= weights (ASSIGNMENT in list order, later entries overwrite)."""
    bw = list(weights_boundary_scaled(D, dim))
    bi = list(D.boundary_indices)
    if corners is None:
        corners = find_corners(bi)
    for pos in sorted(corners[dim], reverse=True):  # deleteat!
        del bw[pos]
        del bi[pos]
    b = np.zeros(D.N)
    for idx, val in zip(bi, bw):
        b[idx] = val
    return b  # diagonal of Diagonal(b)


def neighborhood_sparsity_pattern(nodes, lengths):
    """src/utils/sparsity_patterns.jl:48-60: upper-triangular boolean pattern,
    pattern[i,j] = ||(x_i-x_j)./lengths|| < 1 for i<j, diagonal false."""
    nodes = np.asarray(nodes, dtype=np.float64)
    N = len(nodes)
    diff = (nodes[:, None, :] - nodes[None, :, :]) / np.asarray(lengths)[None, None, :]
    pat = np.sqrt((diff**2).sum(-1)) < 1.0
    pat = np.triu(pat, 1)
    return pat


# --------------------------------------------------------------------------------------
# SubcellOperator  (src/subcell_operators.jl)
# --------------------------------------------------------------------------------------


class SubcellOperator:
    """src/subcell_operators.jl:36-85 fields; apply matrix rebuilt per call in the reference
    (`D = Matrix(Dop)` :272) = inv(Diagonal(w_L (+) w_R)) * (Q_left + Q_right) (:183-184)."""

    def __init__(self, nodes, x_M, weights_left, weights_right, Q_left, Q_right, B_left, B_right,
                 e_L, e_M_L, e_M_R, e_R, accuracy_order):
        if len(nodes) != len(weights_left) + len(weights_right):
            raise ValueError("ArgumentError: length(nodes) != N_L + N_R")  # :72-74
        self.grid = np.asarray(nodes, dtype=np.float64)
        self.x_M = float(x_M)
        self.weights_left = np.asarray(weights_left, dtype=np.float64)
        self.weights_right = np.asarray(weights_right, dtype=np.float64)
        self.Q_left, self.Q_right = np.asarray(Q_left), np.asarray(Q_right)
        self.B_left, self.B_right = np.asarray(B_left), np.asarray(B_right)
        self.e_L, self.e_M_L, self.e_M_R, self.e_R = e_L, e_M_L, e_M_R, e_R
        self.accuracy_order = accuracy_order
        self.N = len(self.grid)
        self.N_L = len(self.weights_left)
        self.N_R = len(self.weights_right)

    def get_weight_left(self, i):  # :222-235 (0-based i)
        return self.weights_left[i] if i < self.N_L else 0.0

    def get_weight_right(self, i):  # :237-250
        return self.weights_right[i - (self.N - self.N_R)] if i >= self.N - self.N_R else 0.0

    def weights(self):  # :121-123
        wl = np.array([self.get_weight_left(i) for i in range(self.N)])
        wr = np.array([self.get_weight_right(i) for i in range(self.N)])
        return wl + wr

    def mass_matrix(self):
        return np.diag(self.weights())

    def mass_matrix_boundary(self):  # :151
        return self.B_left + self.B_right

    def matrix(self):
        """derivative_matrix (:183-184): inv(Diagonal(w)) * (Q_left+Q_right); Julia's
        inv(Diagonal) forms 1/w_i first and then scales row i."""
        winv = 1.0 / self.weights()
        return winv[:, None] * (self.Q_left + self.Q_right)

    def mul(self, u, alpha=1.0, beta=0.0, dest=None):  # :270-280
        D = self.matrix()
        if len(u) != self.N or (dest is not None and len(dest) != self.N):
            raise ValueError("ArgumentError: N == length(u)")
        y = alpha * (D @ u)
        if beta != 0.0:
            y = y + beta * dest
        return y

    def integrate(self, func, u):  # :92-94
        return float(np.sum(func(np.asarray(u)) * self.weights()))

    def integrate_left(self, func, u):
        wl = np.array([self.get_weight_left(i) for i in range(self.N)])
        return float(np.sum(func(np.asarray(u)) * wl))

    def integrate_right(self, func, u):
        wr = np.array([self.get_weight_right(i) for i in range(self.N)])
        return float(np.sum(func(np.asarray(u)) * wr))

    def scale_by_mass_matrix(self, u, factor=1.0):  # :190-204
        if len(u) != self.N:
            raise ValueError("DimensionMismatch")
        return factor * np.asarray(u) * self.weights()

    def scale_by_inverse_mass_matrix(self, u, factor=1.0):  # :206-220
        if len(u) != self.N:
            raise ValueError("DimensionMismatch")
        return factor * np.asarray(u) / self.weights()


def couple_subcell(D_left, D_right, x_M, coupling="no") -> SubcellOperator:
    """src/subcell_operators.jl:328-400 restated.  D_left/D_right expose grid, weights(),
    matrix(), xmin, xmax, accuracy_order."""
    grid_left, grid_right = np.asarray(D_left.grid), np.asarray(D_right.grid)
    N_L, N_R = len(grid_left), len(grid_right)
    tol = 1e-11
    if x_M < grid_left[-1] - tol:
        raise ValueError("ArgumentError: Left sub-cell must be to the left of x_M.")  # :337-339
    if x_M > grid_right[0] + tol:
        raise ValueError("ArgumentError: Right sub-cell must be to the right of x_M.")  # :340-342
    nodes = np.concatenate([grid_left, grid_right])
    wl, wr = np.asarray(D_left.weights()), np.asarray(D_right.weights())
    R_left = interpolation_matrix([D_left.xmin, D_left.xmax], grid_left)  # :347-352
    e_L_l, e_R_l = R_left[0], R_left[1]
    R_right = interpolation_matrix([D_right.xmin, D_right.xmax], grid_right)  # :353-358
    e_L_r, e_R_r = R_right[0], R_right[1]
    Q_l = np.diag(wl) @ D_left.matrix()  # :360
    Q_r = np.diag(wr) @ D_right.matrix()  # :361
    UR = np.zeros((N_L, N_R))
    LL = np.zeros((N_R, N_L))
    if coupling == "no":
        pass
    elif coupling == "central":  # :365-369
        Q_l = Q_l - 0.5 * np.outer(e_R_l, e_R_l)
        UR = 0.5 * np.outer(e_R_l, e_L_r)
        Q_r = Q_r + 0.5 * np.outer(e_L_r, e_L_r)
        LL = -0.5 * np.outer(e_L_r, e_R_l)
    elif coupling == "plus":  # :370-373
        Q_l = Q_l - np.outer(e_R_l, e_R_l)
        UR = np.outer(e_R_l, e_L_r)
    elif coupling == "minus":  # :374-377
        Q_r = Q_r + np.outer(e_L_r, e_L_r)
        LL = -np.outer(e_L_r, e_R_l)
    else:
        raise ValueError("ArgumentError: Unknown coupling type")
    N = N_L + N_R
    Q_left = np.zeros((N, N))
    Q_left[:N_L, :N_L], Q_left[:N_L, N_L:] = Q_l, UR  # :381-382
    Q_right = np.zeros((N, N))
    Q_right[N_L:, :N_L], Q_right[N_L:, N_L:] = LL, Q_r  # :383-384
    B_left, B_right = Q_left + Q_left.T, Q_right + Q_right.T  # :386-387
    z_R, z_L = np.zeros(N_R), np.zeros(N_L)
    return SubcellOperator(nodes, x_M, wl, wr, Q_left, Q_right, B_left, B_right,
                           np.concatenate([e_L_l, z_R]), np.concatenate([e_R_l, z_R]),
                           np.concatenate([z_L, e_L_r]), np.concatenate([z_L, e_R_r]),
                           min(D_left.accuracy_order, D_right.accuracy_order))


# --------------------------------------------------------------------------------------
# 2D linear advection semidiscretization (src/conservation_laws/multidimensional_linear_advection.jl)
# --------------------------------------------------------------------------------------


class AdvectionSemi2D:
    """Constructor :22-41: A = a1*Dx + a2*Dy, B = inv(P) * (a1*B_x + a2*B_y) (Diagonal), tmp1."""

    def __init__(self, D: MultidimensionalOperator, a, bc, corners=None):
        self.D, self.a, self.bc = D, (float(a[0]), float(a[1])), bc
        self.A = self.a[0] * D.Ds[0] + self.a[1] * D.Ds[1]  # :33
        invP = 1.0 / D.w  # :34  inv(Diagonal)
        B_x = mass_matrix_boundary_2d(D, 0, corners)  # :35
        B_y = mass_matrix_boundary_2d(D, 1, corners)  # :36
        self.B = invP * (self.a[0] * B_x + self.a[1] * B_y)  # :37 (diagonal)
        self.tmp1 = np.zeros(D.N)  # :38

    def set_bc(self, u, t):
        """:60-71, sequential loop => last writer wins on duplicated (corner) nodes."""
        D = self.D
        self.tmp1[:] = 0.0
        for i, j in enumerate(D.boundary_indices):
            node = D.nodes[j]
            normal = D.normals[i]
            if normal[0] * self.a[0] + normal[1] * self.a[1] < 0:  # inflow
                self.tmp1[j] = self.bc(node, t)
            else:  # outflow
                self.tmp1[j] = u[j]

    def rhs(self, u, t):
        """:73-81  du .= -A*u + B*(u - tmp1)   (parses as ((-A)*u) + (B*(u-tmp1)))"""
        self.set_bc(u, t)
        return (-self.A) @ u + self.B * (u - self.tmp1)

    def analyze_quantities(self, du, u):
        """:83-108 (tmp1 must hold the boundary values of the matching set_bc! call)."""
        D = self.D
        P = D.w
        mass = np.sum(P * u)
        mass_rate = np.sum(P * du)
        mass_rate_boundary = mass_rate + np.sum(P * self.B * self.tmp1)  # :90
        energy = 0.5 * np.sum(P * (u**2))  # :92
        energy_rate = np.sum(P * (du * u))  # :93
        energy_rate_boundary_dissipation = energy_rate - 0.5 * np.dot(u, P * self.B * (u - 2 * self.tmp1))  # :94-96
        energy_rate_boundary = energy_rate
        for i, j in enumerate(D.boundary_indices):  # :98-105
            normal = D.normals[i]
            an = normal[0] * self.a[0] + normal[1] * self.a[1]
            tau = D.weights_boundary[i] * an
            energy_rate_boundary += 0.5 * tau * self.tmp1[j] ** 2
        return np.array([mass, mass_rate, mass_rate_boundary, energy, energy_rate,
                         energy_rate_boundary, energy_rate_boundary_dissipation])


# --------------------------------------------------------------------------------------
# 1D variable linear advection (upstream; used by examples/RBF_FSBP_advection.jl:37-38)
# --------------------------------------------------------------------------------------


def godunov_flux_variablelinearadvection(ul, ur, a):
    """Upstream SummationByPartsOperators.godunov_flux_variablelinearadvection (called at
    src/conservation_laws/analysis_callback.jl:143-147): upwind flux a>0 ? a*ul : a*ur."""
    return a * ul if a > 0 else a * ur


class AdvectionSemi1D:
    """Upstream VariableLinearAdvectionNonperiodicSemidiscretization(D, nothing, afunc, Val(false),
    left_bc, right_bc): du = -D*(a.*u) with SATs at both ends.  D exposes mul/grid/weights."""

    def __init__(self, D, afunc, left_bc, right_bc):
        self.D = D
        self.a = np.array([afunc(x) for x in D.grid], dtype=np.float64)
        self.left_bc, self.right_bc = left_bc, right_bc

    def rhs(self, u, t):
        D, a = self.D, self.a
        du = -D.mul(a * u)
        w = D.weights()
        fl = godunov_flux_variablelinearadvection(self.left_bc(t), u[0], a[0])
        fr = godunov_flux_variablelinearadvection(u[-1], self.right_bc(t), a[-1])
        du[0] += (fl - a[0] * u[0]) / w[0]
        du[-1] -= (fr - a[-1] * u[-1]) / w[-1]
        return du

    def analyze_quantities(self, du, u, t):
        """src/conservation_laws/analysis_callback.jl:136-173."""
        a, D = self.a, self.D
        P = D.weights()
        lb, rb = self.left_bc(t), self.right_bc(t)
        mass = np.sum(P * u)
        mass_rate = np.sum(P * du)
        fl = godunov_flux_variablelinearadvection(lb, u[0], a[0])
        fr = godunov_flux_variablelinearadvection(u[-1], rb, a[-1])
        mass_rate_boundary = mass_rate - fl + fr  # :148
        energy = 0.5 * np.sum(P * u**2)
        energy_rate = np.sum(P * (du * u))
        erb = energy_rate + 0.5 * np.dot(u**2, P * D.mul(a))  # :152
        if a[0] > 0.0:
            erb += 0.5 * (-a[0] * lb**2 + a[-1] * u[-1] ** 2)
        if a[-1] < 0.0:
            erb += 0.5 * (-a[0] * u[0] ** 2 + a[-1] * rb**2)
        erbd = erb
        if a[0] > 0.0:
            erbd += 0.5 * a[0] * (u[0] - lb) ** 2
        if a[-1] < 0.0:
            erbd += 0.5 * a[-1] * (u[-1] - rb) ** 2
        return np.array([mass, mass_rate, mass_rate_boundary, energy, energy_rate, erb, erbd])


# --------------------------------------------------------------------------------------
# Batched helpers: the reference has no batch axis; a batch is the per-instance call repeated.
# Layout: U is (B, N) C-contiguous == Julia N x B column-major (each instance contiguous).
# --------------------------------------------------------------------------------------


def apply_batch(Dmat, U, alpha=1.0, beta=0.0, dU=None):
    """Per-instance `mul!(dest, D, u, alpha, beta)` for every row of U, as one dgemm."""
    Y = alpha * (U @ np.asarray(Dmat).T)
    if beta != 0.0:
        Y = Y + beta * dU
    return Y


def apply_batch_longdouble(Dmat, U, alpha=1.0):
    """Extended-precision truth (x87 80-bit) used to bound rounding error of both CPU and GPU paths."""
    Dl = np.asarray(Dmat, dtype=np.longdouble)
    Ul = np.asarray(U, dtype=np.longdouble)
    return np.longdouble(alpha) * (Ul @ Dl.T)


def energy_batch(w, U):
    """integrate(abs2, u, D) = sum_i w_i u_i^2 per instance."""
    return (np.asarray(U) ** 2) @ np.asarray(w)


def ssprk53_step(rhs, u, t, dt):
    """One step of SSPRK53 (Ruuth 2006 / OrdinaryDiffEqSSPRK coefficients) - used for the 1D example
    (examples/RBF_FSBP_advection.jl:45-51).  Five-stage, third-order, Shu-Osher form."""
    a30, a32 = 0.355909775063327, 0.644090224936674
    a40, a43 = 0.367933791638137, 0.632066208361863
    a52, a54 = 0.237593836598569, 0.762406163401431
    b10 = 0.377268915331368
    b21 = 0.377268915331368
    b32 = 0.242995220537396
    b43 = 0.238458932846290
    b54 = 0.287632146308408
    c1, c2, c3, c4 = 0.377268915331368, 0.754537830662736, 0.728985661612188, 0.699226135931670
    u1 = u + b10 * dt * rhs(u, t)
    u2 = u1 + b21 * dt * rhs(u1, t + c1 * dt)
    u3 = a30 * u + a32 * u2 + b32 * dt * rhs(u2, t + c2 * dt)
    u4 = a40 * u + a43 * u3 + b43 * dt * rhs(u3, t + c3 * dt)
    u5 = a52 * u2 + a54 * u4 + b54 * dt * rhs(u4, t + c4 * dt)
    return u5


def solve_lowstorage_3sp(rhs, u0, t0, t1, dt0, tab, adaptive=True, abstol=1e-6, reltol=1e-3, pid=(0.38, 0.18, 0.01), tstops=()):
    """Low-storage 3S*+ FSAL Runge-Kutta integration with embedded error estimate and PID step size control, restating
    OrdinaryDiffEqLowStorageRK's LowStorageRK3SpFSAL perform_step + OrdinaryDiffEq's PIDController (third-party, version
    unpinned: the integrator of examples/RBF_MFSBP_advection.jl:21-27 is RDPK3SpFSAL49 of that family).  `tab`: object with
    gamma1, gamma2, gamma3, delta, c (stages 2..s), beta, bhat (stages 1..s), bhat_fsal, order, embedded_order.
    Returns (u, accepted, rejected, list of accepted times)."""
    u = np.array(u0, dtype=np.float64, copy=True)
    s = len(tab.beta)
    k_ctrl = min(tab.order, tab.embedded_order) + 1
    err = [1.0, 1.0, 1.0]
    t, dt = float(t0), float(dt0)
    stops = [x for x in tstops if t0 < x < t1]
    nxt, nacc, nrej, times = 0, 0, 0, []
    kf = rhs(u, t)
    teps = 1e-12 * max(1.0, abs(t1))
    while t < t1 - teps:
        target, h, clipped = (stops[nxt] if nxt < len(stops) else t1), dt, False
        if t + h >= target - teps:
            h, clipped = target - t, True
        uprev, tmp = u.copy(), u.copy()
        un = tmp + tab.beta[0] * h * kf
        ut = tab.bhat[0] * h * kf
        for i in range(1, s):
            k = rhs(un, t + tab.c[i - 1] * h)
            tmp = tmp + tab.delta[i - 1] * un
            un = tab.gamma1[i - 1] * un + tab.gamma2[i - 1] * tmp + tab.gamma3[i - 1] * uprev + tab.beta[i] * h * k
            ut = ut + tab.bhat[i] * h * k
        kl = rhs(un, t + h)
        accept, q = True, 1.0
        if adaptive:
            ut = ut + tab.bhat_fsal * h * kl
            r = ut / (abstol + np.maximum(np.abs(uprev), np.abs(un)) * reltol)
            eest = max(np.sqrt(np.mean(r * r)), np.finfo(float).eps)
            err[0] = 1.0 / eest
            q = err[0] ** (pid[0] / k_ctrl) * err[1] ** (pid[1] / k_ctrl) * err[2] ** (pid[2] / k_ctrl)
            q = 1.0 + np.arctan(q - 1.0)
            accept = q >= 0.81
        if accept:
            t = target if clipped else t + h
            u, kf = un, kl
            nacc += 1
            times.append(t)
            if adaptive:
                err[2], err[1] = err[1], err[0]
            dt = h * q if adaptive else dt0
            if clipped and nxt < len(stops) and target == stops[nxt]:
                nxt += 1
        else:
            nrej += 1
            dt = h * q
    return u, nacc, nrej, times


# --------------------------------------------------------------------------------------
# Construction objective: literal restatement of ext/utils.jl:26-148,164-190 and of the two objective functions
# --------------------------------------------------------------------------------------


def _set_skew_symmetric(M, sigma, k):
    """ext/utils.jl:26-37 (k: 0-based position in sigma; returns the next position)."""
    n = M.shape[0]
    for i in range(n):
        for j in range(i + 1, n):
            M[i, j] = sigma[k]
            M[j, i] = -sigma[k]
            k += 1
    return k


def _set_banded(Dm, sigma, bandwidth, init_k, different_values):
    """ext/utils.jl:40-58."""
    n = Dm.shape[0]
    k = init_k
    for i in range(n):
        for j in range(i + 1, n):
            if j - i <= bandwidth:
                if different_values:
                    l = k
                    k += 1
                else:
                    l = init_k + (j - i) - 1
                Dm[i, j] = sigma[l]
                Dm[j, i] = -sigma[l]
    return k


def _set_triangular(Cm, sigma, bandwidth, size_boundary, init_k, different_values):
    """ext/utils.jl:60-81 (1-based loops of the reference kept, indices shifted on access)."""
    n = Cm.shape[0]
    k = init_k + 1  # 1-based like the reference
    start_i = n - bandwidth + 1 if different_values else size_boundary - bandwidth + 1
    for i in range(start_i, n + 1):
        for j in range(1, i - start_i + 2):
            l = k if different_values else (init_k + 1) - 1 + bandwidth + j - (i - start_i + 1)
            Cm[i - 1, j - 1] = sigma[l - 1]
            k += 1
    return k - 1


def set_S(sigma, N, bandwidth=None, size_boundary=None, different_values=True, sparsity_pattern=None):
    """set_S! / set_S_block_banded! / set_S_sparsity_pattern! (ext/utils.jl:83-148)."""
    sigma = np.asarray(sigma, dtype=np.float64)
    b = N - 1 if bandwidth is None else bandwidth
    c = 2 * b if size_boundary is None else size_boundary
    S = np.zeros((N, N))
    if sparsity_pattern is not None:
        k = 0
        for i in range(N):
            for j in range(i + 1, N):
                if sparsity_pattern[i, j]:
                    S[i, j] = sigma[k]
                    S[j, i] = -sigma[k]
                    k += 1
        return S
    if b == N - 1:
        _set_skew_symmetric(S, sigma, 0)
        return S
    M1 = S[:c, :c]
    k = _set_skew_symmetric(M1, sigma, 0)
    M2 = S[N - c:, N - c:]
    if different_values:
        k = _set_skew_symmetric(M2, sigma, k)
    else:
        M2[:] = -M1[::-1, ::-1]
    Dm = S[c:N - c, c:N - c]
    k = _set_banded(Dm, sigma, b, k, different_values)
    C1 = S[:c, c:N - c]
    k = _set_triangular(C1, sigma, b, c, k, different_values)
    S[c:N - c, :c] = -C1.T
    C2 = S[c:N - c, N - c:]
    if different_values:
        k = _set_triangular(C2, sigma, b, c, k, different_values)
        S[N - c:, c:N - c] = -C2.T
    else:
        C1_bar = C1[::-1, ::-1]
        C2[:] = C1_bar.T
        S[N - c:, c:N - c] = -C1_bar
    return S


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))  # ext/utils.jl:153


def create_P_diag(rho, vol):
    """create_P (ext/utils.jl:164-168): diagonal of P."""
    p = _sig(np.asarray(rho, dtype=np.float64))
    return p * (vol / np.sum(p))


def objective_function_space_operator(x, L, x_length, V, V_x, R, **structure):
    """optimization_function_function_space_operator (ext/function_space_operators_optim.jl:89-105)."""
    x = np.asarray(x, dtype=np.float64)
    N = V.shape[0]
    sigma, rho = x[:L], x[L:]
    S = set_S(sigma, N, **structure)
    A = S @ V - create_P_diag(rho, x_length)[:, None] * V_x + R
    return float(np.sum(A * A))


def objective_multidimensional(x, Ls, vol, normals, moments, boundary_indices, V, V_xis, corners=None, structures=None):
    """optimization_function_multidimensional_function_space_operator
    (ext/multidimensional_function_space_operators.jl:124-161); boundary_indices 0-based, corners: per direction the (0-based)
    positions of the boundary entries that set_B! skips."""
    x = np.asarray(x, dtype=np.float64)
    d, N, Nb = len(V_xis), V.shape[0], len(normals)
    offs = np.concatenate([[0], np.cumsum(Ls)])
    rho, phi = x[len(x) - N - Nb:len(x) - Nb], x[len(x) - Nb:]
    P = create_P_diag(rho, vol)
    corners = corners if corners is not None else tuple([] for _ in range(d))
    res = 0.0
    for i in range(d):
        S = set_S(x[offs[i]:offs[i + 1]], N, **(structures[i] if structures else {}))
        b = np.zeros(N)
        for j, k in enumerate(boundary_indices):  # set_B! (ext/utils.jl:179-190)
            if j not in corners[i]:
                b[k] = phi[j] * normals[j][i]
        BV = b[:, None] * V
        A = S @ V - P[:, None] * V_xis[i] + 0.5 * BV
        Cm = V.T @ BV - moments[i]
        res += np.sum(A * A) + np.sum(Cm * Cm)
    return float(res)
