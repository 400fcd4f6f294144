#!/usr/bin/env python
"""Generates tests/golden/hotpath_golden.npz: small seeded input/output vectors of every hot-path entry point.

The reference is Julia and cannot run in this image, so the outputs come from the CPU oracle (oracle/sbp_oracle.py),
which is pinned to the reference's own known-answer tests by tests/test_oracle.py (incl. the literal numbers of
tests/golden/reference_known_answers.json).  The fixture freezes those outputs: the oracle (CPU suite) and the CUDA
path (GPU suite) are both compared against the stored arrays, so neither can drift silently.

    python tests/golden/make_golden.py        # rewrites hotpath_golden.npz and hotpath_inputs.bin (deterministic)

`hotpath_inputs.bin` holds the seeded INPUT vectors in a flat container (tests/golden/golden_io.py) that
`baseline/julia_oracle.jl` reads on a Julia machine to produce `hotpath_golden_julia.bin`: the same outputs from the real
reference (`mul!`, `semi(du, u, p, t)`, `analyze_quantities`).  tests/test_julia_golden.py compares the oracle and the CUDA
path with that file when it is present (until then those tests are skipped and parity stays "unpinned" for the RHS rows).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sbp_oracle as O  # noqa: E402


def build():
    rng = np.random.default_rng(20240229)
    out = {}
    # (1) mul! with a PolynomialBasesDerivativeOperator (src/polynomialbases_operators.jl:139-149), alpha/beta form
    D = O.PolynomialBasesOperator("LobattoLegendre", -2.0, 1.4, 16)
    U, Y0 = rng.uniform(-1, 1, (5, 16)), rng.uniform(-1, 1, (5, 16))
    out["lgl16_U"], out["lgl16_Y0"] = U, Y0
    out["lgl16_dU"] = np.stack([D.mul(u) for u in U])
    out["lgl16_dU_axpby"] = np.stack([D.mul(u, -1.5, 0.25, y.copy()) for u, y in zip(U, Y0)])
    out["lgl16_energy"] = np.array([D.integrate(np.square, u) for u in U])
    # (2) banded FD-SBP operator stored dense in the reference (upstream MatrixDerivativeOperator), N = 20 and odd N = 21
    for N in (20, 21):
        Dm = O.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        U = rng.uniform(-1, 1, (4, N))
        out[f"mn4_{N}_U"] = U
        out[f"mn4_{N}_dU"] = np.stack([Dm.mul(u) for u in U])
    # (3) SubcellOperator family (src/subcell_operators.jl:270-280) with fused energy (:92-94)
    N = 16
    U = rng.uniform(-1, 1, (8, N))
    dU, E = [], []
    for i, NL in enumerate(range(4, 12)):
        xM = -1.0 + 2.0 * NL / N
        Ds = O.couple_subcell(O.legendre_derivative_operator(-1.0, xM, NL), O.legendre_derivative_operator(xM, 1.0, N - NL),
                              xM, "no" if i % 2 == 0 else "central")
        dU.append(Ds.mul(U[i]))
        E.append(Ds.integrate(np.square, U[i]))
    out["subcell_U"], out["subcell_dU"], out["subcell_energy"] = U, np.stack(dU), np.array(E)
    # (4) 2D advection RHS + analysis quantities (src/conservation_laws/multidimensional_linear_advection.jl:60-108)
    D2 = O.tensor_product_operator_2D(O.mattsson_nordstrom_2004(2, -1.0, 1.0, 8))
    a, t = (1.0, -0.5), 0.37
    g = lambda x, tt: np.exp(-30 * ((x[0] - a[0] * tt) ** 2 + (x[1] - a[1] * tt) ** 2)) + 0.1 * np.sin(3 * x[0] - x[1] + tt)
    semi = O.AdvectionSemi2D(D2, a, g, O.tensor_product_corners(8, 8))
    U = rng.uniform(-1, 1, (3, 64))
    rhs, q = [], []
    for u in U:  # analyze_quantities reads the tmp1 left by the set_bc! of the SAME instance's RHS call
        rhs.append(semi.rhs(u, t))
        q.append(semi.analyze_quantities(rhs[-1], u))
    out["adv2d_U"], out["adv2d_dU"], out["adv2d_quantities"] = U, np.stack(rhs), np.stack(q)
    # (5) 1D advection RHS with Godunov SATs (upstream semidiscretization; examples/RBF_FSBP_advection.jl:37-38)
    D1 = O.mattsson_nordstrom_2004(4, 0.0, 1.0, 21)
    semi1 = O.AdvectionSemi1D(D1, lambda x: 1.0, lambda tt: np.tanh(50 * (-tt - 0.1)), lambda tt: 0.3 * np.cos(tt))
    U = rng.uniform(-1, 1, (3, 21))
    rhs1 = np.stack([semi1.rhs(u, 0.2) for u in U])
    out["adv1d_U"], out["adv1d_dU"] = U, rhs1
    out["adv1d_quantities"] = np.stack([semi1.analyze_quantities(d, u, 0.2) for d, u in zip(rhs1, U)])
    return out


INPUT_KEYS = ("lgl16_U", "lgl16_Y0", "mn4_20_U", "mn4_21_U", "subcell_U", "adv2d_U", "adv1d_U")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "hotpath_golden.npz")
    arrays = build()
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")
    # the INPUT vectors in the flat container that baseline/julia_oracle.jl reads (numpy's RNG does not exist in Julia)
    sys.path.insert(0, here)
    from golden_io import write_bin

    write_bin(os.path.join(here, "hotpath_inputs.bin"), {k: arrays[k] for k in INPUT_KEYS})
    print("wrote hotpath_inputs.bin")
