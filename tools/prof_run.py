#!/usr/bin/env python
"""Short driver for ncu captures: runs a few steps of one BASELINE workload through the C ABI and exits.

    python tools/prof_run.py --workload config2 --steps 3 [--batch-log2 20]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch-log2", type=int, default=None)
    args = ap.parse_args()
    ctx = S.Context(0)
    lib, h = ctx._lib, ctx.handle
    rng = np.random.default_rng(0)
    if args.workload == "config2":
        N, B = 64, 1 << (args.batch_log2 or 20)
        D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
        op = D.device_op(ctx)
        u = ctx.upload(rng.uniform(-1, 1, (B, N)))
        du = u.similar()
        for _ in range(args.steps):
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(du.ptr), C.c_void_p(u.ptr), B, 1.0, 0.0))
    elif args.workload == "config4":
        N, B = 16, 1 << (args.batch_log2 or 22)
        ops = []
        for i, NL in enumerate(range(4, 12)):
            xM = -1.0 + 2.0 * NL / N
            ops.append(S.couple_subcell(S.legendre_derivative_operator(-1.0, xM, NL),
                                        S.legendre_derivative_operator(xM, 1.0, N - NL), xM,
                                        "no" if i % 2 == 0 else "central"))
        table = S.OperatorTable(ops)
        u = ctx.upload(rng.uniform(-1, 1, (B, N)))
        du = u.similar()
        e = S.RawBuffer(ctx, 8 * B + 16)
        op = table.device_op(ctx)
        for _ in range(args.steps):
            S._lib.check(lib.sbp_apply_table(h, op.handle, C.c_void_p(du.ptr), C.c_void_p(u.ptr), B, None, 1.0, 0.0,
                                             C.c_void_p(e.ptr), C.c_void_p(e.ptr + 8 * B)))
    elif args.workload in ("config3", "config3stage"):
        n1, B = 36, 1 << (args.batch_log2 or 16)
        D2 = S.tensor_product_operator_2D(S.mattsson_nordstrom_2004(2, -1.0, 1.0, n1))
        nodes = D2.grid.copy()
        interior = np.ones(D2.N, bool)
        interior[np.unique(D2.boundary_indices)] = False
        nodes[interior] += (2.0 / (n1 - 1)) / 6 * (1 - 2 * rng.uniform(size=(int(interior.sum()), 2)))
        Ds = []
        for dim, lengths in ((0, (0.14, 0.08)), (1, (0.08, 0.14))):
            pat = S.neighborhood_sparsity_pattern(nodes, lengths)
            pat = pat | pat.T | np.eye(D2.N, dtype=bool)
            Dd = D2.Ds[dim].copy()
            extra = pat & (Dd == 0)
            Dd[extra] = 0.05 * rng.normal(size=int(extra.sum()))
            Ds.append(Dd)
        Dm = S.MultidimensionalMatrixDerivativeOperator(nodes, D2.boundary_indices, D2.normals, D2.weights(),
                                                        D2.weights_boundary, Ds, 2, corners=D2.corners)
        semi = S.MultidimensionalLinearAdvectionNonperiodicSemidiscretization(
            Dm, (1.0, 1.0), lambda x, t: float(np.exp(-30 * ((x[0] - t) ** 2 + (x[1] - t) ** 2))), storage="sparse")
        u = ctx.upload(rng.uniform(-1, 1, (B, Dm.N)))
        du = u.similar()
        if args.workload == "config3stage":  # Runge-Kutta stage update riding on the RHS pass, with the extra operand
            e = u.copy()
            for _ in range(args.steps):
                semi.stage(du, u, 0.1, 0.6, 1e-3, 0.4, e)
        else:
            for _ in range(args.steps):
                semi(du, u, None, 0.1)
    elif args.workload == "config5":
        N, B = 4096, 1 << (args.batch_log2 or 12)
        Sk = rng.normal(size=(N, N))
        op = S.upload_dense(ctx, Sk - Sk.T, np.ones(N))
        u = ctx.upload(rng.uniform(-1, 1, (B, N)))
        du = u.similar()
        for _ in range(args.steps):
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(du.ptr), C.c_void_p(u.ptr), B, 1.0, 0.0))
    else:
        raise SystemExit("unknown workload")
    ctx.synchronize()
    print("done", args.workload, ctx.launch_count)


if __name__ == "__main__":
    main()
