#!/usr/bin/env python
"""Time one workload (median of CUDA-event-timed launches) -- used for kernel-variant sweeps on the GPU box.

    SBP_SYM_VARIANT=3 python tools/time_config.py --workload config2 --steps 40 [--nosym]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--nosym", action="store_true")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--batch-log2", type=int, default=20)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    ctx = S.Context(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=dev))
    lib, h = ctx._lib, ctx.handle
    N, B = args.n, 1 << args.batch_log2
    if args.workload == "config2":
        D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
        if args.nosym:
            D.dense_flags = S._lib.DENSE_NO_SYMMETRY
        op = D.device_op(ctx)
        bufs = []
        for _ in range(2):
            U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
            bufs.append((U, torch.empty_like(U)))

        def step(i):
            U, Y = bufs[i & 1]
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))
    elif args.workload == "config4":
        N, B = 16, 1 << 22
        ops = []
        for i, NL in enumerate(range(4, 12)):
            xM = -1.0 + 2.0 * NL / N
            ops.append(S.couple_subcell(S.legendre_derivative_operator(-1.0, xM, NL),
                                        S.legendre_derivative_operator(xM, 1.0, N - NL), xM,
                                        "no" if i % 2 == 0 else "central"))
        op = S.OperatorTable(ops).device_op(ctx)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        E = torch.empty(B + 1, dtype=torch.float64, device=dev)

        def step(i):
            S._lib.check(lib.sbp_apply_table(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, None,
                                             1.0, 0.0, C.c_void_p(E.data_ptr()), C.c_void_p(E.data_ptr() + 8 * B)))
    elif args.workload == "banded":  # plain mul! with a banded FD-SBP operator (no SAT epilogue)
        N, B = (args.n if args.n != 64 else 101), 1 << 19
        D1 = S.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        D1.storage = "dense" if os.environ.get("SBP_BENCH_DENSE") else "sparse"
        op = D1.device_op(ctx)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)

        def step(i):
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))
    elif args.workload == "config5":
        N, B = 4096, 1 << (args.batch_log2 if args.batch_log2 != 20 else 13)
        rng = np.random.default_rng(0)
        Sk = rng.normal(size=(N, N))
        op = S.upload_dense(ctx, Sk - Sk.T, np.ones(N))
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)

        def step(i):
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))
    elif args.workload in ("config3", "config1b", "config3apply"):
        if args.workload in ("config3", "config3apply"):  # 2D MFSBP advection RHS with SAT on a jittered 36x36 cloud
            n1 = args.n if args.n != 64 else 36  # --n: grid points per direction (N = n1^2), batch scaled to ~85 M DOF
            B = 65536 if n1 == 36 else (65536 * 1296 // (n1 * n1)) // 64 * 64
            rng = np.random.default_rng(0)
            D2 = S.tensor_product_operator_2D(S.mattsson_nordstrom_2004(2, -1.0, 1.0, n1))
            nodes = D2.grid.copy()
            interior = np.ones(D2.N, bool)
            interior[np.unique(D2.boundary_indices)] = False
            if not os.environ.get("SBP_BENCH_REGULAR"):  # regular grid => identical stencil offsets on every row
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
                Dm, (1.0, 1.0), lambda x, t: float(np.exp(-30 * ((x[0] - t) ** 2 + (x[1] - t) ** 2))),
                storage="sparse")
            N = Dm.N
        else:  # config 1 batched: banded 1D operator N = 101, SAT at both ends
            N, B = (args.n if args.n != 64 else 101), 1 << 19
            D1 = S.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
            D1.storage = "dense" if os.environ.get("SBP_BENCH_DENSE") else "sparse"
            if args.nosym:
                D1.dense_flags = S._lib.DENSE_NO_SYMMETRY
            semi = S.VariableLinearAdvectionNonperiodicSemidiscretization(D1, None, lambda x: 1.0, False,
                                                                          lambda t: 0.3, lambda t: 0.0)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        u, du = S.DeviceBatch.from_torch(ctx, U), S.DeviceBatch.from_torch(ctx, Y)
        semi(du, u, None, 0.1)
        op = semi.device_op(ctx)

        def step(i):
            semi(du, u, None, 0.1)

        if args.workload == "config3apply":  # the same operator without the SAT terms: mul!(du, A, u, -1)
            inner = semi._dev[id(ctx)][1]

            def step(i):
                S._lib.check(lib.sbp_apply(h, inner.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, -1.0, 0.0))
    else:
        raise SystemExit("unknown workload")
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)])
    ms = float(np.median(per))
    print(json.dumps({"tag": args.tag, "workload": args.workload, "N": N, "B": B, "kernel": op.info()["kernel"],
                      "variant": os.environ.get("SBP_SYM_VARIANT"), "ms_median": ms, "ms_min": float(per.min()),
                      "gdof_s": N * B / ms / 1e6, "hbm_gbs": 16.0 * N * B / ms / 1e6}))


if __name__ == "__main__":
    main()
