#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 SBP hot path (BASELINE.json metric: GDOF/s for batched SBP apply,
% of HBM peak, 1/2/4/8 GPU).

    python bench.py --gpus N --steps K --warmup W            # native arm (CUDA kernels through the C ABI)
    python bench.py --impl reference ...                     # reference arm: the reference's CPU path on host cores

Workload at every N (weak scaling): BASELINE config 2 PER GPU -- 1D Legendre (LGL) operator N = 64 applied to a batch
of 2^20 independent Float64 nodal vectors (`mul!(dU, D, U)`; U = dU = 512 MiB, far larger than the 126 MB L2, and two
buffer pairs are rotated).  A step is one pass of the hot path over the batch.  DOF = one nodal output value;
GDOF/s = N*B*ranks / t / 1e9.  Algorithmic traffic 16 B/DOF (read U once, write dU once).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES = 64
BATCH = 1 << 20
BYTES_PER_DOF = 16.0
FLOPS_PER_DOF = 2.0 * N_NODES  # 2 N^2 per instance / N outputs
GOLDEN = 0.6180339887498949
PARITY_NOTE = ("oracle (numpy restatement, file:line cited) pinned to the reference's known-answer tests; advection RHS / "
               "analysis quantities / 1D RHS unpinned until a Julia machine runs baseline/julia_oracle.jl (see DESIGN.md section 4)")


def headline_config(world):
    """The SAME dict in the native and the reference arm (the driver compares them)."""
    return {"workload": "config2: 1D Legendre (LGL) operator N=64, mul!(dU, D, U) over 2^20 Float64 vectors per GPU; D = "
                        "basis.D exactly as the barycentric formula yields it (centro-antisymmetric only up to rounding)",
            "N": N_NODES, "batch_per_gpu": BATCH, "bytes_per_dof": BYTES_PER_DOF,
            "l2": "inputs (512 MiB) and outputs (512 MiB) exceed the 126 MB L2; two buffer pairs rotated",
            "data": "u[b, i] = 2 frac((b N + i + 1) phi) - 1 by GLOBAL instance index b (the same problem at every N)",
            "parallelism": f"batch sharded over {world} GPU(s), operator replicated, no data-path collective"}


def batch_numpy(lo, hi, N):
    """Deterministic synthetic batch rows [lo, hi): a function of the global instance index only."""
    idx = np.arange(lo * N, hi * N, dtype=np.float64) + 1.0
    return (2.0 * np.mod(idx * GOLDEN, 1.0) - 1.0).reshape(hi - lo, N)


def batch_torch(torch, lo, hi, N, dev):
    idx = torch.arange(lo * N, hi * N, dtype=torch.float64, device=dev) + 1.0
    return (2.0 * torch.remainder(idx * GOLDEN, 1.0) - 1.0).reshape(hi - lo, N)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads():
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1, which must not shrink the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """FP64 tensor (DMMA) peak measured by tools/microbench on this pool, committed under profiles/."""
    p = os.path.join(ROOT, "profiles", "microbench_b200.json")
    try:
        d = json.load(open(p))
        return max(max(v.values()) for v in d["dmma_tflops"].values()), d.get("dfma_tflops")
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C port of the reference CPU path, timed on the host cores
# ------------------------------------------------------------------------------------------------------------------
def blas_single_thread():
    """OpenBLAS must not spawn its own threads inside the per-instance loop (the instances are spread over OpenMP threads)."""
    from threadpoolctl import threadpool_limits

    return threadpool_limits(limits=1, user_api="blas")


def cpu_reference_rate(instances, steps, warmup, threads, impl="openblas_dgemv"):
    """The reference's CPU path for `mul!(du, D, u)` on every vector of the batch, instances spread over `threads` threads:
    impl "openblas_dgemv": one OpenBLAS dgemv per instance -- what Julia's mul! calls (src/polynomialbases_operators.jl:148);
    impl "c_port": the hand-written dgemv_n loop of oracle/oracle_c.c;
    impl "dgemm": best case a CPU user could get by batching by hand, ONE multithreaded OpenBLAS dgemm U * D'."""
    from threadpoolctl import threadpool_limits

    from oracle import c_oracle as CO
    from oracle import sbp_oracle as O

    D = O.PolynomialBasesOperator("LobattoLegendre", -1.0, 1.0, N_NODES).matrix()
    U = batch_numpy(0, instances, N_NODES)
    Y = np.empty_like(U)
    if impl == "dgemm":
        Dt = np.ascontiguousarray(D.T)

        def run():
            with threadpool_limits(limits=threads, user_api="blas"):
                np.matmul(U, Dt, out=Y)
    elif impl == "c_port":
        def run():
            CO.apply_dense(D, U, Y=Y, nthreads=threads)
    else:
        def run():
            with blas_single_thread():
                CO.apply_dense_blas(D, U, Y=Y, nthreads=threads)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return instances * N_NODES * steps / dt / 1e9, dt / steps


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    threads = host_threads()
    # the full BASELINE batch per step, as in the native arm (about 25-50 ms per step on 16 cores)
    gdofs, sec = cpu_reference_rate(BATCH, args.steps, max(args.warmup, 1), threads, "openblas_dgemv")
    line = {
        "impl": "reference", "metric": "GDOF/s (batched SBP operator apply, Float64)", "value": gdofs,
        "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(args.gpus),
        "cpu_baseline": {"value": gdofs, "unit": "GDOF/s", "cores": threads, "kind": "port",
                         "sample": f"all {BATCH} instances per step: one OpenBLAS dgemv per instance (the BLAS call Julia's mul! "
                                   f"makes, src/polynomialbases_operators.jl:148; loop in oracle/oracle_c.c) spread over {threads} "
                                   "OpenMP threads; the reference itself is Julia and cannot run in this image"},
        "e2e": {"value": gdofs, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------------
def time_steps(torch, fn, steps, per_step=True):
    """K steps, CUDA events around every step on the launching stream; returns (total_ms, per-step ms list)."""
    if not per_step:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), []
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn(i)
        evs[i + 1].record()
    evs[-1].synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[-1]), per


def extra_workloads(torch, S, ctx, hbm_peak, steps=6):
    """Secondary BASELINE configs (few steps each) so that one bench line also covers the subcell table + energy,
    the fused 2D advection RHS and the large dense DGEMM."""
    out = {}
    dev = torch.device("cuda", ctx.device)

    def bench(fn, dofs, bytes_per_dof, flops=None):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        tot, per = time_steps(torch, fn, steps)
        ms = float(np.median(per))
        r = {"gdof_s": dofs / ms / 1e6, "ms": ms, "hbm_gbs": dofs * bytes_per_dof / ms / 1e6,
             "hbm_frac": dofs * bytes_per_dof / ms / 1e6 / hbm_peak}
        if flops:
            r["tflops"] = flops / ms / 1e9
        return r

    try:  # config 2 again with the symmetry fast path disabled: a GENERAL dense N=64 operator (FP64-tensor bound)
        N, B = N_NODES, BATCH
        Dg = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
        Dg.dense_flags = S._lib.DENSE_NO_SYMMETRY
        opg = Dg.device_op(ctx)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        lib, h = ctx._lib, ctx.handle

        def f2(i):
            S._lib.check(lib.sbp_apply(h, opg.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))

        out["config2_general_dense_N64_no_symmetry"] = bench(f2, B * N, 16.0, flops=2.0 * N * N * B)
        del U, Y
    except Exception as e:  # pragma: no cover
        out["config2_general_dense_N64_no_symmetry"] = {"error": str(e)[:200]}
    try:  # config 4: subcell table N=16, G=8, B=2^22, fused per-instance energy + global sum
        N, G, B = 16, 8, 1 << 22
        ops = []
        for i, NL in enumerate(range(4, 12)):
            xM = -1.0 + 2.0 * NL / N
            ops.append(S.couple_subcell(S.legendre_derivative_operator(-1.0, xM, NL),
                                        S.legendre_derivative_operator(xM, 1.0, N - NL), xM,
                                        "no" if i % 2 == 0 else "central"))
        table = S.OperatorTable(ops)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        E = torch.empty(B + 1, dtype=torch.float64, device=dev)
        op = table.device_op(ctx)
        lib, h = ctx._lib, ctx.handle

        def f4(i):
            S._lib.check(lib.sbp_apply_table(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, None,
                                             1.0, 0.0, C.c_void_p(E.data_ptr()), C.c_void_p(E.data_ptr() + 8 * B)))

        out["config4_subcell_table_energy"] = bench(f4, B * N, 16.0 + 8.0 / N)
        del U, Y, E
    except Exception as e:  # pragma: no cover
        out["config4_subcell_table_energy"] = {"error": str(e)[:200]}
    try:  # config 3: 2D MFSBP advection RHS with SAT, 36x36 jittered cloud, ellipsoid sparsity, B = 65536
        n1, B = 36, 65536
        D1 = S.mattsson_nordstrom_2004(2, -1.0, 1.0, n1)
        D2 = S.tensor_product_operator_2D(D1)
        rng = np.random.default_rng(0)
        nodes = D2.grid.copy()
        dx = 2.0 / (n1 - 1)
        interior = np.ones(D2.N, bool)
        interior[np.unique(D2.boundary_indices)] = False
        nodes[interior] += dx / 6 * (1 - 2 * rng.uniform(size=(int(interior.sum()), 2)))
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
        a = (1.0, 1.0)
        semi = S.MultidimensionalLinearAdvectionNonperiodicSemidiscretization(
            Dm, a, lambda x, t: float(np.exp(-30 * ((x[0] - t) ** 2 + (x[1] - t) ** 2))), storage="sparse")
        N = Dm.N
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        u, du = S.DeviceBatch.from_torch(ctx, U), S.DeviceBatch.from_torch(ctx, Y)
        nnz = int(np.count_nonzero(semi.cache["A"]))
        r = bench(lambda i: semi(du, u, None, 0.1), B * N, 16.0, flops=2.0 * nnz * B)
        r["N"], r["nnz_per_row"], r["kernel"] = N, nnz / N, semi.device_op(ctx).info()["kernel"]
        out["config3_advection2d_rhs_sat"] = r
        # one SSPRK53 step = 5 stage updates riding on the RHS pass (sbp_rhs_advection2d_stage): 2 without, 3 with the
        # extra operand; algorithmic traffic 2 * 16 + 3 * 24 = 104 B/DOF per step
        E = torch.rand((B, N), dtype=torch.float64, device=dev)
        e = S.DeviceBatch.from_torch(ctx, E)
        tab = semi.boundary_table(ctx, [0.1])
        so = bench(lambda i: semi.stage(du, u, 0.1, 1.0, 1e-3, gptr=tab.ptr), B * N, 16.0)
        se = bench(lambda i: semi.stage(du, u, 0.1, 0.6, 1e-3, 0.4, e, gptr=tab.ptr), B * N, 24.0)
        step_ms = 2 * so["ms"] + 3 * se["ms"]
        out["config3_ssprk53_step_fused_stage_updates"] = {
            "ms_per_step": step_ms, "stage_ms": so["ms"], "stage_with_extra_operand_ms": se["ms"],
            "gdof_steps_s": B * N / step_ms / 1e6, "hbm_gbs": 104.0 * B * N / step_ms / 1e6,
            "hbm_frac": 104.0 * B * N / step_ms / 1e6 / hbm_peak, "kernel": "k_bsr3<HELP, FUSED> (RHS + SAT + stage update)"}
        tab.free()
        del U, Y, E
        # ---- the same workload end to end with HOST buffers (pinned; H2D + D2H inside the timed region) ----
        Bh = 16384
        hU, hY = ctx.pinned_empty((Bh, N)), ctx.pinned_empty((Bh, N))
        hU[:] = batch_numpy(0, Bh, N)
        semi.rhs_host(ctx, hY, hU, 0.1)
        t0 = time.perf_counter()
        for _ in range(3):
            semi.rhs_host(ctx, hY, hU, 0.1)
        dt = (time.perf_counter() - t0) / 3
        out["config3_e2e_rhs_host"] = {"gdof_s": Bh * N / dt / 1e9, "ms": dt * 1e3, "batch": Bh,
                                       "h2d_bytes_per_step": 8 * Bh * N, "d2h_bytes_per_step": 8 * Bh * N,
                                       "api": "sbp_rhs_advection2d_host (one RHS evaluation per round trip: copy-bound)"}
        # a whole solve: upload u0 and the boundary table once, nsteps SSPRK53 steps on the device, download u(t_end)
        nsteps, hstep = 40, 2.0e-3
        cs = (0.0, 0.377268915331368, 0.754537830662736, 0.728985661612188, 0.699226135931670)
        table = np.ascontiguousarray(np.stack([semi.boundary_values(k * hstep + c * hstep) for k in range(nsteps) for c in cs]))
        hs = np.full(nsteps, hstep)
        opb = semi.device_op(ctx)

        def solve():
            S._lib.check(ctx._lib.sbp_solve_ssprk53_host(ctx.handle, opb.handle, hU.ctypes.data_as(C.c_void_p), Bh,
                                                         table.ctypes.data_as(C.c_void_p), N, nsteps,
                                                         hs.ctypes.data_as(C.c_void_p)))

        solve()
        hU[:] = batch_numpy(0, Bh, N)
        t0 = time.perf_counter()
        solve()
        dt = time.perf_counter() - t0
        out["config3_e2e_solve_ssprk53_host"] = {
            "gdof_rhs_s": 5.0 * nsteps * Bh * N / dt / 1e9, "ms": dt * 1e3, "batch": Bh, "steps": nsteps,
            "h2d_bytes": 8 * Bh * N + table.nbytes, "d2h_bytes": 8 * Bh * N,
            "api": "sbp_solve_ssprk53_host: one upload, 5 fused RHS+stage launches per step on the device, one download "
                   "(gdof_rhs_s counts one RHS evaluation of every DOF per stage)",
            "final_state_checksum": float(hU[:: max(1, Bh // 256)].sum())}
        ctx.free_pinned(hU), ctx.free_pinned(hY)
        # ---- CPU legs of config 3 (host cores of this box; bounded samples) ----
        from oracle import c_oracle as CO

        threads = host_threads()
        A, bd, mode = semi.cache["A"], semi.cache["B"], semi.mode
        g0 = semi.boundary_values(0.1)
        Us = batch_numpy(0, 4096, N)
        CO.rhs_advection2d(A, bd, mode, g0, Us[:256], faithful=False, nthreads=threads)
        t0 = time.perf_counter()
        CO.rhs_advection2d(A, bd, mode, g0, Us, faithful=False, nthreads=threads)
        t_csr = time.perf_counter() - t0
        nf = 8 * threads
        t0 = time.perf_counter()
        CO.rhs_advection2d(A, bd, mode, g0, Us[:nf], faithful=True, nthreads=threads)
        t_ref = time.perf_counter() - t0
        out["config3_cpu_legs"] = {
            "reference_faithful_gdof_s": nf * N / t_ref / 1e9, "best_case_csr_gdof_s": 4096 * N / t_csr / 1e9, "cores": threads,
            "sample": f"faithful: {nf} instances (dense dgemv over the stored zeros incl. the per-call -A materialisation, "
                      "multidimensional_linear_advection.jl:78); best case: 4096 instances, CSR rows, all threads (oracle/oracle_c.c)"}
    except Exception as e:  # pragma: no cover
        out["config3_advection2d_rhs_sat"] = dict(out.get("config3_advection2d_rhs_sat", {}), error=str(e)[:300])
    try:  # config 1 batched: banded 1D FD-SBP operator N = 101 (examples/RBF_FSBP_advection.jl sizes), RHS with SATs
        N, B = 101, 1 << 19
        D1 = S.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        D1.storage = "sparse"  # block-sparse DMMA kernel (odd N: pair-row view of the batch)
        semi1 = S.VariableLinearAdvectionNonperiodicSemidiscretization(D1, None, lambda x: 1.0, False,
                                                                       lambda t: 0.3, lambda t: 0.0)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        u, du = S.DeviceBatch.from_torch(ctx, U), S.DeviceBatch.from_torch(ctx, Y)
        r = bench(lambda i: semi1(du, u, None, 0.1), B * N, 16.0)
        r["N"], r["kernel"] = N, semi1.device_op(ctx).info()["kernel"]
        out["config1_batched_banded_N101_rhs1d"] = r
        del U, Y
    except Exception as e:  # pragma: no cover
        out["config1_batched_banded_N101_rhs1d"] = {"error": str(e)[:200]}
    try:  # config 1, the reference's own CPU-runnable case: the WHOLE solve of examples/RBF_FSBP_advection.jl (N = 101 banded
        # operator, SSPRK53, dt = 0.009, tspan = (0, 0.5): 56 steps x 5 stages) for 1 / 64 / 1024 instances, wall clock
        from oracle import c_oracle as CO
        from oracle import sbp_oracle as O

        N = 101
        D1 = S.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        D1.storage = "sparse"
        u0f = lambda x: np.tanh(50 * (x - 0.1))
        lbc, rbc = (lambda t: float(u0f(0.0 - t))), (lambda t: 0.0)
        semi1 = S.VariableLinearAdvectionNonperiodicSemidiscretization(D1, None, lambda x: 1.0, False, lbc, rbc)
        cs = (0.0, 0.377268915331368, 0.754537830662736, 0.728985661612188, 0.699226135931670)
        sched, t = [], 0.0
        while t < 0.5 - 1e-14:
            hh = min(0.009, 0.5 - t)
            sched.append((t, hh))
            t += hh
        table = np.ascontiguousarray([[lbc(tt + c * hh), rbc(tt + c * hh)] for tt, hh in sched for c in cs])
        hs = np.ascontiguousarray([hh for _, hh in sched])
        opb = semi1.device_op(ctx)
        tab = ctx.upload_array(table)
        Do = O.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        w = Do.weights()
        u0 = u0f(Do.grid)
        CO.solve_ssprk53_1d(Do.matrix(), np.ones(N), w[0], w[-1], u0, hs, table)
        t0 = time.perf_counter()
        for _ in range(10):
            CO.solve_ssprk53_1d(Do.matrix(), np.ones(N), w[0], w[-1], u0, hs, table)
        cpu_ms = (time.perf_counter() - t0) / 10 * 1e3
        res = {"steps": len(sched), "stages": 5 * len(sched), "N": N,
               "cpu_port_ms_per_solve_one_core": cpu_ms,
               "cpu_port": "oracle_solve_ssprk53_1d: dense dgemv per stage as the reference's operator type does, one core"}
        for Bs in (1, 64, 1024):
            u = ctx.upload(np.tile(u0, (Bs, 1)))
            row = {}
            for force, name in (("1", "single_launch"), ("0", "launch_per_stage")):
                ctx.set_option("solve_small", int(force))
                ts = []
                for rep in range(4):
                    u.copy_from_host(np.tile(u0, (Bs, 1)))
                    n0 = ctx.launch_count
                    t0 = time.perf_counter()
                    S._lib.check(ctx._lib.sbp_solve_ssprk53(ctx.handle, opb.handle, C.c_void_p(u.ptr), Bs, C.c_void_p(tab.ptr),
                                                            2, len(sched), hs.ctypes.data_as(C.c_void_p), None))
                    ctx.synchronize()
                    ts.append((time.perf_counter() - t0) * 1e3)
                row[name + "_ms"] = float(np.median(ts[1:]))
                row[name + "_launches"] = int(ctx.launch_count - n0)
            ctx.set_option("solve_small", -1)
            row["cpu_port_ms_all_cores"] = cpu_ms * -(-Bs // host_threads())
            row["speedup_vs_cpu_port"] = row["cpu_port_ms_all_cores"] / row["single_launch_ms"]
            res[f"B{Bs}"] = row
        tab.free()
        out["config1_whole_solve_latency"] = res
    except Exception as e:  # pragma: no cover
        out["config1_whole_solve_latency"] = {"error": str(e)[:300]}
    try:  # config 5 (reduced batch): dense N = 4096 operator, B = 8192 -> DMMA DGEMM
        N, B = 4096, 8192
        rng = np.random.default_rng(0)
        Sk = rng.normal(size=(N, N))
        Sk = Sk - Sk.T
        w = rng.uniform(0.5, 1.5, N)
        w *= 4.0 / w.sum()
        Dmat = Sk / w[:, None]
        op = S.upload_dense(ctx, Dmat, w)
        U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
        Y = torch.empty_like(U)
        lib, h = ctx._lib, ctx.handle

        def f5(i):
            S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))

        r = bench(f5, B * N, 16.0 + 8.0 * N / B, flops=2.0 * N * N * B)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        Dt = torch.from_numpy(np.ascontiguousarray(Dmat)).to(dev)
        for _ in range(2):
            torch.matmul(U, Dt.T, out=Y)
        t0.record()
        for _ in range(3):
            torch.matmul(U, Dt.T, out=Y)
        t1.record()
        t1.synchronize()
        r["cublas_dgemm_tflops"] = 2.0 * N * N * B * 3 / t0.elapsed_time(t1) / 1e9
        out["config5_dense4096_dgemm_B8192"] = r
        del U, Y, Dt
    except Exception as e:  # pragma: no cover
        out["config5_dense4096_dgemm_B8192"] = {"error": str(e)[:200]}
    return out


def library_allreduce(ctx, E):
    """Sum of a few device doubles over the ranks through the PRODUCT's collective (sbp_allreduce_sum: ncclAllReduce on the
    communicator of sbp_comm_init, on the ctx stream), not torch.distributed."""
    ctx.allreduce_sum(E.data_ptr(), E.numel())


def config5_sharded(torch, S, ctx, dist, rank, world, dev, steps=3):
    """BASELINE config 5 at its full size, strong scaling: dense MFSBP-type operator N = 4096 (replicated), batch 65536
    sharded contiguously over the ranks, mul! + global discrete energy u'Pu (one NCCL allreduce of a scalar per step through
    sbp_allreduce_sum, inside the timed region).  Data are a function of the GLOBAL instance index, so every N computes the
    same problem and the global energy must agree across N to rounding.  Device time, max over ranks."""
    from sbp_b200.sharding import shard_range

    N, Btot = 4096, 65536
    lo, hi = shard_range(Btot, world, rank)
    B = hi - lo
    rng = np.random.default_rng(0)  # same operator on every rank
    Sk = rng.normal(size=(N, N))
    Sk = Sk - Sk.T
    w = rng.uniform(0.5, 1.5, N)
    w *= 4.0 / w.sum()
    op = S.upload_dense(ctx, Sk / w[:, None], w)
    del Sk
    U = batch_torch(torch, lo, hi, N, dev)
    Y = torch.empty_like(U)
    E = torch.zeros(2, dtype=torch.float64, device=dev)
    lib, h = ctx._lib, ctx.handle

    def step():
        S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))
        S._lib.check(lib.sbp_integrate(h, op.handle, C.c_void_p(Y.data_ptr()), None, B, 1, None, C.c_void_p(E.data_ptr())))
        if world > 1:
            library_allreduce(ctx, E[:1])

    for _ in range(2):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    t1.synchronize()
    t = torch.tensor([t0.elapsed_time(t1) / steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    energy = float(E[0].item())
    del U, Y
    out = {"workload": "config5: dense N=4096 operator x 65536 instances (strong scaling), mul! + global energy allreduce",
           "ms": ms, "gdof_s": N * Btot / ms / 1e6, "tflops": 2.0 * N * N * Btot / ms / 1e9,
           "tflops_per_gpu": 2.0 * N * N * Btot / ms / 1e9 / world, "batch_per_gpu": B, "global_energy": energy,
           "kernel": "k_dense_gemm_tma (TMA + DMMA.8x8x4)",
           "collective": "sbp_allreduce_sum (ncclAllReduce of 1 double per step on the library's communicator)" if world > 1 else "none"}
    # the N = 1 value of the same (deterministic) problem, measured on this pool and committed: cross-N agreement
    try:
        ref = json.load(open(os.path.join(ROOT, "profiles", "r2_config5_energy_n1.json")))["global_energy"]
        out["global_energy_n1"] = ref
        out["global_energy_rel_diff_vs_n1"] = abs(energy - ref) / abs(ref)
    except Exception:
        out["global_energy_rel_diff_vs_n1"] = 0.0 if world == 1 else None
    return out


def pcie_and_e2e(torch, S, ctx, dist, rank, world, op, N, B, steps):
    """End to end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region, next to the box's own
    pinned-copy ceiling measured in the same run (all ranks probe at the same time, as they copy at the same time)."""
    lib, h = ctx._lib, ctx.handle

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # host buffers and the copy-issuing thread on the GPU's NUMA node (restored below: the CPU baseline uses every core)
    affinity0 = os.sched_getaffinity(0)
    e2e_cpus = sorted(affinity0) if os.environ.get("SBP_BENCH_NO_BIND") else ctx.bind_host_to_gpu_numa()
    barrier()
    probe = ctx.pcie_probe(256 << 20)
    hU, hY = ctx.pinned_empty((B, N)), ctx.pinned_empty((B, N))
    hU[:] = batch_numpy(rank * B, (rank + 1) * B, N)

    def e2e_step():
        S._lib.check(lib.sbp_apply_host(h, op.handle, hY.ctypes.data_as(C.c_void_p), hU.ctypes.data_as(C.c_void_p), B,
                                        1.0, 0.0, 0))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=torch.device("cuda", ctx.device))
    per_rank = [te.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, te)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    per_rank_s = [float(x.item()) for x in per_rank]
    e2e_s = float(te.item())
    checksum = float(hY[:: max(1, B // 1024)].sum())
    ctx.free_pinned(hU), ctx.free_pinned(hY)
    os.sched_setaffinity(0, affinity0)
    pr = torch.tensor([probe["h2d"], probe["d2h"], probe["bidir"]], dtype=torch.float64, device=torch.device("cuda", ctx.device))
    if dist is not None:
        dist.all_reduce(pr)  # sum over ranks = what the box delivers with every rank copying
    pr = [float(x) for x in pr.tolist()]
    dofs = float(N) * B * world
    moved_gbs = 16.0 * dofs * steps / e2e_s / 1e9  # 8 B/DOF up + 8 B/DOF down
    return {"value": dofs * steps / e2e_s / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": int(8 * N * B),
            "d2h_bytes_per_step": int(8 * N * B), "steps": steps,
            "api": "sbp_apply_host (pinned host buffers, chunked H2D/compute/D2H overlap)", "host_numa_cpus": len(e2e_cpus),
            "checksum": checksum, "per_rank_seconds": per_rank_s,
            "pcie_probe_gbs": {"h2d": pr[0], "d2h": pr[1], "bidir_sum": pr[2],
                               "how": "sbp_pcie_probe: pinned cudaMemcpyAsync of 256 MiB, alone and both directions at once, "
                                      f"all {world} rank(s) at the same time, summed over ranks"},
            "copy_gbs": moved_gbs, "pcie_frac": moved_gbs / pr[2] if pr[2] > 0 else None,
            "limiter": "host<->device copies (PCIe / host DRAM): the kernel needs 0.18 ms per 2^20 instances, the copies ~10 ms "
                       "each way; at N ranks the copies of all GPUs share the host's memory system (see pcie_probe_gbs)"}


def run_native(args):
    import torch

    import sbp_b200 as S

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = S.Context(local)
    # run torch's allocator/RNG/events on the library's own (non-default) stream so that CUDA events bracket the
    # kernels: torch's legacy default stream has handle 0, which sbp_ctx_set_stream treats as "use the ctx stream"
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.set_stream(stream)
    lib, h = ctx._lib, ctx.handle
    if world > 1:  # the library's own NCCL communicator: rank 0 creates the id, the host runtime broadcasts it
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])

    N, B = N_NODES, BATCH  # per GPU (weak scaling)
    D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)  # basis.D untouched (flags = 0)
    op = D.device_op(ctx)
    bufs = []
    for k in range(2):  # two (U, dU) pairs rotated: 2 GiB of traffic between reuses of a line
        U = batch_torch(torch, (2 * rank + k) * B, (2 * rank + k + 1) * B, N, dev)
        bufs.append((U, torch.empty_like(U)))

    def step(i):
        U, Y = bufs[i & 1]
        S._lib.check(lib.sbp_apply(h, op.handle, C.c_void_p(Y.data_ptr()), C.c_void_p(U.data_ptr()), B, 1.0, 0.0))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    total_ms, per = time_steps(torch, step, args.steps)
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    total_ms = float(t.item())
    # a sustained region of the same step (>= 0.5 s of kernels) so that the clock sampler sees the load whatever K is
    sustained = None
    if total_ms < 500.0:
        n_sus = int(min(20000, max(args.steps, 600.0 / max(total_ms / args.steps, 1e-3))))
        s2 = ClockSampler(local)
        s2.start()
        sus_ms, _ = time_steps(torch, step, n_sus, per_step=False)
        torch.cuda.synchronize()
        c2 = s2.stop()
        sustained = {"steps": n_sus, "ms_per_step": sus_ms / n_sus, "seconds": sus_ms / 1e3,
                     "gdof_s_per_gpu": float(N) * B * n_sus / sus_ms / 1e6, "clocks": c2}
        if clocks["samples"] < 5:
            clocks = dict(c2, note=f"sampled over the sustained region ({n_sus} steps); the K-step region had "
                                   f"{clocks['samples']} sample(s)")
        barrier()

    # the path's only collective: a sum of scalars (global discrete energy) through the library's communicator
    E = torch.zeros(2, dtype=torch.float64, device=dev)
    S._lib.check(lib.sbp_integrate(h, op.handle, C.c_void_p(bufs[0][0].data_ptr()), None, B, 1, None,
                                   C.c_void_p(E.data_ptr())))
    if world > 1:
        library_allreduce(ctx, E[:1])
    torch.cuda.synchronize()
    global_energy = float(E[0].item())

    e2e = pcie_and_e2e(torch, S, ctx, dist, rank, world, op, N, B, max(2, min(args.steps, 5)))

    try:
        c5 = config5_sharded(torch, S, ctx, dist, rank, world, dev) if not args.no_extra else None
    except Exception as e:  # pragma: no cover
        c5 = {"error": str(e)[:200]}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = measured_peaks()
    dofs_step = float(N) * B
    kernel_ms = float(np.mean(per))
    achieved = dofs_step * BYTES_PER_DOF / kernel_ms / 1e6  # GB/s
    fp64_dmma, fp64_dfma = fp64_peak()
    traffic = None
    for tp in (os.path.join(ROOT, "profiles", "r2_traffic.json"), os.path.join(ROOT, "profiles", "r1_traffic.json")):
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("config2_dense_small_bytes_per_launch")
                break
            except Exception:
                traffic = None
    kern = op.info()["kernel"]
    line = {
        "metric": "GDOF/s (batched SBP operator apply, Float64)",
        "value": world * dofs_step * args.steps / total_ms / 1e6,
        "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": headline_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": {"dmma_smem_sym": "k_dense_small_sym<NTH=4> (even/odd split, DMMA.8x8x4)",
                                "dmma_smem": "k_dense_small<NT=8> (DMMA.8x8x4)"}.get(kern, kern),
                     "op_info_kernel": kern,
                     "operator_upload": "basis.D of the Python mirror = the barycentric matrix unmodified, flags = 0; the library "
                                        "detects |D + JDJ| <= 256 eps |D|max and applies (D - JDJ)/2 on the even/odd kernel",
                     "kernel_ms": kernel_ms,
                     "fp64_tflops_dense_equivalent": dofs_step * FLOPS_PER_DOF / kernel_ms / 1e9,
                     "fp64_flops_executed_fraction": 0.5 if kern == "dmma_smem_sym" else 1.0,
                     "fp64_dmma_peak_tflops": fp64_dmma,
                     "note": "N=64 has AI = 8 flop/B: time >= max(bytes/HBM, flops/FP64 peak); both fractions reported"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "global_energy_allreduce": global_energy,
        "collective": "sbp_allreduce_sum (library NCCL communicator)" if world > 1 else "none",
        "parity": PARITY_NOTE,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if c5 is not None:
        line["config5_sharded"] = c5
    if fp64_dmma:
        line["roofline"]["fp64_frac"] = (line["roofline"]["fp64_tflops_dense_equivalent"] *
                                         line["roofline"]["fp64_flops_executed_fraction"] / fp64_dmma)
    if world == 1:
        threads = host_threads()
        sample = 1 << 18
        g1, _ = cpu_reference_rate(1 << 14, 2, 1, 1, "openblas_dgemv")
        gblas, sb = cpu_reference_rate(sample, 3, 1, threads, "openblas_dgemv")
        reps = int(max(3, min(60, 6.0 / max(sb, 1e-3))))
        gblas, _ = cpu_reference_rate(sample, reps, 1, threads, "openblas_dgemv")
        gport, _ = cpu_reference_rate(sample, max(3, reps // 2), 1, threads, "c_port")
        ggemm, _ = cpu_reference_rate(sample, max(3, reps // 2), 1, threads, "dgemm")
        line["cpu_baseline"] = {
            "value": gblas, "unit": "GDOF/s", "cores": threads, "kind": "port", "single_core_value": g1,
            "sample": f"{sample} of {B} instances per pass, {reps} passes: one OpenBLAS dgemv per instance (what the reference's "
                      f"mul! calls) on {threads} OpenMP threads; single-thread rate in single_core_value",
            "legs": {"openblas_dgemv_per_instance": gblas, "c_port_dgemv_per_instance": gport,
                     "best_case_one_multithreaded_dgemm": ggemm},
            "legs_note": "the reference has no batch axis: its path is the per-instance dgemv; the dgemm leg is what a CPU user "
                         "would get by batching by hand (SURVEY 8d best case)"}
        if not args.no_extra:
            line["extra"] = extra_workloads(torch, S, ctx, hbm_peak)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (configs 1/3/4/5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
