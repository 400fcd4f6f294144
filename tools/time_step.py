#!/usr/bin/env python
"""Time SSPRK53 steps (device resident) with the stage updates riding on the RHS pass vs. separate RHS + update passes.

    SBP_FUSED_STAGE=0|1 python tools/time_step.py --workload config3|config1b [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--separate", action="store_true", help="RHS + lincomb_ passes (the driver's non-fused route)")
    args = ap.parse_args()
    ctx = S.Context(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=dev))
    if args.workload == "config3":
        n1, B = 36, 65536
        D2 = S.tensor_product_operator_2D(S.mattsson_nordstrom_2004(2, -1.0, 1.0, n1))
        semi = S.MultidimensionalLinearAdvectionNonperiodicSemidiscretization(
            D2, (1.0, 1.0), lambda x, t: np.exp(-30 * ((x[0] - t) ** 2 + (x[1] - t) ** 2)), storage="sparse")
        N = D2.N
    else:
        N, B = 101, 1 << 19
        D1 = S.mattsson_nordstrom_2004(4, 0.0, 1.0, N)
        D1.storage = "sparse"
        semi = S.VariableLinearAdvectionNonperiodicSemidiscretization(D1, None, lambda x: 1.0, False,
                                                                      lambda t: 0.3, lambda t: 0.0)
    U = torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1
    u0 = S.DeviceBatch.from_torch(ctx, U)
    dt = 1e-3
    f = semi
    if args.separate:  # hide the stage method: the driver takes the RHS + lincomb_ route
        class _Plain:
            def __init__(self, s):
                self._s = s

            def __call__(self, *a, **k):
                return self._s(*a, **k)
        f = _Plain(semi)
    prob = S.ODEProblem(f, u0, (0.0, 2 * dt))
    S.solve_ssprk53(prob, dt)  # warm-up (2 steps)
    torch.cuda.synchronize()
    prob = S.ODEProblem(f, u0, (0.0, args.steps * dt))
    t0 = time.perf_counter()
    S.solve_ssprk53(prob, dt)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({"workload": args.workload, "N": N, "B": B, "separate": args.separate,
                      "fused_env": os.environ.get("SBP_FUSED_STAGE"), "ms_per_step": ms,
                      "gdof_steps_s": N * B / ms / 1e6, "rhs_evals_gdof_s": 5 * N * B / ms / 1e6}))


if __name__ == "__main__":
    main()
