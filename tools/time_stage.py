#!/usr/bin/env python
"""CUDA-event timing of one RHS pass, one fused stage update (with / without the extra operand) and one separate update pass.

    python tools/time_stage.py --workload config3|config1b
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--steps", type=int, default=30)
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
    mk = lambda: S.DeviceBatch.from_torch(ctx, torch.rand((B, N), dtype=torch.float64, device=dev) * 2 - 1)
    u, e, out, du = mk(), mk(), mk(), mk()
    tab = semi.boundary_table(ctx, [0.1])
    cases = {
        "rhs": lambda: semi(du, u, None, 0.1),
        "stage_own": lambda: semi.stage(out, u, 0.1, 1.0, 1e-3, gptr=tab.ptr),
        "stage_extra": lambda: semi.stage(out, u, 0.1, 0.6, 1e-3, 0.4, e, gptr=tab.ptr),
        "lincomb3": lambda: out.lincomb_(0.6, u, 0.4, e, 1e-3, du),
        "lincomb2": lambda: out.lincomb_(1.0, u, 1e-3, du),
    }
    res = {"workload": args.workload, "N": N, "B": B}
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        res[name + "_ms"] = float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]))
    r = res
    res["ssprk53_step_fused_ms"] = 2 * r["stage_own_ms"] + 3 * r["stage_extra_ms"]
    res["ssprk53_step_separate_ms"] = 5 * r["rhs_ms"] + 2 * r["lincomb2_ms"] + 3 * r["lincomb3_ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
