#!/usr/bin/env python
"""Time the host-buffer path (sbp_apply_host: pinned H2D + kernel + D2H pipeline) for several chunk sizes."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    ctx = S.Context(0)
    N, B = 64, 1 << 20
    D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
    op = D.device_op(ctx)
    hU, hY = ctx.pinned_empty((B, N)), ctx.pinned_empty((B, N))
    hU[:] = np.random.default_rng(0).uniform(-1, 1, (B, N))
    for chunk_mib in [4, 8, 16, 32, 64, 128, 256]:
        chunk = chunk_mib * (1 << 20) // (N * 8)
        for _ in range(2):
            S._lib.check(ctx._lib.sbp_apply_host(ctx.handle, op.handle, hY.ctypes.data_as(C.c_void_p),
                                                 hU.ctypes.data_as(C.c_void_p), B, 1.0, 0.0, chunk))
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            S._lib.check(ctx._lib.sbp_apply_host(ctx.handle, op.handle, hY.ctypes.data_as(C.c_void_p),
                                                 hU.ctypes.data_as(C.c_void_p), B, 1.0, 0.0, chunk))
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"chunk_mib": chunk_mib, "ms": dt * 1e3, "gdof_s": N * B / dt / 1e9,
                          "gb_s_each_direction": N * B * 8 / dt / 1e9}))


if __name__ == "__main__":
    main()
