#!/usr/bin/env python
"""Why does the host-buffer path not scale with the number of ranks?  Run under torchrun with 1/2/4/8 ranks:
every rank, at the same time, (a) probes its pinned-copy rates, (b) runs sbp_apply_host with several chunk sizes,
(c) repeats (b) with host buffers of a quarter of the size (256 MiB total instead of 1 GiB per rank).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_probe.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = S.Context(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    out = {"ranks": world, "cpus": len(os.sched_getaffinity(0))}
    barrier()
    p = ctx.pcie_probe(256 << 20)
    out["probe_sum_gbs"] = dict(zip(("h2d", "d2h", "bidir"), gather([p["h2d"], p["d2h"], p["bidir"]])))
    N = 64
    D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
    op = D.device_op(ctx)
    for B in (1 << 20, 1 << 18):
        hU, hY = ctx.pinned_empty((B, N)), ctx.pinned_empty((B, N))
        hU[:] = 0.5
        for chunk_mib in (8, 32, 128):
            chunk = chunk_mib * (1 << 20) // (N * 8)
            call = lambda: S._lib.check(ctx._lib.sbp_apply_host(ctx.handle, op.handle, hY.ctypes.data_as(C.c_void_p),
                                                                hU.ctypes.data_as(C.c_void_p), B, 1.0, 0.0, chunk))
            call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                call()
            dt = (time.perf_counter() - t0) / 4
            tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            out[f"apply_host_B{B}_chunk{chunk_mib}MiB_sum_gbs_both_directions"] = world * 16.0 * N * B / float(tmax.item()) / 1e9
        ctx.free_pinned(hU), ctx.free_pinned(hY)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
