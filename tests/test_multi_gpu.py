"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): one process per GPU, batch sharded contiguously, operator
replicated, and the path's single collective -- a sum of a few scalars -- through the library's own NCCL communicator
(`sbp_comm_unique_id` / `sbp_comm_init` / `sbp_allreduce_sum`).  The host-side sharding logic is also covered on CPU
with gloo in tests/test_host.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import sbp_b200 as S
    from oracle import sbp_oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only: broadcasts the NCCL id
    ctx = S.Context(rank)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.numpy().tobytes()))

    N, B = 64, 10007
    D = S.PolynomialBasesDerivativeOperator(S.LobattoLegendre, -1.0, 1.0, N)
    U = np.random.default_rng(0).uniform(-1, 1, (B, N))
    b0, b1 = S.shard_range(B, world, rank)
    u = ctx.upload(U[b0:b1])
    du = D * u
    ref = O.apply_batch(D.basis.D, U[b0:b1], D.jac)
    err = float(np.max(np.abs(du.to_host() - ref)) / np.max(np.abs(ref)))
    # local energy sums (u'Pu and du'P du) on the device, then ONE allreduce of 2 scalars
    s = S.RawBuffer(ctx, 16)
    op = D.device_op(ctx)
    S._lib.check(ctx._lib.sbp_integrate(ctx.handle, op.handle, C.c_void_p(u.ptr), None, u.B, 1, None, C.c_void_p(s.ptr)))
    S._lib.check(ctx._lib.sbp_integrate(ctx.handle, op.handle, C.c_void_p(du.ptr), None, u.B, 1, None,
                                        C.c_void_p(s.ptr + 8)))
    ctx.allreduce_sum(s.ptr, 2)
    tot = s.to_host(np.float64, 2)
    if rank == 0:
        Yall = O.apply_batch(D.basis.D, U, D.jac)
        q.put((err, tot.tolist(), [float(O.energy_batch(D.weights(), U).sum()),
                                   float(O.energy_batch(D.weights(), Yall).sum())]))
    else:
        q.put((err, None, None))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_sharded_apply_and_nccl_energy_allreduce():
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [mctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=300) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(r[0] <= 1e-12 for r in results)
    tot, ref = next((r[1], r[2]) for r in results if r[1] is not None)
    assert abs(tot[0] - ref[0]) <= 1e-12 * ref[0] and abs(tot[1] - ref[1]) <= 1e-12 * ref[1]
