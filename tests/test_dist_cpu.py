"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU path: row partitioning,
sharding-independent RNG, the NCCL-id rendezvous plumbing, and that row-sharded partial inner
products + allreduce reproduce the single-rank oracle (the contract of SURVEY.md 8e)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightkrylov_b200.api import partition
    from oracle import lk_oracle as lo
    lo.set_threads(1)
    nx, ny = 64, 37                                  # ragged split of the slow axis
    y0, nyl = partition(ny, world, rank)
    n, nloc, row0 = nx * ny, nx * nyl, nx * y0
    # 1. partitions tile the global range
    spans = [None] * world
    dist.all_gather_object(spans, (row0, nloc))
    assert spans[0][0] == 0 and sum(s[1] for s in spans) == n
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    # 2. RNG keyed on the global row: the local slab equals the slice of the global vector
    full = lo.fill(n, "d", "uniform", 42)
    mine = lo.fill(nloc, "d", "uniform", 42, row0)
    assert np.array_equal(mine, full[row0:row0 + nloc])
    # 3. sharded partial dots + allreduce == global dot (what the device multi-dot + NCCL does)
    j = 5
    V = np.asfortranarray(np.stack([lo.fill(n, "d", "normal", 100 + i) for i in range(j)], axis=1))
    part = V[row0:row0 + nloc].T @ mine
    t = torch.from_numpy(part.copy()); dist.all_reduce(t)
    assert np.allclose(t.numpy(), V.T @ full, rtol=1e-13)
    # 4. stencil slab + halo rows == global stencil restricted to the slab
    A = lo.Op.stencil("d", (nx, ny), (4.0, -1.0, -1.0, -1.0, -1.0))
    yfull = A.apply(full)
    lo_row = full[row0 - nx:row0] if y0 > 0 else np.zeros(nx)
    hi_row = full[row0 + nloc:row0 + nloc + nx] if y0 + nyl < ny else np.zeros(nx)
    ext = np.concatenate([lo_row, mine, hi_row])
    Aext = lo.Op.stencil("d", (nx, nyl + 2), (4.0, -1.0, -1.0, -1.0, -1.0))
    assert np.allclose(Aext.apply(ext)[nx:-nx], yfull[row0:row0 + nloc], rtol=1e-14, atol=1e-14)
    # 5. rendezvous plumbing used by Context.from_torch_distributed: rank 0's id reaches everyone
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_partition_properties():
    from lightkrylov_b200.api import partition
    for n in (1, 7, 4096, 4097, 513):
        for w in (1, 2, 3, 4, 8):
            spans = [partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
