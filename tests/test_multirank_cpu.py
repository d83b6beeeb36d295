"""N>1 host logic on CPU (gloo, world size 2): x3-slab ghost exchange in place of bvals_mhd's MPI path
(bvals_mhd.c:423-446) and the all_reduce callback that stands where the reference calls MPI_Allreduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (loads the package under the name athena_b200, also in spawned workers)
import oracle_lib as ol
from athena_b200.driver import exchange_x3_ghosts, slab_bounds, Run

NG = 4


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_with_ghosts(bc, nx=(6, 5, 8), nscal=1, seed=0):
    rng = np.random.default_rng(seed)
    G = ol.OracleGrid(nx, (0, 0, 0), (1, 1, 1), 5 / 3, nscal, bc)
    for v in G.U:
        G.active(v)[...] = rng.normal(size=G.active(v).shape)
    G.bvals()
    return np.stack([G.U[v] for v in G.U])


def _worker(rank, world, port, bc, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Ug = _global_with_ghosts(bc)
        nx3 = Ug.shape[1] - 2 * NG
        k0, k1 = slab_bounds(nx3, world, rank)
        slab = torch.from_numpy(Ug[:, k0:k1 + 2 * NG].copy())
        want = slab.clone()
        # wipe the x3 ghost layers a neighbouring slab owns (physical faces keep what bvals wrote)
        if rank > 0 or bc[4] == 4:
            slab[:, :NG] = -777.0
        if rank < world - 1 or bc[5] == 4:
            slab[:, -NG:] = -777.0
        exchange_x3_ghosts(slab, rank, world, bc[4], bc[5], dist)
        ok_exchange = bool(torch.equal(slab, want))
        # reduction callback (ops: 0 MIN, 1 MAX, 2 SUM)
        class Shell(Run):
            def __init__(self):
                self.dist = dist
        red = Shell()._make_reduce()
        v = np.array([1.0 + rank, 10.0 - rank])
        ok_red = (np.array_equal(red(v, 0), [1.0, 10.0 - (world - 1)]) and np.array_equal(red(v, 1), [float(world), 10.0])
                  and np.array_equal(red(v, 2), [sum(1.0 + r for r in range(world)), sum(10.0 - r for r in range(world))]))
        v4 = np.array([1.0 + rank, 10.0 - rank, 5.0 + rank, 7.0 - rank])
        ok_red = ok_red and np.array_equal(red(v4, 4), [sum(1.0 + r for r in range(world)), 10.0, 5.0, 7.0 - (world - 1)])
        q.put((rank, ok_exchange, ok_red))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bc", [(2, 2, 2, 2, 2, 2), (4, 4, 4, 4, 4, 4), (1, 2, 2, 1, 1, 2)])
def test_slab_exchange_and_reductions_world2(bc):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_ex, ok_red in res:
        assert ok_ex, f"rank {rank}: ghost slabs differ from the single-Grid bvals result"
        assert ok_red, f"rank {rank}: reduction callback"
