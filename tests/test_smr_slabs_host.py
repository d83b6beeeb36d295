"""Host logic of SMR across ranks (driver.smr_slab_layout / balanced_cuts / exchange_x3_with / make_group_reduce), on
CPU: the layout invariants that keep every parent-child operation on one device, and the x3 exchange and grouped
reductions over gloo with world size 2 (a level held by one rank only, a level shared by both)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (loads the package as athena_b200, also in spawned workers)
from cases import DECKS
from athena_b200.athinput import ParFile
from athena_b200.driver import smr_slab_layout, balanced_cuts, exchange_x3_with, make_group_reduce

NG = 4


def hires():
    return ParFile.par_open(os.path.join(DECKS, "ioniz_sphere_hires.athinput"))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5])
def test_layout_invariants_on_the_shipped_five_level_nest(world):
    par = hires()
    cuts = balanced_cuts(par, 5, world)
    lay = smr_slab_layout(par, 5, world, cuts)
    total = 0
    for n, L in enumerate(lay["levels"]):
        covered = 0
        for q, S in enumerate(L["slabs"]):
            if S is None:
                continue
            assert S["nx"][2] >= NG
            covered += S["nx"][2]
            total += S["nx"][0] * S["nx"][1] * S["nx"][2]
            # neighbours exist exactly at cut planes interior to the Domain, and they hold the adjacent slab
            if S["lo"] is not None:
                assert L["slabs"][S["lo"]]["kb"] == S["ka"]
            else:
                assert S["ka"] == S["dom"][0]
            if S["hi"] is not None:
                assert L["slabs"][S["hi"]]["ka"] == S["kb"]
            else:
                assert S["kb"] == S["dom"][1]
            if n > 0:
                pa = lay["levels"][n - 1]["slabs"][q]
                o = S["ov"]
                assert pa is not None and 2 * pa["ka"] <= S["ka"] and S["kb"] <= 2 * pa["kb"], "child slab nested in parent slab"
                for a in range(3):
                    assert NG <= o["cs"][a] <= o["ce"][a] <= pa["nx"][a] + NG - 1, "overlap inside the parent's active zones"
                    assert 2 * (o["ce"][a] - o["cs"][a] + 1) == S["nx"][a]
                # a fine/coarse face in x3 only at the Domain's own ends, never at a cut plane
                assert o["face"][4] == (1 if S["lo"] is None else 0) and o["face"][5] == (1 if S["hi"] is None else 0)
                # at least the prolongation stencil (nghost/2 + 2 coarse cells) of parent data beyond the overlap, ghosts included
                assert o["cs"][2] - 3 >= 0 and o["ce"][2] + 3 <= pa["nx"][2] + 2 * NG - 1
        assert covered == L["nx"][2], f"level {n}: slabs must tile the Domain"
    assert total == 80**3 + 52**3 + 96**3 + 176**3 + 320**3
    zones = [sum(S["nx"][0] * S["nx"][1] * S["nx"][2] for L in lay["levels"] for S in [L["slabs"][q]] if S) for q in range(world)]
    assert total / world / max(zones) > 0.9, f"balance {zones}"


def test_layout_rejects_bad_cuts():
    par = hires()
    with pytest.raises(ValueError):
        smr_slab_layout(par, 5, 2, [0, 30, 80])      # level 4 starts at root layer 30: flux correction would cross ranks
    with pytest.raises(ValueError):
        smr_slab_layout(par, 5, 2, [0, 28, 80])      # leaves rank 0 two level-1 layers (54..56)
    with pytest.raises(ValueError):
        smr_slab_layout(par, 5, 3, [0, 40, 42, 80])  # a 2-layer root slab cannot feed 4 ghost layers
    with pytest.raises(ValueError):
        balanced_cuts(par, 5, 24)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # level A: shared by both ranks (rank 0 below rank 1); level B: held by rank 1 only -> no neighbours, own group
        rng = np.random.default_rng(5)
        full = torch.from_numpy(rng.normal(size=(3, 8 + 2 * NG, 5, 6)))
        k0, k1 = (0, 4) if rank == 0 else (4, 8)
        slab = full[:, k0:k1 + 2 * NG].clone()
        want = slab.clone()
        if rank == 0:
            slab[:, -NG:] = -1.0
        else:
            slab[:, :NG] = -1.0
        exchange_x3_with(slab, None if rank == 0 else 0, 1 if rank == 0 else None, dist)
        ok = bool(torch.equal(slab, want))
        g_all = None
        g_one = dist.new_group([1])                  # collective call on both ranks
        red_all = make_group_reduce(dist, g_all)
        v = np.array([1.0 + rank, 10.0 - rank, 5.0 + rank, 7.0 - rank])
        ok = ok and np.array_equal(red_all(v, 4), [3.0, 10.0, 5.0, 6.0]) and np.array_equal(red_all(v[:2], 0), [1.0, 9.0])
        if rank == 1:
            exchange_x3_with(slab, None, None, dist)  # level B: nothing to exchange
            red_one = make_group_reduce(dist, g_one)
            ok = ok and np.array_equal(red_one(v, 4), v) and np.array_equal(red_one(v[:3], 1), v[:3])
        ok = ok and np.array_equal(red_all(np.array([float(rank)]), 1), [1.0])    # both ranks meet again afterwards
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_exchange_and_grouped_reductions_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
