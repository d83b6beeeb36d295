"""The N>1 SMR path on CPU: driver.SmrRunSlabs (cut planes, per-level exchange, reductions, restriction / prolongation
order) run under gloo with world size 2, with the CPU ORACLE standing in for the CUDA engine behind the same Grid
interface -- test scaffolding only; the product never touches oracle/.  Hydro nests (the oracle's ion sub-cycle is one
C call and cannot take the cross-rank reductions).  The gathered result must equal the reference's dump bit for bit."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (loads the package as athena_b200, also in spawned workers)
import oracle_lib as ol
from smr_cases import SmrCase

NG = 4


class OracleBackedGrid:
    """engine.Grid's interface (the part a hydro SMR run uses) over oracle/athena_oracle.c."""

    def __init__(self, nx, dx, xmin, gamma, nscal, device=0, torch_memory=True):
        self.nx, self.dx = tuple(nx), tuple(dx)
        self.nvar = 5 + nscal
        shp = (self.nx[2] + 2 * NG, self.nx[1] + 2 * NG, self.nx[0] + 2 * NG)
        self.U_torch = torch.zeros((self.nvar,) + shp, dtype=torch.float64)
        self._U = self.U_torch.numpy()                         # shares memory: the exchange writes straight into it
        ptrs = [ol.P(self._U[v]) if v < self.nvar else None for v in range(6)]
        self.g = ol.OrcGrid(self.nx[0], self.nx[1], self.nx[2], nscal, *self.dx, *xmin, float(gamma), *ptrs, None)
        self.Fsave = [np.zeros(shp + (6,)) for _ in range(3)]
        for a in range(3):
            self.g.Fsave[a] = ol.P(self.Fsave[a])

    def upload_active(self, d):
        for n, v in enumerate(ol.VARS[:self.nvar]):
            self._U[n][NG:-NG, NG:-NG, NG:-NG] = d[v]

    def download_active(self):
        return {v: self._U[n][NG:-NG, NG:-NG, NG:-NG].copy() for n, v in enumerate(ol.VARS[:self.nvar])}

    def bvals_mhd(self, bc):
        ol.lib().orc_bvals(C.byref(self.g), (C.c_int * 6)(*bc))

    def max_wave_speeds(self):
        out = (C.c_double * 3)()
        ol.lib().orc_max_speeds(C.byref(self.g), out)
        return [out[0], out[1], out[2]]

    def integrate_3d_ctu(self, dt):
        ol.lib().orc_integrate_3d_ctu(C.byref(self.g), dt)

    @staticmethod
    def _ov(o):
        r = ol.OrcOverlap()
        for a in range(3):
            r.cs[a], r.ce[a], r.cdisp[a] = o.cs[a], o.ce[a], o.cdisp[a]
        for a in range(6):
            r.face[a] = o.face[a]
        return r

    def smr_restrict_correct(self, child, o, dt):
        ol.lib().orc_restrict_correct(C.byref(self.g), C.byref(child.g), C.byref(self._ov(o)), dt)

    def smr_prolongate(self, child, o):
        ol.lib().orc_prolongate(C.byref(self.g), C.byref(child.g), C.byref(self._ov(o)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import athena_b200.driver as drv
        drv.engine.Grid = OracleBackedGrid                     # the oracle behind the engine's interface, this process only
        case = SmrCase(name)
        R = drv.SmrRunSlabs(case.par, case.build0, [lv["U0"] for lv in case.levels], dist)
        ok = R.dt == float(case.gold["dt0"])
        R.run(case.nsteps)
        ok = ok and R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
        bad = []
        for n, lv in enumerate(case.levels):
            full = R.gather_level(n)
            for v in lv["U1"]:
                if not np.array_equal(full[v], lv["U1"][v]):
                    bad.append((n, v))
        held = [L is not None for L in R.levels]
        q.put((rank, ok and not bad, bad, held, R.cuts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["blastsmr_k4", "blastsmr3_k4"])
def test_smr_across_two_ranks_equals_reference_dump(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, bad, held, cuts in res:
        assert ok, f"rank {rank}: differs from the reference dump in {bad} (levels held {held}, cuts {cuts})"
