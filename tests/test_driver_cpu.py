"""The host driver (athena_b200/driver.py: Run, SmrRun, SmrRunSlabs -- the Python mirror of main.c's loop that tests and
bench.py drive the engine with) exercised on CPU: the CPU ORACLE stands in for the CUDA engine behind the same Grid
interface (test scaffolding only; the product never touches oracle/).  What is checked is the driver's call order, scalar
handling, problem hooks and, for SmrRunSlabs under gloo with world size 2, the cut planes, per-level exchange and
reductions: every run must equal the reference's dump bit for bit."""
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

    # ---- the rest of what driver.Run uses ----
    def set_riemann_solver(self, kind):
        self._riemann = kind

    def set_reconstruction(self, order):
        self._order = order

    def upload(self, U):
        for n, v in enumerate(ol.VARS[:self.nvar]):
            self._U[n][...] = U[v]

    def set_static_grav_pot(self, fn):
        self._grav = fn                                        # keep the ctypes function alive
        self.g.grav = C.cast(fn, C.c_void_p)

    def set_fixed_cells(self, cells, vals):
        self._fixed = (np.asarray(cells), np.asarray(vals))

    def apply_fixed_cells(self):
        cells, vals = self._fixed
        for n in range(self.nvar):
            self._U[n].reshape(-1)[cells] = vals[n]

    def new_dt(self, cour_no, dt_old, nstep, time, tlim, reduce=None):
        if reduce is None:
            return ol.lib().orc_new_dt(C.byref(self.g), cour_no, dt_old, nstep, time, tlim)
        v = self.max_wave_speeds()                              # as ab2_new_dt: local dt, MIN over ranks, then the tlim clip
        max_dti = 0.0
        for a in range(3):
            if self.nx[a] > 1:
                x = v[a] / self.dx[a]
                max_dti = max_dti if max_dti > x else x
        dt = cour_no / max_dti if nstep == 0 else min(2.0 * dt_old, cour_no / max_dti)
        dt = float(reduce(np.array([dt]), 0)[0])
        if time < tlim and (tlim - time) < dt:
            dt = tlim - time
        return dt

    # ---- ion RT (single rank: the oracle's sub-cycle is one C call) ----
    def ion_init(self, ip):
        self.ionpar = ol.OrcIonPar()
        for n, _ in ol.OrcIonPar._fields_:
            setattr(self.ionpar, n, getattr(ip, n))
        self.edgeflux = np.zeros((self.nx[2] + 1, self.nx[1] + 1, self.nx[0] + 1))
        self.ion_coarse_time = 0.0

    def bvals_ionrad(self, flux_i):
        ol.lib().orc_bvals_ionrad(C.byref(self.g), flux_i, ol.P(self.edgeflux))

    def ion_radtransfer_3d(self, flux_i, time, cour_no, dt, reduce=None):
        assert reduce is None
        d, nchem, ntherm, tc = C.c_double(dt), C.c_int(0), C.c_int(0), C.c_double(0.0)
        niter = ol.lib().orc_ion_radtransfer_root(C.byref(self.g), C.byref(self.ionpar), flux_i, time, cour_no, C.byref(d),
                                                  ol.P(self.edgeflux), C.byref(nchem), C.byref(ntherm), C.byref(tc))
        self.ion_coarse_time = tc.value
        return d.value, niter, ntherm.value, nchem.value

    def ion_radtransfer_3d_fine(self, tcoarse, dt, reduce=None):
        assert reduce is None
        d, nchem, ntherm = C.c_double(dt), C.c_int(0), C.c_int(0)
        niter = ol.lib().orc_ion_radtransfer_fine(C.byref(self.g), C.byref(self.ionpar), tcoarse, C.byref(d), ol.P(self.edgeflux),
                                                  C.byref(nchem), C.byref(ntherm))
        return d.value, niter, ntherm.value, nchem.value

    def smr_ion_prolong(self, child, o):
        ol.lib().orc_ion_prolong(C.byref(self.g), ol.P(self.edgeflux), C.byref(child.g), ol.P(child.edgeflux), C.byref(self._ov(o)))

    def smr_ion_restrict(self, child, o):
        ol.lib().orc_ion_restrict(C.byref(self.g), C.byref(child.g), C.byref(self._ov(o)))

    def download_edgeflux(self):
        return self.edgeflux.copy()


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


def _worker_unigrid(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import athena_b200.driver as drv
        from cases import Case
        drv.engine.Grid = OracleBackedGrid
        case = Case(name)
        R = drv.Run(case.par, case.problem, dist=dist, U0=case.U0)
        ok = R.dt == float(case.gold["dt0"])
        R.run(case.nsteps)
        ok = ok and R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
        a = R.G.download_active()
        bad = [v for v in case.U1 if not np.array_equal(a[v], case.U1[v][R.k0:R.k1])]
        q.put((rank, ok and not bad, bad, [True], [R.k0, R.k1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["blast16_k6", "blastmix_k5"])
def test_unigrid_across_two_ranks_equals_reference_dump(name):
    """driver.Run with the Grid cut into two x3 slabs (periodic ring for blast16, physical faces for blastmix)"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_unigrid, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, bad, _, ks in res:
        assert ok, f"rank {rank} (layers {ks}): differs from the reference dump in {bad}"


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


@pytest.mark.parametrize("name", ["blastsmr_k4", "ifrontsmr_k3", "ifrontsmr3_k3", "ifrontsmredge_k3"])
def test_driver_smr_loop_on_cpu_equals_reference_dump(name, monkeypatch):
    """driver.SmrRun itself (the Python mirror of main.c's SMR loop that tests and bench.py drive the engine with) over
    the oracle: its call order and scalar handling reproduce the reference's dumps, ion RT and EdgeFlux included."""
    import athena_b200.driver as drv
    monkeypatch.setattr(drv.engine, "Grid", OracleBackedGrid)
    case = SmrCase(name)
    R = drv.SmrRun(case.par, case.build0, *[lv["U0"] for lv in case.levels])
    assert R.dt == float(case.gold["dt0"])
    R.run(case.nsteps)
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    if case.ion:
        assert [r[0] for r in R.rad_log] == [int(x) for x in case.gold["rad"][:, 0]]
    for n, (G, lv) in enumerate(zip(R.grids, case.levels)):
        a = G.download_active()
        for v in lv["U1"]:
            assert np.array_equal(a[v], lv["U1"][v]), f"level {n} {v}"
        if case.ion:
            assert np.array_equal(G.download_edgeflux(), lv["edgeflux1"]), f"level {n} EdgeFlux"


@pytest.mark.parametrize("name", ["blast16_k6", "blastmix_k5", "ifront32_k4", "ioniz24_k3", "ionizodd_k1", "ifronttiny_k3"])
def test_driver_unigrid_loop_on_cpu_equals_reference_dump(name, monkeypatch):
    """driver.Run (main.c's loop for one root Grid: problem hooks, static potential, Userwork replay, ion RT, new_dt) over
    the oracle reproduces the reference's dumps."""
    import athena_b200.driver as drv
    from cases import Case
    monkeypatch.setattr(drv.engine, "Grid", OracleBackedGrid)
    case = Case(name)
    R = drv.Run(case.par, case.problem, U0=case.U0)
    assert R.dt == float(case.gold["dt0"])
    R.run(case.nsteps)
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    a = R.G.download_active()
    for v in case.U1:
        assert np.array_equal(a[v], case.U1[v], equal_nan=True), v
    if case.ion:
        assert [r[0] for r in R.rad_log] == [int(x) for x in case.gold["rad"][:, 0]]
        assert np.array_equal(R.G.download_edgeflux(), case.gold["U1_edgeflux"])
