"""Pins the CPU oracle (oracle/athena_oracle.c) against the reference's own outputs:
 * function level: reference lr_states()/fluxes()/Cons1D_to_Prim1D()/cfast() vectors (tests/golden/units_*.npz,
   produced by oracle/_ref/<cfg>/libref_units.so = the reference's objects) -- bit-exact;
 * run level: restart dumps of the reference binary after K steps (tests/golden/<case>_kK.npz) --
   hydro bit-exact, ion RT bit-exact as well (oracle and reference share the host libm)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from cases import Case, load_golden, load_units, VARS


@pytest.mark.parametrize("nscal", [0, 1])
def test_units_bit_exact(nscal):
    g = load_golden(f"units_nscal{nscal}")
    L = ol.lib()
    gam = float(g["gamma"])
    n = len(g["Wl"])
    c = lambda k: np.ascontiguousarray(g[k])
    U = np.empty((n, 6)); F = np.empty((n, 6)); W = np.empty((n, 6))
    L.orc_prim_to_cons_batch(n, ol.P(c("Wl")), nscal, gam, ol.P(U))
    assert np.array_equal(U, g["Ul"])
    L.orc_roe_flux_batch(n, ol.P(c("Ul")), ol.P(c("Ur")), ol.P(c("Wl")), ol.P(c("Wr")), nscal, gam, ol.P(c("eta")), ol.P(F))
    assert np.array_equal(F, g["F"])
    L.orc_cons_to_prim_batch(n, ol.P(c("Ul")), nscal, gam, ol.P(W))
    assert np.array_equal(W, g["W_of_Ul"])
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    for t in range(len(g["pencil_W"])):
        Wp = np.ascontiguousarray(g["pencil_W"][t])
        n = len(Wp)
        Wl = np.zeros((n, 6)); Wr = np.zeros((n, 6))
        L.orc_lr_states(ol.P(Wp), nscal, gam, float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, ol.P(Wl), ol.P(Wr))
        assert np.array_equal(Wl[il:iu + 2], g["pencil_Wl"][t][il:iu + 2])
        assert np.array_equal(Wr[il - 1:iu + 1], g["pencil_Wr"][t][il - 1:iu + 1])


@pytest.mark.parametrize("nscal", [0, 1])
def test_hllc_flux_bit_exact(nscal):
    """rsolvers/hllc.c's fluxes() (reference object, oracle/_ref/*_hllc/libref_units.so) on the same seeded states"""
    g = load_units("hllc_", nscal)
    n = len(g["Wl"])
    c = lambda k: np.ascontiguousarray(g[k])
    F = np.empty((n, 6))
    ol.lib().orc_hllc_flux_batch(n, ol.P(c("Ul")), ol.P(c("Ur")), ol.P(c("Wl")), ol.P(c("Wr")), nscal, float(g["gamma"]), ol.P(F))
    assert np.array_equal(F, g["F"])
    roe = load_golden(f"units_nscal{nscal}")
    assert not np.array_equal(roe["F"], g["F"]), "same inputs, a different solver"
    # PLM steps 8-9 in their HLL-family form (lr_states_plm.c:257-293)
    L = ol.lib()
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    L.orc_set_riemann_solver(1)
    try:
        for t in range(len(g["pencil_W"])):
            Wp = np.ascontiguousarray(g["pencil_W"][t])
            m = len(Wp)
            Wl = np.zeros((m, 6)); Wr = np.zeros((m, 6))
            L.orc_lr_states(ol.P(Wp), nscal, float(g["gamma"]), float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, ol.P(Wl), ol.P(Wr))
            assert np.array_equal(Wl[il:iu + 2], g["pencil_Wl"][t][il:iu + 2])
            assert np.array_equal(Wr[il - 1:iu + 1], g["pencil_Wr"][t][il - 1:iu + 1])
    finally:
        L.orc_set_riemann_solver(0)
    assert not np.array_equal(roe["pencil_Wl"], g["pencil_Wl"])


@pytest.mark.parametrize("tag,nscal,riemann", [("ppm_", 0, 0), ("ppm_", 1, 0), ("ppm_hllc_", 0, 1)])
def test_ppm_lr_states_bit_exact(tag, nscal, riemann):
    """reconstruction/lr_states_ppm.c (reference object built with THIRD_ORDER_CHAR) on the seeded pencils, in its CTU and
    its HLL-family form"""
    g = load_units(tag, nscal)
    L = ol.lib()
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    L.orc_set_reconstruction(3); L.orc_set_riemann_solver(riemann)
    try:
        for t in range(len(g["pencil_W"])):
            Wp = np.ascontiguousarray(g["pencil_W"][t])
            m = len(Wp)
            Wl = np.zeros((m, 6)); Wr = np.zeros((m, 6))
            L.orc_lr_states(ol.P(Wp), nscal, float(g["gamma"]), float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, ol.P(Wl), ol.P(Wr))
            assert np.array_equal(Wl[il:iu + 2], g["pencil_Wl"][t][il:iu + 2])
            assert np.array_equal(Wr[il - 1:iu + 1], g["pencil_Wr"][t][il - 1:iu + 1])
    finally:
        L.orc_set_reconstruction(2); L.orc_set_riemann_solver(0)
    plm = load_units("hllc_" if riemann else "", nscal)
    assert not np.array_equal(plm["pencil_Wl"], g["pencil_Wl"]), "same pencils, a different reconstruction"


def oracle_run(case, log=None):
    ol.lib().orc_set_riemann_solver(case.riemann)
    ol.lib().orc_set_reconstruction(case.order)
    try:
        return _oracle_run(case, log)
    finally:
        ol.lib().orc_set_riemann_solver(0)
        ol.lib().orc_set_reconstruction(2)


def _oracle_run(case, log=None):
    grav = case.grav_fn()
    G = ol.OracleGrid(case.nx, case.xmin, case.xmax, case.gamma, case.nscal, case.bc, grav=grav)
    G.set_active(case.U0)
    G.bvals()
    if case.ion:
        G.bvals_ionrad(case.flux_i)
    G.new_dt(case.cour_no, case.tlim)
    assert G.dt == float(case.gold["dt0"])
    ionpar = ol.ion_par_from(case.par, G.dx) if case.ion else None
    userwork = None
    tab = case.reset_table()
    if tab is not None:
        cells, vals = tab

        def userwork(G):
            for n, v in enumerate(VARS):
                G.U[v].reshape(-1)[cells] = vals[n]
    ol.run_main_loop(G, case.nsteps, case.cour_no, case.tlim, ionpar, case.flux_i, userwork, log)
    return G


@pytest.mark.parametrize("name", ["blast16_k1", "blast16_k6", "blastmix_k5", "ifront32_k1", "ifront32_k4",
                                  "ioniz24_k1", "ioniz24_k3", "blast16hllc_k6", "ifront32hllc_k4", "blast16noh_k6",
                                  "blast16ppm_k6", "ifront32ppm_k4", "blast16ppmhllc_k6",
                                  "blasttiny_k5", "ifronttiny_k3", "ionizodd_k1"])
def test_run_matches_reference_dump(name):
    case = Case(name)
    log = {}
    G = oracle_run(case, log)
    assert G.nstep == case.nsteps
    assert np.array_equal(np.array(log["next_dt"]), case.gold["next_dt"][1:]) or \
        np.allclose(np.array(log["next_dt"]), case.gold["next_dt"][1:], rtol=5e-7)   # stdout prints 7 digits
    assert G.time == float(case.gold["time1"])
    assert G.dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(G.active(v), case.U1[v]), f"{name}: {v} differs, max abs {np.abs(G.active(v)-case.U1[v]).max()}"
    if case.ion:
        rad = case.gold["rad"]
        assert [r[0] for r in log["rad"]] == [int(x) for x in rad[:, 0]]
        assert np.array_equal(G.edgeflux, case.gold["U1_edgeflux"])
