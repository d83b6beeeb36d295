"""Pins the oracle's SMR restatement (RestrictCorrect, Prolongate/ProCon, ionradRestrictCorrect, ionrad_prolong,
fine-level ion sub-cycle) against runs of the reference's own --enable-smr builds (tests/golden/*smr*): two-level
cases from the serial build, three-level cases from its MPI branch run as one rank (oracle/mpi_stub; the serial
build segfaults below level 1 with ion RT)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from smr_cases import SmrCase


def oracle_run_smr(case, log=None):
    grids, ovs = [], []
    for n, lv in enumerate(case.levels):
        if n == 0:
            G = ol.OracleGrid(lv["nx"], case.xmin, case.xmax, case.gamma, case.nscal, case.bc, keep_fluxes=True)
        else:
            G = ol.OracleGrid(lv["nx"], lv["xmin"], None, case.gamma, case.nscal, lv["bc"], dx=lv["dx"], keep_fluxes=True)
            pa = case.levels[n - 1]
            ovs.append(ol.make_overlap(pa["nx"], lv["nx"], lv["disp"], case.nx, level=n, parent_disp=pa["disp"]))
        G.set_active(lv["U0"])
        grids.append(G)
    L = ol.lib()
    nlev = len(grids)
    for n in range(nlev - 2, -1, -1):                                               # main.c:401
        L.orc_restrict_correct(C.byref(grids[n].g), C.byref(grids[n + 1].g), C.byref(ovs[n]), 0.0)
    for G in grids:
        G.bvals()
    if case.ion:
        for G in grids:
            G.bvals_ionrad(case.flux_i)                                             # bvals_ionrad.c: first Domain's BC fn wins
    for n in range(nlev - 1):                                                       # main.c:446
        L.orc_prolongate(C.byref(grids[n].g), C.byref(grids[n + 1].g), C.byref(ovs[n]))
    dt = ol.new_dt_smr(grids, case.cour_no, 0.0, 0, 0.0, case.tlim)
    assert dt == float(case.gold["dt0"])
    ionpar = ol.ion_par_from(case.par, grids[0].dx) if case.ion else None
    mesh_dt = ol.run_main_loop_levels(grids, ovs, case.nsteps, case.cour_no, case.tlim, ionpar, case.flux_i, log)
    return grids, mesh_dt


@pytest.mark.parametrize("name", ["blastsmr_k1", "blastsmr_k4", "ifrontsmr_k1", "ifrontsmr_k3",
                                  "blastsmr3_k4", "ifrontsmr3_k1", "ifrontsmr3_k3", "blastsmredge_k4", "ifrontsmredge_k3"])
def test_smr_run_matches_reference_dump(name):
    case = SmrCase(name)
    log = {}
    grids, mesh_dt = oracle_run_smr(case, log)
    assert grids[0].time == float(case.gold["time1"])
    assert mesh_dt == float(case.gold["dt1"])
    if case.ion:
        assert [r[0] for r in log["rad"]] == [int(x) for x in case.gold["rad"][:, 0]]
    for n, (G, lv) in enumerate(zip(grids, case.levels)):
        for v in lv["U1"]:
            assert np.array_equal(G.active(v), lv["U1"][v]), f"{name}: level {n} {v}, max abs {np.abs(G.active(v)-lv['U1'][v]).max()}"
        if case.ion:
            assert np.array_equal(G.edgeflux, lv["edgeflux1"]), f"{name}: level {n} EdgeFlux"


def test_three_level_case_hands_down_nontrivial_flux():
    """the fixture only tests ionrad_prolong if both fine levels really receive an attenuated, non-zero flux"""
    case = SmrCase("ifrontsmr3_k3")
    entering = case.levels[0]["edgeflux1"][:, :, 0].max()
    f1 = case.levels[1]["edgeflux1"][:, :, 0]
    f2 = case.levels[2]["edgeflux1"][:, :, 0]
    assert 0 < f2.max() < f1.min() <= f1.max() < entering
