"""Pins the oracle's SMR restatement (RestrictCorrect, Prolongate/ProCon, ionradRestrictCorrect, ionrad_prolong,
fine-level ion sub-cycle) against two-level runs of the reference's own --enable-smr builds (tests/golden/*smr*)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from smr_cases import SmrCase


def oracle_run_smr(case, log=None):
    root = ol.OracleGrid(case.nx, case.xmin, case.xmax, case.gamma, case.nscal, case.bc, keep_fluxes=True)
    child = ol.OracleGrid(case.nx1, case.xmin1, None, case.gamma, case.nscal, case.bc1, dx=case.dx1, keep_fluxes=True)
    o = ol.make_overlap(case.nx, case.nx1, case.disp1, case.nx)
    root.set_active(case.U0); child.set_active(case.U0L1)
    L = ol.lib()
    L.orc_restrict_correct(C.byref(root.g), C.byref(child.g), C.byref(o), 0.0)     # main.c:401
    root.bvals(); child.bvals()
    if case.ion:
        root.bvals_ionrad(case.flux_i); child.bvals_ionrad(case.flux_i)             # bvals_ionrad.c: first Domain's BC fn wins
    L.orc_prolongate(C.byref(root.g), C.byref(child.g), C.byref(o))                # main.c:446
    dt = ol.new_dt_smr([root, child], case.cour_no, 0.0, 0, 0.0, case.tlim)
    assert dt == float(case.gold["dt0"])
    ionpar = ol.ion_par_from(case.par, root.dx) if case.ion else None
    mesh_dt = ol.run_main_loop_smr(root, child, o, case.nsteps, case.cour_no, case.tlim, ionpar, case.flux_i, log)
    return root, child, mesh_dt


@pytest.mark.parametrize("name", ["blastsmr_k1", "blastsmr_k4", "ifrontsmr_k1", "ifrontsmr_k3"])
def test_smr_run_matches_reference_dump(name):
    case = SmrCase(name)
    log = {}
    root, child, mesh_dt = oracle_run_smr(case, log)
    assert root.time == float(case.gold["time1"])
    assert mesh_dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(root.active(v), case.U1[v]), f"{name}: root {v}, max abs {np.abs(root.active(v)-case.U1[v]).max()}"
        assert np.array_equal(child.active(v), case.U1L1[v]), f"{name}: level-1 {v}, max abs {np.abs(child.active(v)-case.U1L1[v]).max()}"
    if case.ion:
        assert [r[0] for r in log["rad"]] == [int(x) for x in case.gold["rad"][:, 0]]
        assert np.array_equal(root.edgeflux, case.gold["U1_edgeflux"])
        assert np.array_equal(child.edgeflux, case.gold["U1L1_edgeflux"])
