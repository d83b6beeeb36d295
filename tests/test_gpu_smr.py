"""SMR parity on the GPU through the C ABI against the dumps of the reference's own --enable-smr builds: two-level
runs (serial build) and three-level runs (its MPI branch as one rank, oracle/mpi_stub -- the only build of the
reference that survives a Grid below level 1 with ion RT).  Hydro and ion RT bit for bit on every level, identical
sub-cycle counts, EdgeFlux on every level."""
import numpy as np
import pytest

from smr_cases import SmrCase

pytestmark = pytest.mark.gpu


def gpu_run_smr(case):
    from athena_b200.driver import SmrRun
    R = SmrRun(case.par, case.build0, *[lv["U0"] for lv in case.levels])
    assert R.dt == float(case.gold["dt0"])
    R.run(case.nsteps)
    return R


@pytest.mark.parametrize("name", ["blastsmr_k1", "blastsmr_k4", "blastsmr3_k4", "blastsmredge_k4"])
def test_smr_hydro_bit_exact(name):
    case = SmrCase(name)
    R = gpu_run_smr(case)
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for n, (G, lv) in enumerate(zip(R.grids, case.levels)):
        a = G.download_active()
        for v in lv["U1"]:
            assert np.array_equal(a[v], lv["U1"][v]), f"level {n} {v}: max abs {np.abs(a[v]-lv['U1'][v]).max()}"


@pytest.mark.parametrize("name", ["ifrontsmr_k1", "ifrontsmr_k3", "ifrontsmr3_k1", "ifrontsmr3_k3", "ifrontsmredge_k3"])
@pytest.mark.parametrize("pair", [False, True])
def test_smr_ion(name, pair):
    """pair: force the rates / update pair with its sub-cycle memo and task lists (AB2_ION_FUSE=0; Grids this small take the
    fused update+rates kernel by default), so that the fine levels' entering flux from EdgeFlux goes through k_ion_walk too."""
    import os
    case = SmrCase(name)
    if pair:
        os.environ["AB2_ION_FUSE"] = "0"
    try:
        R = gpu_run_smr(case)
    finally:
        os.environ.pop("AB2_ION_FUSE", None)
    assert [r[0] for r in R.rad_log] == [int(x) for x in case.gold["rad"][:, 0]]
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for n, (G, lv) in enumerate(zip(R.grids, case.levels)):
        a = G.download_active()
        for v in lv["U1"]:
            assert np.array_equal(a[v], lv["U1"][v]), f"level {n} {v}: max rel {np.abs(a[v]/lv['U1'][v]-1).max()}"
        assert np.array_equal(G.download_edgeflux(), lv["edgeflux1"]), f"level {n} EdgeFlux"
