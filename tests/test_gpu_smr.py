"""SMR parity on the GPU: two-level runs (root + one level-1 Grid) through the C ABI against the dumps of the
reference's own --enable-smr builds.  Hydro (blast) bit for bit; hydro + ion RT (ifront) within the ion tolerance,
identical sub-cycle counts on both levels, EdgeFlux on both levels."""
import numpy as np
import pytest

from smr_cases import SmrCase

pytestmark = pytest.mark.gpu


def gpu_run_smr(case):
    from athena_b200.driver import SmrRun
    R = SmrRun(case.par, case.build0, case.U0, case.U0L1)
    assert R.dt == float(case.gold["dt0"])
    R.run(case.nsteps)
    return R


@pytest.mark.parametrize("name", ["blastsmr_k1", "blastsmr_k4"])
def test_smr_hydro_bit_exact(name):
    case = SmrCase(name)
    R = gpu_run_smr(case)
    a, b = R.root.download_active(), R.child.download_active()
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(a[v], case.U1[v]), f"root {v}: max abs {np.abs(a[v]-case.U1[v]).max()}"
        assert np.array_equal(b[v], case.U1L1[v]), f"level-1 {v}: max abs {np.abs(b[v]-case.U1L1[v]).max()}"


@pytest.mark.parametrize("name", ["ifrontsmr_k1", "ifrontsmr_k3"])
def test_smr_ion(name):
    case = SmrCase(name)
    R = gpu_run_smr(case)
    a, b = R.root.download_active(), R.child.download_active()
    assert [r[0] for r in R.rad_log] == [int(x) for x in case.gold["rad"][:, 0]]
    assert abs(R.time - float(case.gold["time1"])) <= 1e-10 * float(case.gold["time1"])
    for v in ("d", "E", "s"):
        assert np.allclose(a[v], case.U1[v], rtol=1e-10, atol=0), f"root {v}"
        assert np.allclose(b[v], case.U1L1[v], rtol=1e-10, atol=0), f"level-1 {v}"
    assert np.allclose(R.root.download_edgeflux(), case.gold["U1_edgeflux"], rtol=1e-10, atol=0)
    assert np.allclose(R.child.download_edgeflux(), case.gold["U1L1_edgeflux"], rtol=1e-10, atol=0)
