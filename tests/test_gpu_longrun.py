"""north_star bars that need many steps: volume-L1 <= 1e-8 after 100 steps and the ionisation-front position within
one cell, against the reference binary (oracle/_ref, prebuilt; travels with the snapshot) run on the box's host."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from cases import ROOT, DECKS, VARS
from athena_b200.athinput import ParFile
from athena_b200.restart import read_restart

pytestmark = pytest.mark.gpu


def reference_run(build, deck, overrides, nlim):
    exe = os.path.join(ROOT, "oracle", "_ref", build, "athena")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, "-i", os.path.join(DECKS, deck), f"time/nlim={nlim}"] + overrides, cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
        d0, d1 = read_restart(os.path.join(tmp, rst[0])), read_restart(os.path.join(tmp, rst[-1]))
    rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    return d0, d1, rad


def engine_run(problem, deck, overrides, nlim, U0):
    from athena_b200.driver import Run
    par = ParFile.par_open(os.path.join(DECKS, deck))
    par.par_cmdline(overrides)
    R = Run(par, problem, device=0, U0=U0)
    R.run(nlim)
    return R


def volume_l1(a, b):
    """sum |a-b| dV / sum |b| dV (uniform cells)"""
    return float(np.abs(a - b).sum() / np.abs(b).sum())


@pytest.mark.parametrize("problem,deck,ov", [
    ("blast", "blast.athinput", ["domain1/Nx1=32", "domain1/Nx2=32", "domain1/Nx3=32", "problem/radius=0.2"]),
    ("ifront", "ifront.athinput", ["domain1/Nx1=48", "domain1/Nx2=16", "domain1/Nx3=16"]),
])
def test_100_steps_volume_l1(problem, deck, ov):
    d0, d1, rad = reference_run(problem, deck, ov, 100)
    nv = 6 if problem != "blast" else 5
    U0 = {v: d0["grids"][0][v] for v in VARS[:nv]}
    R = engine_run(problem, deck, ov, 100, U0)
    assert R.nstep == d1["nstep"] == 100
    out = R.G.download_active()
    assert abs(R.time - d1["time"]) <= 1e-10 * d1["time"]
    for v in ("d", "E") + (("s",) if nv == 6 else ()):
        l1 = volume_l1(out[v], d1["grids"][0][v])
        assert l1 <= 1e-8, f"{problem}: volume-L1 of {v} after 100 steps = {l1:.3e}"
    # momenta relative to the total momentum scale sum(d*cs)
    ref = d1["grids"][0]
    scale = (ref["d"] * np.sqrt(5 / 3 * (2 / 3) * np.abs(ref["E"]) / ref["d"])).sum()
    for v in ("M1", "M2", "M3"):
        assert np.abs(out[v] - ref[v]).sum() / scale <= 1e-8, v
    if problem == "ifront":
        assert [r[0] for r in R.rad_log] == rad, "sub-cycle iteration counts over 100 steps"
        front = lambda U: np.argmax((U["s"] / U["d"]) >= 0.5, axis=2)
        assert np.abs(front(out) - front(ref)).max() <= 1, "I-front position within one cell on every ray"


def test_full_size_conservation_blast_256():
    """BASELINE configs[1] at its full size (256^3 periodic blast, problem() data generated as the reference does): no
    reference dump of this size travels, so the check is a size-independent property of the conservative update --
    total mass, momentum and energy after 10 steps equal the initial totals to round-off of a 16.7 M-term sum."""
    from athena_b200.driver import Run
    par = ParFile.par_open(os.path.join(DECKS, "blast.athinput"))
    par.par_cmdline(["job/num_domains=1", "domain1/Nx1=256", "domain1/Nx2=256", "domain1/Nx3=256"])
    R = Run(par, "blast", device=0)
    a = R.G.download_active()
    tot0 = {v: float(np.sum(a[v], dtype=np.longdouble)) for v in ("d", "M1", "M2", "M3", "E")}
    R.run(10)
    b = R.G.download_active()
    assert R.nstep == 10 and np.isfinite(b["E"]).all() and (b["d"] > 0).all()
    assert np.abs(b["d"] - a["d"]).max() > 1e-3, "the blast wave must actually have moved"
    n = a["d"].size
    for v in ("d", "E"):
        tot1 = float(np.sum(b[v], dtype=np.longdouble))
        assert abs(tot1 - tot0[v]) <= 1e-12 * abs(tot0[v]), f"{v}: {tot0[v]!r} -> {tot1!r}"
    scale = float(np.sum(np.abs(b["M1"]), dtype=np.longdouble))
    for v in ("M1", "M2", "M3"):
        tot1 = float(np.sum(b[v], dtype=np.longdouble))
        assert abs(tot1 - tot0[v]) <= 1e-12 * scale, f"{v}: {tot0[v]!r} -> {tot1!r} (scale {scale!r})"


def test_full_size_determinism_and_bounds_ioniz_256():
    """BASELINE configs[2] at its full size (256^3 ioniz_sphere: hydro + ion RT + static potential + Userwork reset).
    Size-independent properties: two independent runs agree bit for bit (every cross-thread reduction is a min, a max or an
    integer sum), the sub-cycle is identical, the neutral density stays within [0, d], and the radiation step leaves d alone."""
    from athena_b200.driver import Run
    par = ParFile.par_open(os.path.join(DECKS, "ioniz_sphere.athinput"))
    par.par_cmdline(["job/num_domains=1", "domain1/Nx1=256", "domain1/Nx2=256", "domain1/Nx3=256"])
    runs = []
    for _ in range(2):
        R = Run(par, "ioniz_sphere", device=0)
        R.run(3)
        runs.append((R.G.download_active(), list(R.rad_log), R.time, R.dt))
        del R
    (a, rad_a, ta, dta), (b, rad_b, tb, dtb) = runs
    assert rad_a == rad_b and ta == tb and dta == dtb and len(rad_a) == 3
    for v in a:
        assert np.array_equal(a[v], b[v], equal_nan=True), v
    assert (a["s"] > 0).all() and (a["s"] <= a["d"]).all()
    assert np.isfinite(a["E"]).all()


def run_against_fingerprint(name, setup=None):
    """Run the deck of tests/golden/<name>.json and compare SHA-256 of every variable (and EdgeFlux), time, dt and the sub-cycle
    iteration counts with the reference's own dump.  Returns the Run."""
    import importlib.util
    import json
    from athena_b200.driver import Run
    path = os.path.join(ROOT, "tests", "golden", name + ".json")
    if not os.path.exists(path):
        pytest.skip("fingerprint not generated")
    gold = json.load(open(path))
    spec = importlib.util.spec_from_file_location("fp5", os.path.join(ROOT, "tests", "golden", "make_hires5_fingerprint.py"))
    fp5 = importlib.util.module_from_spec(spec); spec.loader.exec_module(fp5)
    par = ParFile.par_open(os.path.join(DECKS, gold["deck"]))
    par.par_cmdline(gold["overrides"])
    R = Run(par, gold["build"], device=0)
    if setup is not None:
        setup(R)
    R.run(gold["nstep"])
    g = R.G.download_active()
    ion = gold["build"] != "blast"
    if ion:
        g["edgeflux"] = R.G.download_edgeflux()
    mine = fp5.fingerprint(dict(time=R.time, dt=R.dt, nstep=R.nstep, grids=[g]), [r[0] for r in R.rad_log], vars_=tuple(gold["levels"][0]))
    assert mine["rad"] == gold["rad"]
    assert mine["time"] == gold["time"] and mine["dt"] == gold["dt"]
    for v, ref in gold["levels"][0].items():
        got = mine["levels"][0][v]
        assert got["shape"] == ref["shape"]
        assert got["sha256"] == ref["sha256"], f"{v}: sum {got['sum']!r} vs {ref['sum']!r}, absmax {got['absmax']!r} vs {ref['absmax']!r}"
    return R


@pytest.mark.parametrize("name", ["ioniz256_k1", "blast256_k2", "ifront64_k30"])
def test_full_size_matches_reference_fingerprint(name):
    """BASELINE configs[2] / configs[1] / configs[0] AT FULL SIZE (256^3, 256^3, 64^3) against the reference itself: SHA-256 of every conserved
    variable (and EdgeFlux) of the reference's dump, generated by tests/golden/make_fullsize_fingerprint.py from the
    reference binary on this container's CPU; the engine's dump must hash to the same bytes."""
    run_against_fingerprint(name)


@pytest.mark.parametrize("name", ["ioniz96_k12", "ioniz80_k6", "ioniz1100_k4", "ifront64s_k3"])
@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("memo", [1, 0])
def test_sparse_subcycle_matches_reference_fingerprint(name, fuse, memo):
    """The sub-cycle memo (csrc/ion.cu "sparse sub-cycle": chunks of inert ambient cells re-use the previous iteration's rates
    and reductions; k_ion_walk lists the rays with chunks left to evaluate) changes nothing: ioniz_sphere boxes most of whose chunks are inert -- 96^3 over 12 steps and 51 sub-cycle
    iterations, 80x48x40 (ragged last chunk), 1100x12x10 (35 chunks per ray: more than one warp of records) -- and an ifront box where no chunk is, with the memo on and off, through the
    rates/update pair (fuse 0) and the fused update+rates kernel (fuse 1), hash to the reference's dump: every variable,
    EdgeFlux, time, dt and the iteration count of every step."""
    os.environ["AB2_ION_FUSE"] = str(fuse)
    try:
        R = run_against_fingerprint(name, setup=lambda R: R.G.set_ion_memo(memo))
    finally:
        os.environ.pop("AB2_ION_FUSE", None)
    nskip, ntask, n = R.G.ion_memo_stats()
    rays = R.nx[1] * R.nx[2]
    if not memo or fuse:                       # the memo belongs to the rates / update pair; the fused kernel evaluates everything
        assert nskip == 0 and ntask == 0
    elif name.startswith("ioniz"):
        # the last scan of the records: the bulk of the box was re-used, and not all of it (the planet is active)
        assert 0.5 * n < nskip < n and 0 < ntask <= rays, (nskip, ntask, n, rays)
    else:
        assert nskip == 0 and ntask == rays, (nskip, ntask)   # ifront: every chunk is ionising, nothing to re-use
