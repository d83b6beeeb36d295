"""Drop-in test: the reference's UNMODIFIED driver (main.c loop, par.c, problem file, restart dumps) linked
against shim/athena_shim.c + libathena_b200.so (build/dropin/<cfg>/athena, built by shim/build_dropin.py in the
container) runs the golden decks on the GPU; its restart dumps are compared with the reference's own."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from cases import Case, ROOT, DECKS
from athena_b200.restart import read_restart

pytestmark = pytest.mark.gpu


def run_dropin(case):
    exe = os.path.join(ROOT, "build", "dropin", case.build, "athena")
    if not os.path.exists(exe):
        pytest.skip("build/dropin not built (needs the reference sources: shim/build_dropin.py)")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-i", os.path.join(DECKS, str(case.gold["deck"])), f"time/nlim={case.nsteps}"] + [str(x) for x in case.gold["overrides"]]
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
        d1 = read_restart(os.path.join(tmp, rst[-1]))
    rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    return d1, rad, r


@pytest.mark.parametrize("name", ["blast16_k6", "blastmix_k5", "blast16hllc_k6", "blast16noh_k6", "blast16ppm_k6"])
def test_dropin_hydro_bit_exact(name):
    case = Case(name)
    d1, _, _ = run_dropin(case)
    assert d1["nstep"] == case.nsteps and d1["time"] == float(case.gold["time1"]) and d1["dt"] == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(d1["grids"][0][v], case.U1[v]), v


@pytest.mark.parametrize("name", ["ifront32_k4", "ioniz24_k3"])
def test_dropin_ion(name):
    case = Case(name)
    d1, rad, r = run_dropin(case)
    assert rad == [int(x) for x in case.gold["rad"][:, 0]]
    assert d1["time"] == float(case.gold["time1"]) and d1["dt"] == float(case.gold["dt1"])
    g = d1["grids"][0]
    for v in case.U1:
        assert np.array_equal(g[v], case.U1[v], equal_nan=True), v
    assert np.array_equal(g["edgeflux"], case.gold["U1_edgeflux"])
    if name.startswith("ioniz"):
        assert "Userwork_in_loop replayed on the device" in r.stdout


# ---- SMR builds (--enable-smr): the reference's init_grid.c builds the overlap tables, the shim hands them to the engine ----
def run_dropin_smr(case, build):
    exe = os.path.join(ROOT, "build", "dropin", build, "athena")
    if not os.path.exists(exe):
        pytest.skip("build/dropin not built")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-i", os.path.join(DECKS, str(case.gold["deck"])), f"time/nlim={case.nsteps}"] + [str(x) for x in case.gold["overrides"]]
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
        d1 = read_restart(os.path.join(tmp, rst[-1]))
    rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    return d1, rad


@pytest.mark.parametrize("name,dropin", [("blastsmr_k4", "blast_smr"), ("blastsmr3_k4", "blast_smr"),
                                         ("ifrontsmr_k3", "ifront_smr"), ("ifrontsmr3_k3", "ifront_smr"),
                                         ("blastsmredge_k4", "blast_smr"), ("ifrontsmredge_k3", "ifront_smr")])
def test_dropin_smr_bit_exact(name, dropin):
    """the unmodified SMR driver linked against the engine, two and three nested levels, against the reference's dumps
    (three levels: its MPI branch as one rank -- the reference's serial build segfaults there)"""
    from smr_cases import SmrCase
    case = SmrCase(name)
    d1, rad = run_dropin_smr(case, dropin)
    assert d1["time"] == float(case.gold["time1"]) and d1["dt"] == float(case.gold["dt1"])
    if case.ion:
        assert rad == [int(x) for x in case.gold["rad"][:, 0]]
    for n, lv in enumerate(case.levels):
        for v in lv["U1"]:
            assert np.array_equal(d1["grids"][n][v], lv["U1"][v]), f"level {n} {v}"
        if case.ion:
            assert np.array_equal(d1["grids"][n]["edgeflux"], lv["edgeflux1"]), f"level {n} EdgeFlux"


def test_dropin_smr_shipped_two_level_geometry_vs_reference_live():
    """SURVEY App. C 'SMR golden': the 2-level cut of the shipped ioniz_sphere_hires deck (root 80^3 + level-1 52^3 at
    displacement 54), static potential + Userwork reset on both levels, 3 steps -- against the reference binary run here."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ioniz_smr", "athena")
    exe = os.path.join(ROOT, "build", "dropin", "ioniz_smr", "athena")
    if not (os.path.exists(ref) and os.path.exists(exe)):
        pytest.skip("needs oracle/_ref/ioniz_smr and build/dropin/ioniz_smr")
    out = {}
    for tag, binary in (("ref", ref), ("b200", exe)):
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [binary, "-i", os.path.join(DECKS, "ioniz_sphere.athinput"), "job/num_domains=2", "time/nlim=3"]
            r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 0, tag + r.stdout[-1500:] + r.stderr[-1500:]
            rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
            out[tag] = (read_restart(os.path.join(tmp, rst[-1])), [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)])
    (a, rad_a), (b, rad_b) = out["ref"], out["b200"]
    assert rad_a == rad_b
    assert a["time"] == b["time"] and a["dt"] == b["dt"]
    for lev in (0, 1):
        for v in ("d", "M1", "M2", "M3", "E", "s", "edgeflux"):
            assert np.array_equal(b["grids"][lev][v], a["grids"][lev][v]), (lev, v)


def test_dropin_shipped_five_level_deck_matches_reference_fingerprint():
    """The SHIPPED 5-level deck (80^3/52^3/96^3/176^3/320^3, 39.8 M zones, static potential, Userwork reset, ion RT on
    every level) for one step through the unmodified driver, against SHA-256 fingerprints of the reference's own dump
    (its MPI branch as one rank; 6 minutes of host time, generated by tests/golden/make_hires5_fingerprint.py)."""
    import importlib.util
    import json
    exe = os.path.join(ROOT, "build", "dropin", "ioniz_smr", "athena")
    if not os.path.exists(exe):
        pytest.skip("needs build/dropin/ioniz_smr")
    spec = importlib.util.spec_from_file_location("fp5", os.path.join(ROOT, "tests", "golden", "make_hires5_fingerprint.py"))
    fp5 = importlib.util.module_from_spec(spec); spec.loader.exec_module(fp5)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "hires5_k1.json")))
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, "-i", os.path.join(DECKS, "ioniz_sphere_hires.athinput"), "time/nlim=1"], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
        d1 = read_restart(os.path.join(tmp, rst[-1]))
    rad = [int(a) for a in re.findall(r"Radiation done in (\d+) iterations", r.stderr)]
    mine = fp5.fingerprint(d1, rad)
    assert mine["rad"] == gold["rad"]
    assert mine["time"] == gold["time"] and mine["dt"] == gold["dt"]
    for lev, (a, b) in enumerate(zip(mine["levels"], gold["levels"])):
        for v in fp5.VARS:
            assert a[v]["shape"] == b[v]["shape"]
            assert a[v]["sha256"] == b[v]["sha256"], f"level {lev} {v}: sum {a[v]['sum']!r} vs {b[v]['sum']!r}, absmax {a[v]['absmax']!r} vs {b[v]['absmax']!r}"


def test_dropin_restart_from_a_reference_dump():
    """`athena -r <dump>` (main.c:388 restart_grids instead of problem()): the drop-in continues a run from a restart
    dump written by the REFERENCE binary -- U and EdgeFlux from the file, the radiator plane and StaticGravPot from
    problem_read_restart (ioniz_sphere.c:200-236), Userwork_in_loop discovered on an evolved state -- and ends where the
    reference's own restarted run ends, bit for bit.  (ifront cannot be restarted in the reference itself: its
    problem_read_restart does not restore the radiator plane and restart.c:284-294 is commented out.  Steps 2-3 only:
    from step 4 on this under-resolved 24^3 planet puts NaN into U, after which the reference's new_dt depends on the
    order it visits the cells.)"""
    case = Case("ioniz24_k3")
    ref = os.path.join(ROOT, "oracle", "_ref", case.build, "athena")
    exe = os.path.join(ROOT, "build", "dropin", case.build, "athena")
    if not (os.path.exists(ref) and os.path.exists(exe)):
        pytest.skip("needs oracle/_ref and build/dropin")
    deck = os.path.join(DECKS, str(case.gold["deck"]))
    ov = [str(x) for x in case.gold["overrides"]]
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b, tempfile.TemporaryDirectory() as c:
        r = subprocess.run([ref, "-i", deck, "time/nlim=1"] + ov, cwd=a, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        dump = os.path.join(a, sorted(f for f in os.listdir(a) if f.endswith(".rst"))[-1])
        out = {}
        for tag, binary, cwd in (("ref", ref, b), ("b200", exe, c)):
            r = subprocess.run([binary, "-r", dump, "time/nlim=3"], cwd=cwd, capture_output=True, text=True)
            assert r.returncode == 0, tag + r.stdout[-1500:] + r.stderr[-1500:]
            rst = sorted(f for f in os.listdir(cwd) if f.endswith(".rst"))
            out[tag] = (read_restart(os.path.join(cwd, rst[-1])), [int(x) for x in re.findall(r"Radiation done in (\d+) iterations", r.stderr)])
    (x, rad_x), (y, rad_y) = out["ref"], out["b200"]
    assert x["nstep"] == y["nstep"] == 3 and rad_x == rad_y and len(rad_x) == 2
    assert x["time"] == y["time"] and x["dt"] == y["dt"]
    for v in x["grids"][0]:
        assert not np.isnan(x["grids"][0][v]).any()
        assert np.array_equal(x["grids"][0][v], y["grids"][0][v]), v


# The reference's own (and only) automated check, SURVEY section 4: linear-wave L1 convergence.  prob/linear_wave3d.c stores
# the eigenmode it starts from and writes the L1 error after 5 periods in Userwork_after_loop.  Values below are what the
# REFERENCE binary (oracle/_ref/linwave_noh: the deck's own configuration, Roe without H-correction, NSCALARS=1) writes
# into LinWave-errors.0.dat for the two grids: Nx1 Nx2 Nx3 RMS d M1 M2 M3 E.
LINWAVE_REFERENCE = {
    16: "32  16  16  4.293960e-05  2.076469e-05  7.037402e-06  1.401739e-05  1.401739e-05  3.114703e-05",
    32: "64  32  32  1.225627e-05  5.928812e-06  1.990201e-06  4.000947e-06  4.000947e-06  8.893219e-06",
}


# the same deck with the reference configured --with-order=3 (oracle/_ref/linwave_ppm_noh)
LINWAVE_REFERENCE_PPM = {
    16: "32  16  16  2.703872e-05  1.306692e-05  4.423587e-06  8.848876e-06  8.848876e-06  1.960037e-05",
    32: "64  32  32  8.017620e-06  3.877901e-06  1.299328e-06  2.619171e-06  2.619171e-06  5.816850e-06",
}


@pytest.mark.parametrize("build,ref", [("linwave_noh", LINWAVE_REFERENCE), ("linwave_ppm_noh", LINWAVE_REFERENCE_PPM)])
def test_dropin_linear_wave_convergence_known_answer(build, ref):
    exe = os.path.join(ROOT, "build", "dropin", build, "athena")
    if not os.path.exists(exe):
        pytest.skip(f"needs build/dropin/{build}")
    rms = {}
    with tempfile.TemporaryDirectory() as tmp:
        for n in (16, 32):
            r = subprocess.run([exe, "-i", os.path.join(DECKS, "linear_wave3d.athinput"), f"domain1/Nx1={2 * n}", f"domain1/Nx2={n}",
                                f"domain1/Nx3={n}"], cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        rows = [ln.split() for ln in open(os.path.join(tmp, "LinWave-errors.0.dat")) if not ln.startswith("#")]
    assert len(rows) == 2
    for row, n in zip(rows, (16, 32)):
        assert row == ref[n].split(), f"N={n}: {row}"       # the very digits the reference prints
        rms[n] = float(row[3])
    assert rms[16] / rms[32] > 3.3, "second-order convergence of the sound wave (the reference reaches 3.4-3.5 on these grids)"


def test_dropin_user_boundary_function_as_device_table():
    """prob/jet.c enrols jet_iib on the ix1 face with bvals_mhd_fun (inflow constants inside the jet radius,
    reflection outside).  The shim turns it into a device table (constants + signed copies) instead of a host round trip
    per call; both paths must reproduce the reference's dump bit for bit."""
    case = Case("jet24_k6")
    for force_host in (False, True):
        if force_host:
            os.environ["AB2_USER_BC"] = "host"
        try:
            d1, _, r = run_dropin(case)
        finally:
            os.environ.pop("AB2_USER_BC", None)
        assert d1["nstep"] == case.nsteps and d1["time"] == float(case.gold["time1"]) and d1["dt"] == float(case.gold["dt1"])
        for v in case.U1:
            assert np.array_equal(d1["grids"][0][v], case.U1[v]), (force_host, v)
        assert ("boundary function of face 0 runs on the device" in r.stdout) == (not force_host), r.stdout[-1500:]


@pytest.mark.parametrize("out", ["cons", "prim"])
@pytest.mark.parametrize("cfg,deck,nlim", [("ifront", "ifront.athinput", 3), ("blast", "blast.athinput", 4)])
def test_dropin_vtk_dumps_are_the_reference_bytes(cfg, deck, nlim, out):
    """dump_vtk's float32 big-endian fields are produced on the device (ab2_output_vtk_field) and written by the
    shim; conserved and primitive dumps must be byte-identical to the files the reference binary writes."""
    ref = os.path.join(ROOT, "oracle", "_ref", cfg, "athena")
    exe = os.path.join(ROOT, "build", "dropin", cfg, "athena")
    if not (os.path.exists(ref) and os.path.exists(exe)):
        pytest.skip("needs oracle/_ref and build/dropin")
    text = open(os.path.join(DECKS, deck)).read()
    assert "maxout      = 1" in text
    text = text.replace("maxout      = 1", "maxout      = 2")
    text += f"\n<output2>\nout_fmt = vtk\nout     = {out}\ndt      = 1.0e30\n"
    files = {}
    for tag, binary in (("ref", ref), ("b200", exe)):
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "deck"), "w").write(text)
            cmd = [binary, "-i", "deck", f"time/nlim={nlim}", "domain1/Nx1=24", "domain1/Nx2=16", "domain1/Nx3=20"]
            r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 0, tag + r.stdout[-1500:] + r.stderr[-1500:]
            files[tag] = {f: open(os.path.join(tmp, f), "rb").read() for f in sorted(os.listdir(tmp)) if f.endswith(".vtk")}
    assert len(files["ref"]) == 2 and sorted(files["ref"]) == sorted(files["b200"]), (sorted(files["ref"]), sorted(files["b200"]))
    for f in files["ref"]:
        assert files["b200"][f] == files["ref"][f], f"{f}: {len(files['b200'][f])} vs {len(files['ref'][f])} bytes"


def test_dropin_forced_host_paths_give_the_same_bits():
    """The opt-outs of INTEGRATION.md ("Problem-file hooks that touch the state"): AB2_USERWORK=host runs the problem file's
    Userwork_in_loop on the host every step (download, hook, upload) instead of the device replay; the dump must not change."""
    case = Case("ioniz24_k3")
    os.environ["AB2_USERWORK"] = "host"
    try:
        d1, rad, r = run_dropin(case)
    finally:
        os.environ.pop("AB2_USERWORK", None)
    assert "Userwork_in_loop replayed on the device" not in r.stdout
    assert rad == [int(x) for x in case.gold["rad"][:, 0]]
    assert d1["time"] == float(case.gold["time1"]) and d1["dt"] == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(d1["grids"][0][v], case.U1[v], equal_nan=True), v
    assert np.array_equal(d1["grids"][0]["edgeflux"], case.gold["U1_edgeflux"])


def test_dropin_replay_survives_its_revalidation_steps():
    """20 steps of ioniz_sphere 24^3 through the drop-in cross the re-validation steps 1, 4 and 16 of the Userwork replay; the
    replay must stay enabled (no fall-back message) and the run must equal the engine's own 20-step run bit for bit."""
    from athena_b200.driver import Run
    case = Case("ioniz24_k3")
    exe = os.path.join(ROOT, "build", "dropin", case.build, "athena")
    if not os.path.exists(exe):
        pytest.skip("build/dropin not built")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-i", os.path.join(DECKS, str(case.gold["deck"])), "time/nlim=20"] + [str(x) for x in case.gold["overrides"]]
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rst = sorted(f for f in os.listdir(tmp) if f.endswith(".rst"))
        d1 = read_restart(os.path.join(tmp, rst[-1]))
    assert "Userwork_in_loop replayed on the device" in r.stdout and "no longer matches" not in r.stdout
    R = Run(case.par, case.problem, device=0, U0=case.U0)
    R.run(20)
    out = R.G.download_active()
    assert d1["nstep"] == 20 and d1["time"] == R.time
    for v in out:
        assert np.array_equal(d1["grids"][0][v], out[v], equal_nan=True), v
