"""Parity tests proper: the CUDA engine, called through the C ABI (libathena_b200.so), against
 (1) the committed golden vectors produced by the reference itself (tests/golden), and
 (2) the CPU oracle run live on the same inputs.
Bars (BASELINE.json north_star): hydro -- bit-exact (the library is built --fmad=false and keeps the reference's
operation order); ion RT -- per-cell relative error <= 1e-10 after one step, identical sub-cycle iteration counts."""
import ctypes as C

import os

import numpy as np
import pytest

from cases import Case, load_golden, load_units, VARS

pytestmark = pytest.mark.gpu

ION_RTOL_1STEP = 1e-10      # north_star: per-cell relative error <= 1e-10 after one step
ION_RTOL_KSTEP = 1e-9       # a few steps (the 100-step bar is volume-L1 <= 1e-8)


def engine():
    from athena_b200 import engine as e
    return e


def P(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("nscal", [0, 1])
def test_units_bit_exact(nscal):
    """fluxes(), lr_states(), Cons1D_to_Prim1D, Prim1D_to_Cons1D, cfast vs the reference's own objects."""
    L = engine().units()
    g = load_golden(f"units_nscal{nscal}")
    gam = float(g["gamma"])
    c = lambda k: np.ascontiguousarray(g[k])
    n = len(g["Wl"])
    F = np.empty((n, 6)); U = np.empty((n, 6)); W = np.empty((n, 6)); cf = np.empty(n)
    assert L.ab2_unit_convert(n, P(c("Wl")), nscal, gam, 0, P(U), None) == 0
    assert np.array_equal(U, g["Ul"])
    assert L.ab2_unit_roe_flux(n, P(c("Ul")), P(c("Ur")), P(c("Wl")), P(c("Wr")), nscal, gam, P(c("eta")), P(F)) == 0
    bad = np.argwhere(F != g["F"])
    assert len(bad) == 0, (bad[:5], F[bad[0][0]], g["F"][bad[0][0]])
    assert L.ab2_unit_convert(n, P(c("Ul")), nscal, gam, 1, P(W), P(cf)) == 0
    assert np.array_equal(W, g["W_of_Ul"])
    assert np.array_equal(cf, g["cfast_Ul"])
    Wp = c("pencil_W")
    npen, npc, _ = Wp.shape
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    Wl = np.zeros_like(Wp); Wr = np.zeros_like(Wp)
    assert L.ab2_unit_lr_states(npen, npc, P(Wp), nscal, gam, float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, P(Wl), P(Wr)) == 0
    assert np.array_equal(Wl[:, il:iu + 2], g["pencil_Wl"][:, il:iu + 2])
    assert np.array_equal(Wr[:, il - 1:iu + 1], g["pencil_Wr"][:, il - 1:iu + 1])


def test_fast_ieee_division_and_sqrt_are_correctly_rounded():
    """Arith<true> (branch-free MUFU + FMA refinement) must return the bits of a/b and sqrt(a) for every operand its
    range test accepts; rejected operands go through the plain operators.  4e6 operand pairs over 600 binades plus
    the special values."""
    L = engine().units()
    rng = np.random.default_rng(7)
    n = 1 << 22
    a = rng.uniform(1, 2, n) * 2.0 ** rng.integers(-300, 300, n) * rng.choice([-1.0, 1.0], n)
    b = rng.uniform(1, 2, n) * 2.0 ** rng.integers(-300, 300, n) * rng.choice([-1.0, 1.0], n)
    sp = np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1e308, 1.7e308, 1e-20, 2.2250738585072014e-308])
    a[:len(sp) ** 2] = np.repeat(sp, len(sp)); b[:len(sp) ** 2] = np.tile(sp, len(sp))
    a[1000:2000] = rng.uniform(1, 2, 1000) * 2.0 ** rng.integers(-1074, -960, 1000)       # tiny numerators / radicands
    b[2000:3000] = rng.uniform(1, 2, 1000) * 2.0 ** rng.integers(1000, 1023, 1000)        # huge divisors
    qf, qe, sf, se = (np.empty(n) for _ in range(4))
    nrej = (C.c_int * 2)()
    assert L.ab2_unit_ieee(n, P(a), P(b), P(qf), P(qe), P(sf), P(se), nrej) == 0
    assert np.array_equal(qf.view(np.int64), qe.view(np.int64)) or np.array_equal(qf, qe, equal_nan=True)
    assert np.array_equal(sf, se, equal_nan=True)
    host_q = a / b
    ok = np.isfinite(host_q)
    assert np.array_equal(qe[ok], host_q[ok])                  # and the device operator agrees with the host's IEEE division
    assert nrej[0] < n // 100 and nrej[1] > n // 3             # divisions: few rejected; square roots: the negative half is


def test_glibc_math_matches_host_libm_bit_for_bit():
    """The ion kernels' exp()/pow() (csrc/glibc_math.cuh) against the host libm the reference links, on the arguments
    the path produces: exp(-tau), exp(-118348/T), pow(T/1e4,-0.7), pow(T,-0.89), plus wide random ranges."""
    import ctypes.util
    import math
    L = engine().units()
    rng = np.random.default_rng(11)
    n = 1 << 21
    x = np.concatenate([-10.0 ** rng.uniform(-12, 3.2, n // 4), rng.uniform(-745, 709, n // 4),
                        np.full(n // 4, -0.7), np.full(n // 4, -0.89)])
    y = np.concatenate([rng.uniform(0.5, 2, n // 4), 10.0 ** rng.uniform(-300, 300, n // 4),
                        10.0 ** rng.uniform(-4, 6, n // 4), 10.0 ** rng.uniform(0, 9, n // 4)])
    x[:8] = [0.0, -0.0, 1e-300, -1e-300, -745.2, -800.0, 710.0, -1e-17]
    e = np.empty(n); p = np.empty(n)
    assert L.ab2_unit_libm(n, P(x), P(y), P(e), P(p)) == 0
    libm = C.CDLL(ctypes.util.find_library("m"))          # the host libm itself: no Python range errors in the way
    libm.exp.restype = C.c_double; libm.exp.argtypes = [C.c_double]
    libm.pow.restype = C.c_double; libm.pow.argtypes = [C.c_double, C.c_double]
    he = np.array([libm.exp(v) for v in x[:n // 2]])
    assert np.array_equal(e[:n // 2], he), f"exp differs in {(e[:n//2] != he).sum()} of {n//2}"
    for sel in (np.r_[n // 2:n], np.r_[0:n // 4]):
        hp = np.array([libm.pow(b, a) for b, a in zip(y[sel], x[sel])])
        bad = p[sel] != hp
        assert not bad.any(), f"pow differs in {bad.sum()} of {len(sel)}: {y[sel][bad][:4]} ^ {x[sel][bad][:4]}"


def gpu_run(case, fuse=None, planes=None):
    """fuse: 0 forces the unfused k_correct / k_final pair (AB2_HYDRO_FUSE, read at Grid creation); planes: x3-chunked scratch"""
    from athena_b200.driver import Run
    if fuse is not None:
        os.environ["AB2_HYDRO_FUSE"] = str(fuse)
    try:
        R = Run(case.par, case.problem, device=0, U0=case.U0, riemann=case.riemann, order=case.order)
    finally:
        os.environ.pop("AB2_HYDRO_FUSE", None)
    if planes is not None:
        R.G.set_scratch_planes(planes)
        assert R.G.scratch_planes == (planes if planes < case.nx[2] else 0)
    assert R.dt == float(case.gold["dt0"]), "first new_dt must be bit-identical (max is order independent)"
    R.run(case.nsteps)
    return R


@pytest.mark.parametrize("name", ["blast16_k6", "blastmix_k5", "blasttiny_k5", "blast16hllc_k6", "blast16ppm_k6", "ifront32_k4",
                                  "ioniz24_k3", "ionizodd_k1", "ifronttiny_k3"])
@pytest.mark.parametrize("variant", ["fused", "chunk4_fused", "chunk8", "chunk5", "pairwise_correct"])
def test_hydro_pass_variants_bit_exact_vs_reference(name, variant):
    """The step's pass structure does not show in the results: the fused correct+eta+final kernel (k_cfinal: corrected states
    and eta in shared memory only), pass C as one or three launches, and the x3-chunked scratch (chunks of 4, 5 or 8 active
    planes, incl. a ragged last chunk) reproduce the reference's dump bit for bit, exactly like the default path of the
    tests below (k_correct3 / k_final pair, whole-Grid scratch)."""
    case = Case(name)
    fuse = 1 if "fused" in variant else 0
    planes = {"fused": None, "chunk4_fused": 4, "chunk8": 8, "chunk5": 5, "pairwise_correct": None}[variant]
    if variant == "pairwise_correct":
        os.environ["AB2_CORRECT3"] = "0"          # pass C as three launches, one per face direction
    try:
        R = gpu_run(case, fuse=fuse, planes=planes)
    finally:
        os.environ.pop("AB2_CORRECT3", None)
    out = R.G.download_active()
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(out[v], case.U1[v], equal_nan=True), f"{name}/{variant}:{v} differs in {(out[v] != case.U1[v]).sum()} cells"


def rel_err(a, b):
    scale = np.maximum(np.abs(b), np.abs(b).max() * 1e-30 + 1e-300)
    return np.abs(a - b) / scale


@pytest.mark.parametrize("nscal", [0, 1])
def test_hllc_units_bit_exact(nscal):
    """rsolvers/hllc.c fluxes() and the HLL-family form of PLM steps 8-9, against the reference's own objects"""
    g = load_units("hllc_", nscal)
    L = engine().units()
    gam = float(g["gamma"])
    n = len(g["Wl"])
    c = lambda k: np.ascontiguousarray(g[k])
    F = np.empty((n, 6))
    assert L.ab2_unit_hllc_flux(n, P(c("Ul")), P(c("Ur")), P(c("Wl")), P(c("Wr")), nscal, gam, P(F)) == 0
    assert np.array_equal(F, g["F"])
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    Wp = c("pencil_W")
    npen, m = Wp.shape[:2]
    Wl = np.zeros_like(Wp); Wr = np.zeros_like(Wp)
    assert L.ab2_unit_lr_states_hll(npen, m, P(Wp), nscal, gam, float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, P(Wl), P(Wr)) == 0
    assert np.array_equal(Wl[:, il:iu + 2], g["pencil_Wl"][:, il:iu + 2])
    assert np.array_equal(Wr[:, il - 1:iu + 1], g["pencil_Wr"][:, il - 1:iu + 1])


@pytest.mark.parametrize("tag,nscal,hll", [("ppm_", 0, 0), ("ppm_", 1, 0), ("ppm_hllc_", 0, 1)])
def test_ppm_units_bit_exact(tag, nscal, hll):
    """reconstruction/lr_states_ppm.c (the reference's object built with THIRD_ORDER_CHAR) on the seeded pencils, incl. the
    passive scalar as the reference's aliased work arrays treat it, and the HLL-family form of the tracing"""
    g = load_units(tag, nscal)
    L = engine().units()
    il, iu = int(g["pencil_il"]), int(g["pencil_iu"])
    Wp = np.ascontiguousarray(g["pencil_W"])
    npen, m = Wp.shape[:2]
    Wl = np.zeros_like(Wp); Wr = np.zeros_like(Wp)
    assert L.ab2_unit_lr_states_ppm(npen, m, P(Wp), nscal, float(g["gamma"]), float(g["pencil_dt"]), float(g["pencil_dx"]), il, iu, hll, P(Wl), P(Wr)) == 0
    assert np.array_equal(Wl[:, il:iu + 2], g["pencil_Wl"][:, il:iu + 2])
    assert np.array_equal(Wr[:, il - 1:iu + 1], g["pencil_Wr"][:, il - 1:iu + 1])


@pytest.mark.parametrize("name", ["blast16_k1", "blast16_k6", "blastmix_k5", "blast16hllc_k6", "blast16noh_k6", "blast16ppm_k6",
                                  "blast16ppmhllc_k6", "blasttiny_k5"])
def test_hydro_bit_exact_vs_reference(name):
    case = Case(name)
    R = gpu_run(case)
    out = R.G.download_active()
    assert R.nstep == case.nsteps
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(out[v], case.U1[v]), f"{name}:{v} max abs diff {np.abs(out[v]-case.U1[v]).max()}"


@pytest.mark.parametrize("name", ["ifront32_k1", "ifront32_k4", "ioniz24_k1", "ioniz24_k3", "ifront32hllc_k4", "ifront32ppm_k4",
                                  "ifronttiny_k3", "ionizodd_k1"])
@pytest.mark.parametrize("pair", [False, True])
def test_ion_hydro_bit_exact_vs_reference(name, pair):
    """stricter than north_star asks: with the host libm's exp/pow restated on the device (csrc/glibc_math.cuh) the
    hydro + ion-RT step reproduces the reference's dump to the last bit, EdgeFlux and sub-cycle counts included --
    through the fused update+rates kernel Grids this small take by default, and (pair) through the rates / update pair with
    its sparse sub-cycle (k_ion_walk and the task-list passes; rows shorter than one 32-cell chunk, odd extents)"""
    case = Case(name)
    if pair:
        os.environ["AB2_ION_FUSE"] = "0"
    try:
        R = gpu_run(case)
    finally:
        os.environ.pop("AB2_ION_FUSE", None)
    out = R.G.download_active()
    assert [tuple(r[:3]) for r in R.rad_log] == [(int(a), int(b), int(c)) for a, b, c in case.gold["rad"][:, :3]]
    assert R.time == float(case.gold["time1"]) and R.dt == float(case.gold["dt1"])
    for v in case.U1:
        assert np.array_equal(out[v], case.U1[v], equal_nan=True), f"{name}:{v} differs in {(out[v] != case.U1[v]).sum()} cells"
    assert np.array_equal(R.G.download_edgeflux(), case.gold["U1_edgeflux"])


@pytest.mark.parametrize("name,rtol", [("ifront32_k1", ION_RTOL_1STEP), ("ifront32_k4", ION_RTOL_KSTEP),
                                       ("ioniz24_k1", ION_RTOL_1STEP), ("ioniz24_k3", ION_RTOL_KSTEP)])
def test_ion_hydro_vs_reference(name, rtol):
    case = Case(name)
    R = gpu_run(case)
    out = R.G.download_active()
    rad = case.gold["rad"]
    assert [r[0] for r in R.rad_log] == [int(x) for x in rad[:, 0]], "sub-cycle iteration counts (ionrad_3d.c:1041)"
    assert [r[1] for r in R.rad_log] == [int(x) for x in rad[:, 1]]
    assert R.nstep == case.nsteps
    assert abs(R.time - float(case.gold["time1"])) <= rtol * abs(float(case.gold["time1"]))
    assert abs(R.dt - float(case.gold["dt1"])) <= rtol * abs(float(case.gold["dt1"]))
    for v in ("d", "E", "s"):
        e = rel_err(out[v], case.U1[v]).max()
        assert e <= rtol, f"{name}:{v} max rel err {e:.3e}"
    # momenta pass through zero: compare against the momentum scale d*cs of the cell
    cs = np.sqrt(case.gamma * (case.gamma - 1) * np.abs(case.U1["E"]) / case.U1["d"])
    for v in ("M1", "M2", "M3"):
        e = (np.abs(out[v] - case.U1[v]) / (case.U1["d"] * cs)).max()
        assert e <= rtol, f"{name}:{v} max err / (d cs) {e:.3e}"
    ef = R.G.download_edgeflux()
    gef = case.gold["U1_edgeflux"]
    assert np.array_equal(ef == 0.0, gef == 0.0), "ray cut (1e-3) must fall in the same cells"
    assert rel_err(ef, gef).max() <= max(rtol, 1e-9)
    # ionisation front position per ray: first i with s/d < 0.5 -- within one cell (north_star)
    front = lambda U: np.argmax((U["s"] / U["d"]) >= 0.5, axis=2)
    assert np.abs(front(out) - front(case.U1)).max() <= 1


@pytest.mark.parametrize("name", ["blast16_k1", "ifront32_k1", "ioniz24_k1"])
def test_pieces_vs_oracle(name):
    """bvals_mhd / new_dt / integrate_3d_ctu one call at a time against the oracle on the same state."""
    import oracle_lib as ol
    from athena_b200 import engine as eng
    case = Case(name)
    grav = case.grav_fn()
    O = ol.OracleGrid(case.nx, case.xmin, case.xmax, case.gamma, case.nscal, case.bc, grav=grav)
    O.set_active(case.U1)                          # a developed state
    G = eng.Grid(case.nx, O.dx, case.xmin, case.gamma, case.nscal)
    G.upload({v: O.U[v] for v in O.U})
    O.bvals(); G.bvals_mhd(case.bc)
    got = G.download()
    for v in O.U:
        assert np.array_equal(got[v], O.U[v]), f"bvals {v}"
    vo = (C.c_double * 3)(); ol.lib().orc_max_speeds(C.byref(O.g), vo)
    assert tuple(vo) == G.max_wave_speeds()
    O.nstep = 1; O.dt = 1e300
    dt = O.new_dt(case.cour_no, case.tlim)
    assert dt == G.new_dt(case.cour_no, 1e300, 1, 0.0, case.tlim)
    if grav is not None:
        G.set_static_grav_pot(grav)
    O.integrate(); G.integrate_3d_ctu(dt)
    got = G.download_active()
    for v in O.U:
        # equal_nan: a full-CFL hydro step from a post-ionisation state can produce NaN through cfast() (no floor,
        # convert_var.c:486-491); the engine must then produce NaN in the same cells (ternary MAX/MIN semantics)
        assert np.array_equal(got[v], O.active(v), equal_nan=True), f"integrate {v}: max abs {np.nanmax(np.abs(got[v]-O.active(v)))}"


def test_aos_round_trip_and_layout():
    """ConsS AoS <-> device SoA (coherence points of SURVEY 8b) and ghost-zone layout."""
    from athena_b200 import engine as eng
    rng = np.random.default_rng(3)
    G = eng.Grid((12, 8, 6), (0.1, 0.1, 0.1), (0, 0, 0), 5 / 3, 1)
    A = rng.normal(size=G.shape + (6,))
    G.upload_aos(A)
    S = G.download()
    for n, v in enumerate(VARS):
        assert np.array_equal(S[v], A[..., n])
    assert np.array_equal(G.download_aos(), A)


def test_no_device_memory_growth_and_launch_count():
    case = Case("blast16_k1")
    R = gpu_run(case)
    b0, l0 = R.G.bytes_allocated, R.G.kernel_launches
    R.run(case.nsteps + 3)
    assert R.G.bytes_allocated == b0
    assert R.G.kernel_launches > l0


def test_device_side_writer_fields_and_bc_table():
    """ab2_output_vtk_field / ab2_output_active_var against numpy on the downloaded state; ab2_set_bc_table against the built-in
    reflecting boundary (the same ghost cells written as a signed-copy table)."""
    from athena_b200 import engine as eng
    rng = np.random.default_rng(11)
    G = eng.Grid((12, 8, 6), (0.1, 0.1, 0.1), (0, 0, 0), 5 / 3, 1)
    U = {v: rng.uniform(0.5, 2.0, size=G.shape) for v in VARS}
    U["E"] += 5.0
    G.upload(U)
    act = {v: U[v][4:-4, 4:-4, 4:-4] for v in VARS}
    be = lambda a: np.ascontiguousarray(a, dtype=np.float64).astype(np.float32).astype(">f4").view(np.uint8).ravel()
    assert np.array_equal(G.output_vtk_field(0), be(act["d"]))
    assert np.array_equal(G.output_vtk_field(1), be(np.stack([act["M1"], act["M2"], act["M3"]], axis=-1)))
    assert np.array_equal(G.output_vtk_field(2), be(act["E"]))
    assert np.array_equal(G.output_vtk_field(3), be(act["s"]))
    di = 1.0 / act["d"]
    assert np.array_equal(G.output_vtk_field(1, prim=True), be(np.stack([act["M1"] * di, act["M2"] * di, act["M3"] * di], axis=-1)))
    P = (act["E"] - 0.5 * (act["M1"] * act["M1"] + act["M2"] * act["M2"] + act["M3"] * act["M3"]) * di) * (5 / 3 - 1.0)
    P = np.maximum(P, 1.0e-20)            # convert_var.c:409 MAX(P, TINY_NUMBER): some of the random cells have KE > E
    assert np.array_equal(G.output_vtk_field(2, prim=True), be(P))
    for n, v in enumerate(VARS):
        assert np.array_equal(G.output_active_var(n), act[v])
    # reflecting ix1 as a table: ghost (k,j,4-l) <- active (k,j,4+l-1), M1 negated, j,k active (bvals_mhd.c:959-977)
    n3, n2, n1 = G.shape
    dst, src = [], []
    for k in range(4, n3 - 4):
        for j in range(4, n2 - 4):
            for l in range(1, 5):
                dst.append((k * n2 + j) * n1 + 4 - l); src.append((k * n2 + j) * n1 + 4 + l - 1)
    nv = 6
    S = np.tile(np.array(src), (nv, 1)); Cf = np.ones((nv, len(dst))); Cf[1] = -1.0
    G.set_bc_table(0, dst, S, Cf)
    G.bvals_mhd([-2, 0, 0, 0, 0, 0])
    a = G.download()
    G.upload(U)
    G.bvals_mhd([1, 0, 0, 0, 0, 0])
    b = G.download()
    for v in VARS:
        assert np.array_equal(a[v], b[v]), v
    # constants
    Cf2 = rng.normal(size=(nv, len(dst)))
    G.set_bc_table(0, dst, -np.ones_like(S), Cf2)
    G.bvals_mhd([-2, 0, 0, 0, 0, 0])
    a = G.download()
    for n, v in enumerate(VARS):
        assert np.array_equal(a[v].ravel()[np.array(dst)], Cf2[n])
