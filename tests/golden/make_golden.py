#!/usr/bin/env python3
"""Generate the golden parity fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Runs the unmodified reference binaries built by oracle/build_ref.py (oracle/_ref/<cfg>/athena) on the
decks in tests/decks/ with small grids, and stores the t=0 and t=K restart dumps (FP64, active zones)
plus the per-cycle `next dt` / radiation sub-cycle counts it printed.  Also stores function-level
vectors from the reference's own lr_states()/fluxes() objects (oracle/_ref/<cfg>/libref_units.so).

    python tests/golden/make_golden.py          # needs /root/reference (to build oracle/_ref) -- container only

The committed .npz files are what the GPU box sees; nothing under /root/reference is read at test time.
"""
import ctypes as C
import importlib.util
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
from athena_b200.restart import read_restart  # noqa: E402

DECKS = os.path.join(ROOT, "tests", "decks")
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name -> (reference build, deck, overrides, [nlim...])
CASES = {
    "blast16": ("blast", "blast.athinput",
                ["domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25"], [1, 6]),
    "blastmix": ("blast", "blast.athinput",
                 ["domain1/Nx1=20", "domain1/Nx2=12", "domain1/Nx3=16", "problem/radius=0.2",
                  "domain1/x2min=-0.3", "domain1/x2max=0.3", "domain1/x3min=-0.4", "domain1/x3max=0.4",
                  "domain1/bc_ix1=1", "domain1/bc_ox1=2", "domain1/bc_ix2=2", "domain1/bc_ox2=1",
                  "domain1/bc_ix3=1", "domain1/bc_ox3=2"], [5]),
    # smallest Grids the 3D CTU path accepts (4 zones = one ghost width per direction), all three extents different
    "blasttiny": ("blast", "blast.athinput",
                  ["domain1/Nx1=8", "domain1/Nx2=4", "domain1/Nx3=6", "problem/radius=0.3"], [5]),
    "ifronttiny": ("ifront", "ifront.athinput",
                   ["domain1/Nx1=5", "domain1/Nx2=4", "domain1/Nx3=7"], [3]),
    # static potential + Userwork reset on a Grid with three different, partly odd extents (the potential tables cannot all alias)
    "ionizodd": ("ioniz_sphere", "ioniz_sphere.athinput",
                 ["domain1/Nx1=21", "domain1/Nx2=18", "domain1/Nx3=15"], [1]),
    # a problem-file boundary function (prob/jet.c jet_iib on ix1: inflow constants inside the jet radius, reflection outside)
    "jet24": ("jet", "jet.athinput", [], [6]),
    "ifront32": ("ifront", "ifront.athinput",
                 ["domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16"], [1, 4]),
    "ioniz24": ("ioniz_sphere", "ioniz_sphere.athinput",
                ["domain1/Nx1=24", "domain1/Nx2=24", "domain1/Nx3=24"], [1, 3]),
    # two-level SMR (root + one level-1 Domain, both single Grid); builds configured with --enable-smr
    "blastsmr": ("blast_smr", "blast.athinput",
                 ["job/num_domains=2", "domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25",
                  "domain2/Nx1=16", "domain2/Nx2=12", "domain2/Nx3=8", "domain2/iDisp=8", "domain2/jDisp=10", "domain2/kDisp=12"], [1, 4]),
    "ifrontsmr": ("ifront_smr", "ifront.athinput",
                  ["job/num_domains=2", "domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16",
                   "domain2/Nx1=24", "domain2/Nx2=16", "domain2/Nx3=16", "domain2/iDisp=8", "domain2/jDisp=8", "domain2/kDisp=8"], [1, 3]),
    # the reference configured --with-flux=hllc (no H-correction: etah is Roe-only)
    "blast16hllc": ("blast_hllc", "blast.athinput",
                    ["domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25"], [6]),
    # --with-order=3 (PPM, characteristic variables)
    "blast16ppm": ("blast_ppm", "blast.athinput",
                   ["domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25"], [6]),
    "ifront32ppm": ("ifront_ppm", "ifront.athinput",
                    ["domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16"], [4]),
    "blast16ppmhllc": ("blast_ppm_hllc", "blast.athinput",
                       ["domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25"], [6]),
    "blast16noh": ("blast_noh", "blast.athinput",
                   ["domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25"], [6]),
    "ifront32hllc": ("ifront_hllc", "ifront.athinput",
                     ["domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16"], [4]),
    # a child Domain that touches root boundaries: its x1 faces get the root's physical BCs (reflect / outflow) instead of
    # prolongation, the rays enter it through the root boundary itself, and two of its faces need no flux correction
    "ifrontsmredge": ("ifront_smr", "ifront.athinput",
                      ["job/num_domains=2", "domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16", "problem/n_H=0.05",
                       "domain2/Nx1=24", "domain2/Nx2=16", "domain2/Nx3=20", "domain2/iDisp=0", "domain2/jDisp=16", "domain2/kDisp=8"], [3]),
    "blastsmredge": ("blast_smr", "blast.athinput",
                     ["job/num_domains=2", "domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25",
                      "domain1/bc_ix1=2", "domain1/bc_ox1=2", "domain1/bc_ix2=1", "domain1/bc_ox2=2", "domain1/bc_ix3=2", "domain1/bc_ox3=1",
                      "domain2/Nx1=16", "domain2/Nx2=12", "domain2/Nx3=16", "domain2/iDisp=16", "domain2/jDisp=0", "domain2/kDisp=16"], [4]),
    # three nested levels.  The reference's serial build segfaults below level 1 with ion RT, so these
    # come from its MPI branch run as one rank against oracle/mpi_stub (builds *_smr_mpi).  n_H lowered so the box
    # is optically thin-ish (tau ~ 0.3 per root cell) and non-trivial fluxes are handed down to both fine levels.
    "ifrontsmr3": ("ifront_smr_mpi", "ifront.athinput",
                   ["job/num_domains=3", "domain1/Nx1=32", "domain1/Nx2=16", "domain1/Nx3=16", "problem/n_H=0.05",
                    "domain2/Nx1=24", "domain2/Nx2=16", "domain2/Nx3=16", "domain2/iDisp=8", "domain2/jDisp=8", "domain2/kDisp=8",
                    "domain3/Nx1=16", "domain3/Nx2=16", "domain3/Nx3=16", "domain3/iDisp=24", "domain3/jDisp=24", "domain3/kDisp=24"], [1, 3]),
    "blastsmr3": ("blast_smr_mpi", "blast.athinput",
                  ["job/num_domains=3", "domain1/Nx1=16", "domain1/Nx2=16", "domain1/Nx3=16", "problem/radius=0.25",
                   "domain2/Nx1=16", "domain2/Nx2=12", "domain2/Nx3=16", "domain2/iDisp=8", "domain2/jDisp=10", "domain2/kDisp=8",
                   "domain3/Nx1=12", "domain3/Nx2=8", "domain3/Nx3=12", "domain3/iDisp=28", "domain3/jDisp=28", "domain3/kDisp=24"], [4]),
}


def run_reference(build, deck, overrides, nlim):
    exe = os.path.join(REFDIR, build, "athena")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [exe, "-i", os.path.join(DECKS, deck), f"time/nlim={nlim}"] + overrides
        r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, check=True)
        out = os.path.join(tmp, "id0") if build.endswith("_mpi") else tmp        # MPI builds write into id<rank>/
        rst = sorted(f for f in os.listdir(out) if f.endswith(".rst"))
        assert len(rst) == 2, rst
        d0 = read_restart(os.path.join(out, rst[0]))
        d1 = read_restart(os.path.join(out, rst[1]))
    next_dt = [float(x) for x in re.findall(r"cycle=\d+ time=\S+ next dt=(\S+)", r.stdout)]
    rad = [(int(a), int(b), int(c), float(d)) for a, b, c, d in re.findall(
        r"Radiation done in (\d+) iterations: (\d+) thermal, (\d+) chemical; new dt = (\S+)", r.stderr)]
    return d0, d1, next_dt, rad, cmd


def pack(prefix, g, out):
    for k, v in g.items():
        out[f"{prefix}_{k}"] = v


def make_runs():
    for name, (build, deck, ov, nlims) in CASES.items():
        for nlim in nlims:
            d0, d1, next_dt, rad, cmd = run_reference(build, deck, ov, nlim)
            out = {}
            pack("U0", d0["grids"][0], out)
            pack("U1", d1["grids"][0], out)
            for lev in range(1, len(d0["grids"])):
                pack(f"U0L{lev}", d0["grids"][lev], out)
                pack(f"U1L{lev}", d1["grids"][lev], out)
            out["dt0"] = d0["dt"]; out["time0"] = d0["time"]
            out["dt1"] = d1["dt"]; out["time1"] = d1["time"]; out["nstep1"] = d1["nstep"]
            out["next_dt"] = np.array(next_dt)
            out["rad"] = np.array(rad, dtype=np.float64).reshape(-1, 4)
            out["deck"] = deck
            out["overrides"] = np.array(ov)
            out["build"] = build
            path = os.path.join(HERE, f"{name}_k{nlim}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path) // 1024, "KiB;", " ".join(cmd[3:]))


def make_units(seed=20261018):
    """Function-level vectors: reference lr_states()/fluxes()/Cons1D_to_Prim1D() on seeded random states."""
    dp = C.POINTER(C.c_double)
    P = lambda a: a.ctypes.data_as(dp)
    for cfg, nscal in (("ifront", 1), ("blast", 0), ("ifront_hllc", 1), ("blast_hllc", 0), ("ifront_ppm", 1), ("blast_ppm", 0),
                       ("blast_ppm_hllc", 0)):
        hllc = cfg.endswith("_hllc")
        tag = {"ifront": "", "blast": "", "ifront_hllc": "hllc_", "blast_hllc": "hllc_", "ifront_ppm": "ppm_", "blast_ppm": "ppm_",
               "blast_ppm_hllc": "ppm_hllc_"}[cfg]
        ref = C.CDLL(os.path.join(REFDIR, cfg, "libref_units.so"))
        ref.ref_set_gamma.argtypes = [C.c_double]
        ref.ref_fluxes.argtypes = [dp, dp, dp, dp, C.c_double, dp]
        ref.ref_prim_to_cons.argtypes = [dp, dp]
        ref.ref_cons_to_prim.argtypes = [dp, dp]
        ref.ref_cfast.argtypes = [dp]; ref.ref_cfast.restype = C.c_double
        ref.ref_lr_states.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        gam = 5.0 / 3.0
        ref.ref_set_gamma(gam)
        rng = np.random.default_rng(seed + nscal)
        nf = 4096
        Wl = np.empty((nf, 6)); Wr = np.empty((nf, 6))
        for W in (Wl, Wr):
            W[:, 0] = rng.uniform(0.1, 2, nf); W[:, 1] = 2 * rng.normal(size=nf)
            W[:, 2:4] = rng.normal(size=(nf, 2)); W[:, 4] = rng.uniform(0.01, 3, nf)
            W[:, 5] = rng.uniform(0, 1, nf) if nscal else 0.0
        Wr[::3, 0] *= 1e-3; Wr[::3, 4] *= 1e-4          # strong rarefactions: exercises the HLLE fallback
        Wr[1::7, 1] += 6.0; Wl[1::7, 1] += 6.0          # supersonic: upwind shortcut
        eta = np.where(np.arange(nf) % 2 == 0, rng.uniform(0, 2, nf), 0.0)
        Ul = np.empty((nf, 6)); Ur = np.empty((nf, 6)); F = np.empty((nf, 6)); Wb = np.empty((nf, 6))
        cf = np.empty(nf)
        for t in range(nf):
            ref.ref_prim_to_cons(P(Wl[t]), P(Ul[t])); ref.ref_prim_to_cons(P(Wr[t]), P(Ur[t]))
            ref.ref_fluxes(P(Ul[t]), P(Ur[t]), P(Wl[t]), P(Wr[t]), eta[t], P(F[t]))
            ref.ref_cons_to_prim(P(Ul[t]), P(Wb[t]))
            cf[t] = ref.ref_cfast(P(Ul[t]))
        npen, n = 64, 48
        W = np.empty((npen, n, 6)); WL = np.zeros((npen, n, 6)); WR = np.zeros((npen, n, 6))
        for t in range(npen):
            w = W[t]
            w[:, 0] = rng.uniform(0.1, 2, n); w[:, 1:4] = rng.normal(size=(n, 3)); w[:, 4] = rng.uniform(0.01, 3, n)
            w[:, 5] = rng.uniform(0, 1, n) if nscal else 0.0
            if t % 2:
                w[:] = w[0] + np.cumsum(0.02 * w, axis=0)
                w[:, 0] = np.abs(w[:, 0]) + 0.1; w[:, 4] = np.abs(w[:, 4]) + 0.01
            ref.ref_lr_states(P(w), n, 0.01, 0.05, 3, n - 4, P(WL[t]), P(WR[t]))
        path = os.path.join(HERE, f"units_{tag}nscal{nscal}.npz")
        if tag:         # a variant build sees the same seeded inputs: keep only what it computes differently (cases.load_units merges)
            np.savez_compressed(path, F=F, pencil_Wl=WL, pencil_Wr=WR)
        else:
            np.savez_compressed(path, gamma=gam, nscal=nscal, Wl=Wl, Wr=Wr, Ul=Ul, Ur=Ur, eta=eta, F=F, W_of_Ul=Wb, cfast_Ul=cf,
                                pencil_W=W, pencil_Wl=WL, pencil_Wr=WR, pencil_dt=0.01, pencil_dx=0.05, pencil_il=3, pencil_iu=n - 4)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    br = importlib.util.module_from_spec(spec); spec.loader.exec_module(br)
    if os.path.isdir(os.path.join(br.REF, "src")):
        for c in ("ifront", "ioniz_sphere", "blast", "ifront_smr", "blast_smr", "ifront_smr_mpi", "blast_smr_mpi",
                  "blast_hllc", "ifront_hllc", "blast_noh", "blast_ppm", "ifront_ppm", "blast_ppm_hllc"):
            br.build(c)
        for c in ("ifront", "blast", "ifront_hllc", "blast_hllc", "ifront_ppm", "blast_ppm", "blast_ppm_hllc"):
            br.build_units(c)
    only = sys.argv[1:]
    if only:
        CASES = {k: v for k, v in CASES.items() if k in only}
        globals()["CASES"] = CASES
    make_runs()
    if not only or "units" in only:
        make_units()
