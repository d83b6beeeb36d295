"""Shared set-up for parity tests: turns a golden fixture (tests/golden/*.npz) into grid geometry, boundary
flags, ion parameters and problem hooks, for the oracle and for the CUDA engine alike."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from athena_b200.athinput import ParFile  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DECKS = os.path.join(ROOT, "tests", "decks")
NG = 4
VARS = ("d", "M1", "M2", "M3", "E", "s")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def load_units(tag, nscal):
    """Function-level vectors of a reference build: the seeded inputs live in units_nscal<N>.npz, a variant build's file
    (tag "hllc_", "ppm_", "ppm_hllc_") holds only the outputs that differ (F, pencil_Wl, pencil_Wr)."""
    g = dict(load_golden(f"units_nscal{nscal}"))
    if tag:
        g.update(dict(load_golden(f"units_{tag}nscal{nscal}")))
    return g


def par_for(gold):
    par = ParFile.par_open(os.path.join(DECKS, str(gold["deck"])))
    par.par_cmdline([str(x) for x in gold["overrides"]])
    return par


class Case:
    def __init__(self, name):
        self.name = name
        self.gold = load_golden(name)
        self.par = par_for(self.gold)
        p = self.par
        self.build = str(self.gold["build"])
        # 0 Roe + H-correction, 1 HLLC (reference --with-flux=hllc), 2 Roe built without --enable-h-correction
        self.riemann = 1 if self.build.endswith("_hllc") else (2 if self.build.endswith("_noh") else 0)
        self.order = 3 if "_ppm" in self.build else 2                    # lr_states_plm.c / lr_states_ppm.c (--with-order)
        self.problem = self.build.replace("_hllc", "").replace("_noh", "").replace("_ppm", "")
        self.nx = tuple(p.par_geti("domain1", f"Nx{a}") for a in (1, 2, 3))
        self.xmin = tuple(p.par_getd("domain1", f"x{a}min") for a in (1, 2, 3))
        self.xmax = tuple(p.par_getd("domain1", f"x{a}max") for a in (1, 2, 3))
        self.bc = tuple(p.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3"))
        self.gamma = p.par_getd("problem", "gamma")
        self.cour_no = p.par_getd("time", "cour_no")
        self.tlim = p.par_getd("time", "tlim")
        self.ion = self.problem in ("ifront", "ioniz_sphere")
        self.nscal = 1 if self.ion else 0
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.nsteps = int(self.gold["nstep1"])
        self.U0 = {v: self.gold[f"U0_{v}"] for v in VARS[:5 + self.nscal]}
        self.U1 = {v: self.gold[f"U1_{v}"] for v in VARS[:5 + self.nscal]}
        self.has_grav = self.problem == "ioniz_sphere"

    # ---- problem hooks (host C, athena_b200/csrc/problems.c) ------------------------------------
    def problem_lib(self):
        from athena_b200 import problems
        return problems.lib()

    def grav_fn(self):
        """StaticGravPot as a C function pointer (ioniz_sphere.c: PlanetPot), or None."""
        if not self.has_grav:
            return None
        from athena_b200 import problems
        problems.ioniz_sphere_setup(self.par)
        return problems.lib().ab2_planet_pot

    def reset_table(self):
        """Userwork_in_loop of ioniz_sphere.c as (flat cell indices incl. ghosts, 6 x n values), or None."""
        if not self.has_grav:
            return None
        from athena_b200 import problems
        problems.ioniz_sphere_setup(self.par)
        dx = tuple((self.xmax[a] - self.xmin[a]) / self.nx[a] for a in range(3))
        return problems.ioniz_sphere_reset_table(self.nx, self.xmin, dx)
