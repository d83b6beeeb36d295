"""Python handle on the host-side C problem generators / user hooks (csrc/problems.c).

Mirrors what the reference's problem files give the driver: problem() initial data, the StaticGravPot
function pointer and Userwork_in_loop (src/prob/{blast,ifront,ioniz_sphere}.c).  Host code only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "problems.c")
_SO = os.path.join(_HERE, "csrc", "libathena_problems.so")
NG = 4
dp = C.POINTER(C.c_double)


class ProbGrid(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int),
                ("x1min", C.c_double), ("x2min", C.c_double), ("x3min", C.c_double),
                ("dx1", C.c_double), ("dx2", C.c_double), ("dx3", C.c_double),
                ("d", dp), ("M1", dp), ("M2", dp), ("M3", dp), ("E", dp), ("s", dp)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", _SRC, "-o", _SO, "-lm"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ab2_planet_pot.argtypes = [C.c_double] * 3
        L.ab2_planet_pot.restype = C.c_double
        L.ab2_ioniz_sphere_setup.argtypes = [C.c_double] * 6
        L.ab2_problem_blast.argtypes = [C.POINTER(ProbGrid)] + [C.c_double] * 6
        L.ab2_problem_ifront.argtypes = [C.POINTER(ProbGrid)] + [C.c_double] * 4
        L.ab2_problem_ioniz_sphere.argtypes = [C.POINTER(ProbGrid)]
        L.ab2_ioniz_sphere_reset_table.argtypes = [C.POINTER(ProbGrid), C.POINTER(C.c_long), dp, C.c_long]
        L.ab2_ioniz_sphere_reset_table.restype = C.c_long
        _lib = L
    return _lib


def _pg(nx, xmin, dx, U=None):
    ptr = lambda v: U[v].ctypes.data_as(dp) if (U is not None and v in U) else None
    return ProbGrid(nx[0], nx[1], nx[2], xmin[0], xmin[1], xmin[2], dx[0], dx[1], dx[2],
                    ptr("d"), ptr("M1"), ptr("M2"), ptr("M3"), ptr("E"), ptr("s"))


def alloc_state(nx, nscal):
    shp = (nx[2] + 2 * NG, nx[1] + 2 * NG, nx[0] + 2 * NG)
    return {v: np.zeros(shp) for v in ("d", "M1", "M2", "M3", "E", "s")[:5 + nscal]}


def ioniz_sphere_setup(par):
    lib().ab2_ioniz_sphere_setup(par.par_getd("problem", "gamma"), par.par_getd("ionradiation", "mu"),
                                 par.par_getd("problem", "cs"), par.par_getd_def("problem", "rp", 1.2e10),
                                 par.par_getd_def("problem", "mp", 1.0e30), par.par_getd_def("problem", "np", 6.0e8))


def ioniz_sphere_reset_table(nx, xmin, dx):
    g = _pg(nx, xmin, dx)
    n = lib().ab2_ioniz_sphere_reset_table(C.byref(g), None, None, 0)
    cells = np.zeros(max(n, 1), dtype=np.int64)
    vals = np.zeros((6, max(n, 1)))
    lib().ab2_ioniz_sphere_reset_table(C.byref(g), cells.ctypes.data_as(C.POINTER(C.c_long)), vals.ctypes.data_as(dp), max(n, 1))
    return cells[:n], np.ascontiguousarray(vals[:, :n])


def problem(name, par, nx, xmin, dx):
    """problem(DomainS*) of the named problem file: returns SoA state dict (with ghosts, ghosts unset)."""
    nscal = 0 if name == "blast" else 1
    U = alloc_state(nx, nscal)
    g = _pg(nx, xmin, dx, U)
    gamma = par.par_getd("problem", "gamma")
    if name == "blast":
        lib().ab2_problem_blast(C.byref(g), gamma, par.par_getd("problem", "pamb"), par.par_getd("problem", "prat"),
                                par.par_getd_def("problem", "damb", 1.0), par.par_getd_def("problem", "drat", 1.0),
                                par.par_getd("problem", "radius"))
    elif name == "ifront":
        lib().ab2_problem_ifront(C.byref(g), gamma, par.par_getd("problem", "n_H"),
                                 par.par_getd("ionradiation", "m_H"), par.par_getd("problem", "cs"))
    elif name == "ioniz_sphere":
        ioniz_sphere_setup(par)
        lib().ab2_problem_ioniz_sphere(C.byref(g))
    else:
        raise ValueError(f"unknown problem '{name}'")
    return U
