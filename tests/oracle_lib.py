"""ctypes binding of oracle/_build/libathena_oracle.so + a restatement of the reference's main loop
(main.c:519-669) on top of it.  TEST INFRASTRUCTURE: imported only by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NG = 4
dp = C.POINTER(C.c_double)
GRAVFN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_double)


class OrcGrid(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int), ("nscal", C.c_int),
                ("dx1", C.c_double), ("dx2", C.c_double), ("dx3", C.c_double),
                ("x1min", C.c_double), ("x2min", C.c_double), ("x3min", C.c_double),
                ("gamma", C.c_double),
                ("d", dp), ("M1", dp), ("M2", dp), ("M3", dp), ("E", dp), ("s", dp),
                ("grav", C.c_void_p), ("Fsave", dp * 3)]


class OrcOverlap(C.Structure):
    _fields_ = [("cs", C.c_int * 3), ("ce", C.c_int * 3), ("face", C.c_int * 6), ("cdisp", C.c_int * 3)]


class OrcIonPar(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("sigma_ph", "m_H", "mu", "e_gamma", "alpha_C", "k_B", "time_unit", "max_de_iter",
                 "max_de_therm_iter", "max_dx_iter", "max_de_step", "max_de_therm_step", "max_dx_step",
                 "tfloor", "tceil")] + [("maxiter", C.c_int), ("min_area", C.c_double), ("d_nlo", C.c_double)]


def build_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libathena_oracle.so")
    src = os.path.join(ROOT, "oracle", "athena_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py"), "--oracle-only"], check=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        gp = C.POINTER(OrcGrid)
        L.orc_bvals.argtypes = [gp, C.POINTER(C.c_int)]
        L.orc_max_speeds.argtypes = [gp, dp]
        L.orc_new_dt.argtypes = [gp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
        L.orc_new_dt.restype = C.c_double
        L.orc_integrate_3d_ctu.argtypes = [gp, C.c_double]
        L.orc_lr_states.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        L.orc_roe_flux.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, dp]
        L.orc_cons_to_prim.argtypes = [dp, C.c_int, C.c_double, dp]
        L.orc_prim_to_cons.argtypes = [dp, C.c_int, C.c_double, dp]
        L.orc_bvals_ionrad.argtypes = [gp, C.c_double, dp]
        L.orc_roe_flux_batch.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp, dp]
        L.orc_hllc_flux_batch.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp]
        L.orc_set_riemann_solver.argtypes = [C.c_int]
        L.orc_set_reconstruction.argtypes = [C.c_int]
        L.orc_prim_to_cons_batch.argtypes = [C.c_int, dp, C.c_int, C.c_double, dp]
        L.orc_cons_to_prim_batch.argtypes = [C.c_int, dp, C.c_int, C.c_double, dp]
        L.orc_ion_derived.argtypes = [C.POINTER(OrcIonPar), C.c_double, C.c_double, C.c_double]
        L.orc_ion_entering_flux.argtypes = [C.c_double, C.c_double]
        L.orc_ion_entering_flux.restype = C.c_double
        L.orc_ion_radtransfer.argtypes = [gp, C.POINTER(OrcIonPar), C.c_double, C.c_double, C.c_double,
                                          dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_ion_radtransfer.restype = C.c_int
        L.orc_ion_radtransfer_root.argtypes = L.orc_ion_radtransfer.argtypes + [dp]
        L.orc_ion_radtransfer_root.restype = C.c_int
        op = C.POINTER(OrcOverlap)
        L.orc_restrict_correct.argtypes = [gp, gp, op, C.c_double]
        L.orc_ion_restrict.argtypes = [gp, gp, op]
        L.orc_prolongate.argtypes = [gp, gp, op]
        L.orc_ion_prolong.argtypes = [gp, dp, gp, dp, op]
        L.orc_ion_radtransfer_fine.argtypes = [gp, C.POINTER(OrcIonPar), C.c_double, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_ion_radtransfer_fine.restype = C.c_int
        _lib = L
    return _lib


def P(a):
    return a.ctypes.data_as(dp)


VARS = ("d", "M1", "M2", "M3", "E", "s")


class OracleGrid:
    """SoA state with ghosts + the scalars the reference keeps in MeshS/GridS."""

    def __init__(self, nx, xmin, xmax, gamma, nscal, bc, grav=None, dx=None, keep_fluxes=False):
        self.nx = tuple(int(n) for n in nx)                     # (Nx1,Nx2,Nx3)
        self.dx = tuple(dx) if dx is not None else tuple((xmax[a] - xmin[a]) / self.nx[a] for a in range(3))   # init_mesh.c: dx = (max-min)/Nx
        self.xmin = tuple(xmin)
        self.gamma, self.nscal = float(gamma), int(nscal)
        self.bc = (C.c_int * 6)(*bc)
        shp = (self.nx[2] + 2 * NG, self.nx[1] + 2 * NG, self.nx[0] + 2 * NG)
        self.U = {v: np.zeros(shp) for v in VARS[:5 + self.nscal]}
        self.grav = grav                                        # ctypes function pointer or None
        self.g = OrcGrid(self.nx[0], self.nx[1], self.nx[2], self.nscal, *self.dx, *self.xmin, self.gamma,
                         *[P(self.U[v]) if v in self.U else None for v in VARS],
                         C.cast(grav, C.c_void_p) if grav is not None else None)
        self.edgeflux = np.zeros((self.nx[2] + 1, self.nx[1] + 1, self.nx[0] + 1))
        self.time, self.dt, self.nstep = 0.0, 0.0, 0
        self.Fsave = None
        if keep_fluxes:                                         # SMR: myFlx source, zero before the first step (calloc)
            self.Fsave = [np.zeros(shp + (6,)) for _ in range(3)]
            for a in range(3):
                self.g.Fsave[a] = P(self.Fsave[a])

    def active(self, v):
        return self.U[v][NG:-NG, NG:-NG, NG:-NG]

    def set_active(self, d):
        for v in self.U:
            self.active(v)[...] = d[v]

    def bvals(self):
        lib().orc_bvals(C.byref(self.g), self.bc)

    def bvals_ionrad(self, flux_i):
        lib().orc_bvals_ionrad(C.byref(self.g), flux_i, P(self.edgeflux))

    def new_dt(self, cour_no, tlim):
        self.dt = lib().orc_new_dt(C.byref(self.g), cour_no, self.dt, self.nstep, self.time, tlim)
        return self.dt

    def integrate(self):
        lib().orc_integrate_3d_ctu(C.byref(self.g), self.dt)

    def ion_radtransfer(self, ionpar, flux_i, cour_no):
        dt = C.c_double(self.dt)
        nchem, ntherm, tc = C.c_int(0), C.c_int(0), C.c_double(0.0)
        niter = lib().orc_ion_radtransfer_root(C.byref(self.g), C.byref(ionpar), flux_i, self.time, cour_no,
                                               C.byref(dt), P(self.edgeflux), C.byref(nchem), C.byref(ntherm), C.byref(tc))
        self.dt = dt.value
        self.tcoarse = tc.value
        return niter, ntherm.value, nchem.value

    def ion_radtransfer_fine(self, ionpar, tcoarse):
        dt = C.c_double(self.dt)
        nchem, ntherm = C.c_int(0), C.c_int(0)
        niter = lib().orc_ion_radtransfer_fine(C.byref(self.g), C.byref(ionpar), tcoarse, C.byref(dt), P(self.edgeflux),
                                               C.byref(nchem), C.byref(ntherm))
        self.dt = dt.value
        return niter, ntherm.value, nchem.value

    def max_speeds(self):
        v = (C.c_double * 3)()
        lib().orc_max_speeds(C.byref(self.g), v)
        return tuple(v)


def ion_par_from(par, dx):
    p = OrcIonPar()
    for n, _ in OrcIonPar._fields_[:15]:
        setattr(p, n, par.par_getd("ionradiation", n))
    p.maxiter = int(par.par_getd("ionradiation", "maxiter"))   # ionrad_3d.c:757 assigns a double to an int
    lib().orc_ion_derived(C.byref(p), *dx)
    return p


def run_main_loop(G, nsteps, cour_no, tlim, ionpar=None, flux_i=0.0, userwork=None, log=None):
    """main.c:519-669 for one root Grid: [ion RT; bvals] -> integrate -> Userwork_in_loop -> time += dt ->
    new_dt -> bvals.  G must already hold problem() data, ghost zones and the first dt (main.c:430-451)."""
    for _ in range(nsteps):
        if G.time >= tlim:
            break
        if ionpar is not None:
            rad = G.ion_radtransfer(ionpar, flux_i, cour_no)
            G.bvals()
            if log is not None:
                log.setdefault("rad", []).append(rad + (G.dt,))
        G.integrate()
        if userwork is not None:
            userwork(G)
        G.nstep += 1
        G.time += G.dt
        G.new_dt(cour_no, tlim)
        G.bvals()
        if log is not None:
            log.setdefault("next_dt", []).append(G.dt)


def make_overlap(parent_nx, child_nx, child_disp, root_nx, level=1, parent_disp=(0, 0, 0)):
    """CGrid.ijks/ijke and the myFlx != NULL pattern of init_grid.c:395-660 for a single-Grid parent Domain with
    displacement parent_disp (parent-level cells; 0 for the root) and a single-Grid child Domain at `level` with
    displacement child_disp (child-level cells, from the root origin)."""
    o = OrcOverlap()
    ref = 2 ** level
    for a in range(3):
        o.cs[a] = child_disp[a] // 2 - parent_disp[a] + NG
        o.ce[a] = (child_disp[a] + child_nx[a]) // 2 - parent_disp[a] + NG - 1
        o.cdisp[a] = child_disp[a]
        o.face[2 * a] = 1 if child_disp[a] != 0 else 0
        o.face[2 * a + 1] = 1 if (child_disp[a] + child_nx[a]) // ref != root_nx[a] else 0
    return o


def run_main_loop_smr(root, child, o, nsteps, cour_no, tlim, ionpar=None, flux_i=0.0, log=None):
    """Two-level form of run_main_loop_levels (root Grid + one level-1 Grid)."""
    return run_main_loop_levels([root, child], [o], nsteps, cour_no, tlim, ionpar, flux_i, log)


def run_main_loop_levels(grids, ovs, nsteps, cour_no, tlim, ionpar=None, flux_i=0.0, log=None):
    """main.c:519-669 with STATIC_MESH_REFINEMENT for a nest of single-Grid Domains, grids[n] at level n and
    ovs[n] the overlap of grids[n+1] on grids[n].  All must hold problem() data, ghost zones (bvals + Prolongate)
    and the first dt.  Mesh scalars live on the root (time, nstep)."""
    L = lib()
    root = grids[0]
    nlev = len(grids)
    mesh_dt = root.dt
    for _ in range(nsteps):
        if root.time >= tlim:
            break
        if ionpar is not None:
            for n, G in enumerate(grids):                        # main.c:546-556, coarse to fine
                if n == 0:
                    r = G.ion_radtransfer(ionpar, flux_i, cour_no)
                else:                                            # ionrad_3d.c:880 ionrad_prolong_rcv from the parent's :1037 snd
                    pa = grids[n - 1]
                    L.orc_ion_prolong(C.byref(pa.g), P(pa.edgeflux), C.byref(G.g), P(G.edgeflux), C.byref(ovs[n - 1]))
                    r = G.ion_radtransfer_fine(ionpar, root.tcoarse)
                mesh_dt = G.dt                                   # ionrad_3d.c:1033 pMesh->dt = pGrid->dt
                G.bvals()
                if log is not None:
                    log.setdefault("rad", []).append(r + (G.dt,))
            for n in range(nlev - 2, -1, -1):                    # ionradRestrictCorrect, finest pair first (smr.c:111)
                L.orc_ion_restrict(C.byref(grids[n].g), C.byref(grids[n + 1].g), C.byref(ovs[n]))
        for G in grids:
            G.integrate()
        for n in range(nlev - 2, -1, -1):                        # RestrictCorrect (smr.c:1233), finest pair first
            L.orc_restrict_correct(C.byref(grids[n].g), C.byref(grids[n + 1].g), C.byref(ovs[n]), grids[n].dt)
        root.nstep += 1
        root.time += mesh_dt
        for G in grids[1:]:
            G.time = root.time
        mesh_dt = new_dt_smr(grids, cour_no, mesh_dt, root.nstep, root.time, tlim)
        for G in grids:
            G.bvals()
        for n in range(nlev - 1):                                # Prolongate (smr.c:2386), root first
            L.orc_prolongate(C.byref(grids[n].g), C.byref(grids[n + 1].g), C.byref(ovs[n]))
        if log is not None:
            log.setdefault("next_dt", []).append(mesh_dt)
    return mesh_dt


def new_dt_smr(grids, cour_no, dt_old, nstep, time, tlim):
    """new_dt.c:58-195 over all levels: one dt for the whole Mesh."""
    max_dti = 0.0
    max_v = [0.0, 0.0, 0.0]          # new_dt.c:53: never reset between Grids -> running maximum over the levels so far
    for G in grids:
        v = G.max_speeds()
        for a in range(3):
            if G.nx[a] > 1:
                max_v[a] = max_v[a] if max_v[a] > v[a] else v[a]
                x = max_v[a] / G.dx[a]
                max_dti = max_dti if max_dti > x else x
    dt = cour_no / max_dti if nstep == 0 else min(2.0 * dt_old, cour_no / max_dti)
    if time < tlim and (tlim - time) < dt:
        dt = tlim - time
    for G in grids:
        G.dt = dt
    return dt
