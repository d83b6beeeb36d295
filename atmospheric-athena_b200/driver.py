"""Host driver: the reference's main.c time loop (main.c:355-669) over one root-level Domain whose Grids are
x3 slabs, one per GPU / process.

The reference keeps this in C and so does the drop-in (shim/athena_shim.c); this Python mirror exists so that
tests and bench.py can drive libathena_b200.so through the same call sequence -- bvals_mhd, bvals_ionrad, new_dt,
then per step [ion_radtransfer_3d; bvals_mhd] -> integrate_3d_ctu -> Userwork_in_loop -> new_dt -> bvals_mhd --
with the same scalars (MeshS.time/dt/nstep, GridS.dt).  It does no arithmetic on the state itself.

Decomposition: rays run along x1 (dir = -1), so Grids are cut along x3 only: NGrid_x1 = NGrid_x2 = 1,
NGrid_x3 = world size.  Ghost slabs at x3 faces are contiguous per variable in the SoA layout and are exchanged
in place with torch.distributed point-to-point (NCCL over NVLink on GPUs, gloo in the CPU tests); timestep and
sub-cycle scalars go through all_reduce, standing where the reference calls MPI (bvals_mhd.c:423-446,
new_dt.c:175-179, ionrad_3d.c:275,399,554,672).
"""
from __future__ import annotations

import numpy as np

from . import engine, problems

NG = 4
VARS = engine.VARS


def slab_bounds(nx3: int, world: int, rank: int):
    """k-range [k0,k1) of rank's slab; the reference requires Nx3 % NGrid_x3 == 0 (init_mesh.c dom_decomp)."""
    if nx3 % world != 0:
        raise ValueError(f"Nx3={nx3} is not divisible by the number of x3 slabs {world}")
    n = nx3 // world
    return rank * n, (rank + 1) * n


def exchange_x3_ghosts(U, rank: int, world: int, bc_ix3: int, bc_ox3: int, dist=None, group=None):
    """In-place exchange of the 4 ghost layers at the x3 faces shared with neighbouring slabs.

    U: torch tensor [nvar][n3][n2][n1] (any device).  Lower neighbour's top active layers land in U[:,0:4], upper
    neighbour's bottom active layers in U[:,n3-4:n3]; with periodic x3 (flag 4) the slabs form a ring
    (bvals_mhd.c:768-774, 814-820).  Slabs span all of x1,x2 incl. ghosts, so corners are filled by the ordering
    x1 -> x2 -> x3 exactly as in bvals_mhd.c:170-172."""
    if world == 1:
        return
    n3 = U.shape[1]
    lo = rank - 1 if rank > 0 else (world - 1 if bc_ix3 == 4 else None)
    hi = rank + 1 if rank < world - 1 else (0 if bc_ox3 == 4 else None)
    ops = []
    nvar = U.shape[0]
    # Upward transfers first (send top active layers to hi / receive bottom ghosts from lo), then downward ones.
    # With a 2-slab periodic ring both messages travel between the same pair of ranks and NCCL matches
    # point-to-point calls by order, so every rank must post them in the same global order.
    for v in range(nvar):
        if hi is not None:
            ops.append(dist.P2POp(dist.isend, U[v, n3 - 2 * NG:n3 - NG], hi, group))
        if lo is not None:
            ops.append(dist.P2POp(dist.irecv, U[v, 0:NG], lo, group))
    for v in range(nvar):
        if lo is not None:
            ops.append(dist.P2POp(dist.isend, U[v, NG:2 * NG], lo, group))
        if hi is not None:
            ops.append(dist.P2POp(dist.irecv, U[v, n3 - NG:n3], hi, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class Run:
    """MeshS + one DomainS + this rank's GridS, driven like main.c."""

    def __init__(self, par, problem: str, device: int = 0, dist=None, U0: dict | None = None, torch_memory: bool | None = None,
                 riemann: int = 0):
        self.par = par
        self.problem = problem
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        p = par
        self.ion = problem in ("ifront", "ioniz_sphere")
        self.nscal = 1 if self.ion else 0
        self.root_nx = tuple(p.par_geti("domain1", f"Nx{a}") for a in (1, 2, 3))
        xmin = [p.par_getd("domain1", f"x{a}min") for a in (1, 2, 3)]
        xmax = [p.par_getd("domain1", f"x{a}max") for a in (1, 2, 3)]
        self.dx = tuple((xmax[a] - xmin[a]) / self.root_nx[a] for a in range(3))        # init_mesh.c: dx = (max-min)/Nx
        self.bc = [p.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3")]
        self.gamma = p.par_getd("problem", "gamma")
        self.cour_no = p.par_getd("time", "cour_no")
        self.tlim = p.par_getd("time", "tlim")
        k0, k1 = slab_bounds(self.root_nx[2], self.world, self.rank)
        self.k0, self.k1 = k0, k1
        self.nx = (self.root_nx[0], self.root_nx[1], k1 - k0)
        # GridS.MinX of the slab: init_grid.c forms MinX = root MinX + Disp*dx
        self.xmin = (xmin[0], xmin[1], xmin[2] + k0 * self.dx[2])
        # faces owned by a neighbouring slab are skipped by bvals and filled by the exchange
        self.bc_local = list(self.bc)
        if self.world > 1:
            if self.rank > 0 or self.bc[4] == 4:
                self.bc_local[4] = 0
            if self.rank < self.world - 1 or self.bc[5] == 4:
                self.bc_local[5] = 0
        self.G = engine.Grid(self.nx, self.dx, self.xmin, self.gamma, self.nscal, device=device,
                             torch_memory=(self.world > 1) if torch_memory is None else torch_memory)
        if riemann:
            self.G.set_riemann_solver(riemann)       # the reference's --with-flux=hllc build
        self.time, self.dt, self.nstep = 0.0, 0.0, 0
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.rad_log = []
        self._reduce = self._make_reduce() if self.world > 1 else None

        # --- problem() (main.c:391-395)
        if U0 is None:
            U = problems.problem(problem, p, self.nx, self.xmin, self.dx)
            self.G.upload(U)
        else:
            self.G.upload_active({v: U0[v][k0:k1] for v in VARS[:5 + self.nscal]})
        if problem == "ioniz_sphere":
            problems.ioniz_sphere_setup(p)
            self.G.set_static_grav_pot(problems.lib().ab2_planet_pot)        # StaticGravPot = PlanetPot
            cells, vals = problems.ioniz_sphere_reset_table(self.nx, self.xmin, self.dx)
            self.G.set_fixed_cells(cells, vals)                               # Userwork_in_loop replay
            self.userwork = self.G.apply_fixed_cells
        else:
            self.userwork = None
        if self.ion:
            self.G.ion_init(engine.ion_par_from(p, self.dx))                  # main.c:473
        # --- main.c:430-451
        self.bvals_mhd()
        if self.ion:
            self.G.bvals_ionrad(self.flux_i)
        self.new_dt()

    # ---- pieces -----------------------------------------------------------------------------------
    def _make_reduce(self):
        import torch
        dist = self.dist
        ops = {0: dist.ReduceOp.MIN, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.SUM}
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        world = dist.get_world_size()
        buf = torch.zeros(4, dtype=torch.float64, device=dev)
        gat = torch.zeros(4 * world, dtype=torch.float64, device=dev)
        host = torch.zeros(4, dtype=torch.float64).pin_memory() if dev == "cuda" else torch.zeros(4, dtype=torch.float64)

        def reduce(vals, op):
            n = len(vals)
            host[:n] = torch.from_numpy(vals)
            buf[:n].copy_(host[:n], non_blocking=True)
            if op == 4:          # {SUM, MAX, MIN, MIN} in one exchange: gather the quadruples, combine locally
                dist.all_gather_into_tensor(gat, buf)
                g = gat.cpu().view(world, 4)
                return np.array([float(g[:, 0].sum()), float(g[:, 1].max()), float(g[:, 2].min()), float(g[:, 3].min())])
            dist.all_reduce(buf[:n], op=ops[op])
            return buf[:n].cpu().numpy()
        return reduce

    def bvals_mhd(self):
        self.G.bvals_mhd(self.bc_local)
        if self.world > 1:
            exchange_x3_ghosts(self.G.U_torch, self.rank, self.world, self.bc[4], self.bc[5], self.dist)

    def new_dt(self):
        self.dt = self.G.new_dt(self.cour_no, self.dt, self.nstep, self.time, self.tlim, self._reduce)
        return self.dt

    def step(self):
        """One pass of the main loop body (main.c:519-669)."""
        if self.ion:
            self.dt, niter, ntherm, nchem = self.G.ion_radtransfer_3d(self.flux_i, self.time, self.cour_no, self.dt, self._reduce)
            self.rad_log.append((niter, ntherm, nchem, self.dt))
            self.bvals_mhd()
        self.G.integrate_3d_ctu(self.dt)
        if self.userwork is not None:
            self.userwork()
        self.nstep += 1
        self.time += self.dt
        self.new_dt()
        self.bvals_mhd()

    def run(self, nlim: int):
        while self.time < self.tlim and self.nstep < nlim:
            self.step()

    @property
    def zones(self):
        return self.nx[0] * self.nx[1] * self.nx[2]


class SmrRun:
    """main.c:355-669 with STATIC_MESH_REFINEMENT for a nest of single-Grid Domains on one GPU (root + one Domain per
    finer level, `job/num_domains` of them): RestrictCorrect / Prolongate / ionradRestrictCorrect / ionrad_prolong
    between the device mirrors of adjacent levels, one dt for the whole Mesh (new_dt.c:58-195), every fine level's
    radiation sub-cycle running to the root's coarse time."""

    def __init__(self, par, problem: str, U0: dict, *U0_fine: dict, device: int = 0):
        p = par
        self.ion = problem in ("ifront", "ioniz_sphere")
        self.nscal = 1 if self.ion else 0
        nlev = 1 + len(U0_fine)
        nx0 = tuple(p.par_geti("domain1", f"Nx{a}") for a in (1, 2, 3))
        xmin = [p.par_getd("domain1", f"x{a}min") for a in (1, 2, 3)]
        xmax = [p.par_getd("domain1", f"x{a}max") for a in (1, 2, 3)]
        dx0 = tuple((xmax[a] - xmin[a]) / nx0[a] for a in range(3))
        bc0 = [p.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3")]
        self.gamma = p.par_getd("problem", "gamma")
        self.cour_no = p.par_getd("time", "cour_no")
        self.tlim = p.par_getd("time", "tlim")
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.grids, self.bcs, self.ovs, disps = [], [], [], []
        for n in range(nlev):
            blk = f"domain{n + 1}"
            nx = tuple(p.par_geti(blk, f"Nx{a}") for a in (1, 2, 3))
            disp = (0, 0, 0) if n == 0 else tuple(p.par_geti(blk, k) for k in ("iDisp", "jDisp", "kDisp"))
            dx = tuple(dx0[a] / 2.0 ** n for a in range(3))
            xm = tuple(xmin[a] if disp[a] == 0 else xmin[a] + float(disp[a]) * dx[a] for a in range(3))   # init_mesh.c:377-382
            bc = []             # interior faces of a child are prolongated, not BC'd (bvals_mhd.c:549-550 ProlongateLater)
            for a in range(3):
                bc.append(bc0[2 * a] if disp[a] == 0 else 0)
                bc.append(bc0[2 * a + 1] if (disp[a] + nx[a]) // 2 ** n == nx0[a] else 0)
            if n > 0:
                self.ovs.append(engine.make_overlap(nx, disp, nx0, level=n, parent_disp=disps[-1]))
            G = engine.Grid(nx, dx, xm, self.gamma, self.nscal, device=device)
            G.upload_active(U0 if n == 0 else U0_fine[n - 1])
            self.grids.append(G); self.bcs.append(bc); disps.append(disp)
        if self.ion:
            ip = engine.ion_par_from(p, dx0)                    # min_area, d_nlo from the ROOT spacing (ionrad.c:115-131)
            for G in self.grids:
                G.ion_init(ip)
        self.root = self.grids[0]
        self.child = self.grids[1] if nlev > 1 else None
        self.time, self.dt, self.nstep = 0.0, 0.0, 0
        self.gdt = [0.0] * nlev                                 # GridS.dt per level
        self.rad_log = []
        # main.c:400-451
        for n in range(nlev - 2, -1, -1):
            self.grids[n].smr_restrict_correct(self.grids[n + 1], self.ovs[n], 0.0)
        for G, bc in zip(self.grids, self.bcs):
            G.bvals_mhd(bc)
        if self.ion:                                            # bvals_ionrad.c:27-29,178: the first Domain's BC function wins
            for G in self.grids:
                G.bvals_ionrad(self.flux_i)
        for n in range(nlev - 1):
            self.grids[n].smr_prolongate(self.grids[n + 1], self.ovs[n])
        self.new_dt()

    def new_dt(self):
        max_dti = 0.0
        max_v = [0.0, 0.0, 0.0]      # new_dt.c:53: never reset between Grids -> running maximum over the levels so far
        for G in self.grids:
            v = G.max_wave_speeds()
            for a in range(3):
                if G.nx[a] > 1:
                    max_v[a] = max_v[a] if max_v[a] > v[a] else v[a]
                    x = max_v[a] / G.dx[a]
                    max_dti = max_dti if max_dti > x else x
        dt = self.cour_no / max_dti if self.nstep == 0 else min(2.0 * self.dt, self.cour_no / max_dti)
        if self.time < self.tlim and (self.tlim - self.time) < dt:
            dt = self.tlim - self.time
        self.dt = dt
        self.gdt = [dt] * len(self.grids)
        return dt

    def step(self):
        g, nlev = self.grids, len(self.grids)
        if self.ion:
            for n in range(nlev):                               # main.c:546-556, coarse to fine
                if n == 0:
                    self.gdt[0], ni, nt, nc = g[0].ion_radtransfer_3d(self.flux_i, self.time, self.cour_no, self.gdt[0])
                else:
                    g[n - 1].smr_ion_prolong(g[n], self.ovs[n - 1])
                    self.gdt[n], ni, nt, nc = g[n].ion_radtransfer_3d_fine(g[0].ion_coarse_time, self.gdt[n])
                self.rad_log.append((ni, nt, nc, self.gdt[n]))
                self.dt = self.gdt[n]                           # ionrad_3d.c:1033
                g[n].bvals_mhd(self.bcs[n])
            for n in range(nlev - 2, -1, -1):                   # ionradRestrictCorrect, finest pair first
                g[n].smr_ion_restrict(g[n + 1], self.ovs[n])
        for n in range(nlev):
            g[n].integrate_3d_ctu(self.gdt[n])
        for n in range(nlev - 2, -1, -1):                       # RestrictCorrect, finest pair first
            g[n].smr_restrict_correct(g[n + 1], self.ovs[n], self.gdt[n])
        self.nstep += 1
        self.time += self.dt
        self.new_dt()
        for n in range(nlev):
            g[n].bvals_mhd(self.bcs[n])
        for n in range(nlev - 1):                               # Prolongate, root first
            g[n].smr_prolongate(g[n + 1], self.ovs[n])

    def run(self, nlim):
        while self.time < self.tlim and self.nstep < nlim:
            self.step()
