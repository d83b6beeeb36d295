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
                 riemann: int = 0, order: int = 2):
        self.par = par
        self.problem = problem
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        p = par
        if p.par_geti_def("job", "num_domains", 1) != 1:
            raise ValueError("driver.Run drives one root Domain; use SmrRun / SmrRunSlabs for job/num_domains > 1")
        self.nscal = 1 if problem in ("ifront", "ioniz_sphere") else 0      # an ion BUILD carries the scalar either way
        # main.c:546 runs the radiation step only when a plane exists; both problem files add theirs iff problem/nradplanes > 0
        self.ion = self.nscal == 1 and p.par_geti("problem", "nradplanes") > 0
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
        # More than one GPU: NCCL lives INSIDE the library (csrc/comm.cu) -- ab2_bvals_mhd exchanges the x3 ghost slabs and
        # the dt / sub-cycle reductions span the communicator; torch.distributed only carries the unique id.  With a CPU
        # backend (gloo: the host-logic tests, where a CPU stand-in sits behind engine.Grid) the same exchange and reductions
        # go through torch.distributed below.
        use_comm = self.world > 1 and dist.get_backend() == "nccl"
        self.G = engine.Grid(self.nx, self.dx, self.xmin, self.gamma, self.nscal, device=device,
                             torch_memory=(self.world > 1 and not use_comm) if torch_memory is None else torch_memory)
        self.comm = None
        if use_comm:
            lo = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.bc[4] == 4 else None)
            hi = self.rank + 1 if self.rank < self.world - 1 else (0 if self.bc[5] == 4 else None)
            self.comm = engine.Comm(dist, device)
            self.G.set_comm(self.comm, lo, hi)
        if riemann:
            self.G.set_riemann_solver(riemann)       # the reference's --with-flux=hllc build
        if order != 2:
            self.G.set_reconstruction(order)         # --with-order=3
        self.time, self.dt, self.nstep = 0.0, 0.0, 0
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.rad_log = []
        self._reduce = self._make_reduce() if (self.world > 1 and self.comm is None) else None

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
        if self.nscal:
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
            return buf[:n].cpu().numpy().copy()
        return reduce

    def bvals_mhd(self):
        self.G.bvals_mhd(self.bc_local)          # with a communicator attached this includes the NCCL ghost exchange
        if self.world > 1 and self.comm is None:
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



def _setup_problem_hooks(G, problem, par, nx, xmin, dx, U0=None):
    """problem() data (when U0 is None), StaticGravPot and the Userwork_in_loop replay table for one Grid, as Run does."""
    if U0 is None:
        G.upload(problems.problem(problem, par, nx, xmin, dx))
    else:
        G.upload_active(U0)
    if problem == "ioniz_sphere":
        problems.ioniz_sphere_setup(par)
        G.set_static_grav_pot(problems.lib().ab2_planet_pot)                # StaticGravPot = PlanetPot (ioniz_sphere.c:166)
        cells, vals = problems.ioniz_sphere_reset_table(nx, xmin, dx)
        G.set_fixed_cells(cells, vals)                                       # Userwork_in_loop (ioniz_sphere.c:255-306)
        return G.apply_fixed_cells
    return None


class SmrRun:
    """main.c:355-669 with STATIC_MESH_REFINEMENT for a nest of single-Grid Domains on one GPU (root + one Domain per
    finer level, `job/num_domains` of them): RestrictCorrect / Prolongate / ionradRestrictCorrect / ionrad_prolong
    between the device mirrors of adjacent levels, one dt for the whole Mesh (new_dt.c:58-195), every fine level's
    radiation sub-cycle running to the root's coarse time."""

    def __init__(self, par, problem: str, U0: dict | None = None, *U0_fine: dict, device: int = 0):
        p = par
        self.nscal = 1 if problem in ("ifront", "ioniz_sphere") else 0
        self.ion = self.nscal == 1 and p.par_geti("problem", "nradplanes") > 0       # main.c:546
        nlev = (1 + len(U0_fine)) if U0 is not None else p.par_geti("job", "num_domains")     # U0 None: problem() data
        nx0 = tuple(p.par_geti("domain1", f"Nx{a}") for a in (1, 2, 3))
        xmin = [p.par_getd("domain1", f"x{a}min") for a in (1, 2, 3)]
        xmax = [p.par_getd("domain1", f"x{a}max") for a in (1, 2, 3)]
        dx0 = tuple((xmax[a] - xmin[a]) / nx0[a] for a in range(3))
        bc0 = [p.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3")]
        self.gamma = p.par_getd("problem", "gamma")
        self.cour_no = p.par_getd("time", "cour_no")
        self.tlim = p.par_getd("time", "tlim")
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.grids, self.bcs, self.ovs, disps, self.userwork = [], [], [], [], []
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
            self.userwork.append(_setup_problem_hooks(G, problem, p, nx, xm, dx, None if U0 is None else (U0 if n == 0 else U0_fine[n - 1])))
            self.grids.append(G); self.bcs.append(bc); disps.append(disp)
        if self.nscal:
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
        for w in self.userwork:                                 # Userwork_in_loop(&Mesh), main.c step 9e
            if w is not None:
                w()
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


# ---------------------------------------------------------------------------------------------------------------------
# SMR across GPUs: every level cut along x3 on planes common to all levels
# ---------------------------------------------------------------------------------------------------------------------
def make_group_reduce(dist, group=None):
    """The MPI_Allreduce sites of new_dt.c / ionrad_3d.c over `group` (see Run._make_reduce): op 0 MIN, 1 MAX, 2 SUM,
    4 = {SUM, MAX, MIN, MIN} in one exchange."""
    import torch
    ops = {0: dist.ReduceOp.MIN, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.SUM}
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = dist.get_world_size(group)
    buf = torch.zeros(4, dtype=torch.float64, device=dev)
    gat = torch.zeros(4 * n, dtype=torch.float64, device=dev)

    def reduce(vals, op):
        k = len(vals)
        buf[:k].copy_(torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float64)))
        if op == 4:
            dist.all_gather_into_tensor(gat, buf, group=group)
            g = gat.cpu().view(n, 4)
            return np.array([float(g[:, 0].sum()), float(g[:, 1].max()), float(g[:, 2].min()), float(g[:, 3].min())])
        dist.all_reduce(buf[:k], op=ops[op], group=group)
        return buf[:k].cpu().numpy().copy()        # a copy: on a CPU backend .cpu().numpy() aliases `buf`, which the next call overwrites
    return reduce


def exchange_x3_with(U, lo, hi, dist):
    """exchange_x3_ghosts with explicit neighbour ranks (None = no neighbour on that side).  Every rank posts the
    upward transfers of a level before the downward ones, so point-to-point calls match by order."""
    if lo is None and hi is None:
        return
    n3 = U.shape[1]
    ops = []
    for v in range(U.shape[0]):
        if hi is not None:
            ops.append(dist.P2POp(dist.isend, U[v, n3 - 2 * NG:n3 - NG], hi))
        if lo is not None:
            ops.append(dist.P2POp(dist.irecv, U[v, 0:NG], lo))
    for v in range(U.shape[0]):
        if lo is not None:
            ops.append(dist.P2POp(dist.isend, U[v, NG:2 * NG], lo))
        if hi is not None:
            ops.append(dist.P2POp(dist.irecv, U[v, n3 - NG:n3], hi))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def smr_slab_layout(par, nlev: int, world: int, cuts=None):
    """Geometry of SmrRunSlabs, no device involved: for every level the Domain (nx, disp, dx), the ranks that hold a
    slab of it, and per rank the slab (ka, kb in level cells), boundary flags, x3 neighbours and the overlap of the
    child slab on the parent slab.  Raises ValueError where the cut planes cannot serve the nest."""
    p = par
    nx0 = tuple(p.par_geti("domain1", f"Nx{a}") for a in (1, 2, 3))
    xmin = [p.par_getd("domain1", f"x{a}min") for a in (1, 2, 3)]
    xmax = [p.par_getd("domain1", f"x{a}max") for a in (1, 2, 3)]
    dx0 = tuple((xmax[a] - xmin[a]) / nx0[a] for a in range(3))
    bc0 = [p.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3")]
    if cuts is None:
        cuts = [r * nx0[2] // world for r in range(world + 1)]
    if len(cuts) != world + 1 or cuts[0] != 0 or cuts[-1] != nx0[2] or any(b <= a for a, b in zip(cuts, cuts[1:])):
        raise ValueError(f"cuts {cuts} do not partition the {nx0[2]} root layers over {world} ranks")
    periodic = world > 1 and (bc0[4] == 4 or bc0[5] == 4)
    levels = []
    for n in range(nlev):
        blk = f"domain{n + 1}"
        ref = 2 ** n
        nx = tuple(p.par_geti(blk, f"Nx{a}") for a in (1, 2, 3))
        disp = (0, 0, 0) if n == 0 else tuple(p.par_geti(blk, k) for k in ("iDisp", "jDisp", "kDisp"))
        dx = tuple(dx0[a] / 2.0 ** n for a in range(3))
        dom_k0, dom_k1 = disp[2], disp[2] + nx[2]
        if n > 0:
            for c in cuts[1:-1]:
                if c * ref in (dom_k0, dom_k1):
                    raise ValueError(f"level {n}: cut plane at root layer {c} coincides with the Domain's x3 boundary")
        slabs, holders = [], []
        for q in range(world):
            ka, kb = max(dom_k0, cuts[q] * ref), min(dom_k1, cuts[q + 1] * ref)
            if kb <= ka:
                slabs.append(None)
                continue
            if kb - ka < NG:
                raise ValueError(f"level {n}: rank {q} would hold a slab of {kb - ka} < {NG} layers; choose other cuts")
            holders.append(q)
            lnx, ldisp = (nx[0], nx[1], kb - ka), (disp[0], disp[1], ka)
            xm = tuple(xmin[a] if ldisp[a] == 0 else xmin[a] + float(ldisp[a]) * dx[a] for a in range(3))   # init_mesh.c:377-382
            bc = []
            for a in range(2):    # x1, x2: as on one GPU
                bc.append(bc0[2 * a] if disp[a] == 0 else 0)
                bc.append(bc0[2 * a + 1] if (disp[a] + nx[a]) // ref == nx0[a] else 0)
            at_lo, at_hi = (ka == 0), (kb == nx0[2] * ref)
            lo = q - 1 if ka > dom_k0 else None
            hi = q + 1 if kb < dom_k1 else None
            if periodic:
                if n > 0 and (at_lo or at_hi):
                    raise NotImplementedError("a fine Domain on a periodic x3 boundary across ranks")
                if n == 0:        # periodic root: the slabs form a ring (bvals_mhd.c:768-774, 814-820)
                    lo = lo if lo is not None else world - 1
                    hi = hi if hi is not None else 0
            bc.append(bc0[4] if (at_lo and lo is None) else 0)
            bc.append(bc0[5] if (at_hi and hi is None) else 0)
            ov = None
            if n > 0:
                pa = levels[n - 1]["slabs"][q]
                assert pa is not None and pa["ka"] * 2 <= ka and kb <= pa["kb"] * 2, "a child slab is nested in its parent slab"
                ddisp = disp
                ov = dict(cs=[ldisp[a] // 2 - pa["disp"][a] + NG for a in range(3)],
                          ce=[(ldisp[a] + lnx[a]) // 2 - pa["disp"][a] + NG - 1 for a in range(3)],
                          cdisp=list(ldisp), face=[0] * 6)
                for a in range(2):
                    ov["face"][2 * a] = 1 if ddisp[a] != 0 else 0
                    ov["face"][2 * a + 1] = 1 if (ddisp[a] + nx[a]) // ref != nx0[a] else 0
                # x3: a fine/coarse boundary only where the slab ends at the Domain's own end (not at a cut plane)
                ov["face"][4] = 1 if (ka == dom_k0 and dom_k0 != 0) else 0
                ov["face"][5] = 1 if (kb == dom_k1 and dom_k1 // ref != nx0[2]) else 0
            slabs.append(dict(ka=ka, kb=kb, nx=lnx, disp=ldisp, xmin=xm, bc=bc, lo=lo, hi=hi, ov=ov, dom=(dom_k0, dom_k1)))
        levels.append(dict(nx=nx, disp=disp, dx=dx, holders=holders, slabs=slabs))
    return dict(levels=levels, cuts=list(cuts), dx0=dx0)


def balanced_cuts(par, nlev: int, world: int):
    """Cut planes (root layers) for SmrRunSlabs that balance the zones per rank over all levels, subject to the
    layout's rules (slabs >= 4 layers on every level, no cut on a fine Domain's x3 boundary).  Dynamic programming
    over the root layers: minimise the most loaded rank.  Raises ValueError if no valid partition exists."""
    p = par
    n3 = p.par_geti("domain1", "Nx3")
    load = np.zeros(n3)                               # zones above each root layer, all levels
    doms = []
    for n in range(nlev):
        blk = f"domain{n + 1}"
        ref = 2 ** n
        nx = tuple(p.par_geti(blk, f"Nx{a}") for a in (1, 2, 3))
        k0 = 0 if n == 0 else p.par_geti(blk, "kDisp")
        doms.append((n, k0, k0 + nx[2]))
        for layer in range(n3):
            a, b = max(k0, layer * ref), min(k0 + nx[2], (layer + 1) * ref)
            if b > a:
                load[layer] += nx[0] * nx[1] * (b - a)
    pre = np.concatenate([[0.0], np.cumsum(load)])

    def ok(a, b):                                     # may one rank own root layers [a, b)?
        if b - a < NG:
            return False
        for n, k0, k1 in doms[1:]:
            ref = 2 ** n
            for c in (a, b):
                if 0 < c < n3 and c * ref in (k0, k1):
                    return False
            lo, hi = max(k0, a * ref), min(k1, b * ref)
            if hi > lo and hi - lo < NG:
                return False
        return True

    INF = float("inf")
    best = [[INF] * (n3 + 1) for _ in range(world + 1)]
    back = [[-1] * (n3 + 1) for _ in range(world + 1)]
    best[0][0] = 0.0
    for r in range(1, world + 1):
        for b in range(1, n3 + 1):
            for a in range(0, b):
                if best[r - 1][a] < INF and ok(a, b):
                    cost = max(best[r - 1][a], pre[b] - pre[a])
                    if cost < best[r][b]:
                        best[r][b], back[r][b] = cost, a
    if best[world][n3] == INF:
        raise ValueError(f"no valid x3 partition of {n3} root layers over {world} ranks for this nest")
    cuts, b = [n3], n3
    for r in range(world, 0, -1):
        b = back[r][b]
        cuts.append(b)
    return cuts[::-1]


class SmrRunSlabs:
    """SmrRun over several GPUs.  The reference cuts every Domain into NGrid_x3 equal Grids (init_mesh.c), which puts a
    child Grid and the parent cells under it on different ranks and needs its pack/send/unpack machinery
    (smr.c:1322-2352).  Here the cut planes are x3 planes of the ROOT Domain, the same physical planes on every level:
    rank r owns root layers [cuts[r], cuts[r+1]) and, of every finer Domain, the cells above exactly those layers.
    A rank's child slab is then nested in its own parent slab, so RestrictCorrect, Prolongate, ionradRestrictCorrect
    and ionrad_prolong stay the same-device pair kernels of the one-GPU path; what crosses NVLink per level is only the
    x3 ghost exchange of Run (rays run along x1, never cut) and the scalar reductions.  A rank holds no Grid on a level
    whose Domain does not reach its layers.  Requirements (checked): a cut plane may not coincide with a fine Domain's
    x3 boundary (the coarse cells flux-corrected across that face would live on the other rank), and every local slab
    is at least 4 cells thick.  Results equal the one-GPU SmrRun bit for bit (same kernels, same order; reductions are
    min / max / integer sums), except that a static potential is tabulated from slab-local MinX as in the reference."""

    def __init__(self, par, problem: str, U0_levels, dist, device: int = 0, cuts=None):
        p = par
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.nscal = 1 if problem in ("ifront", "ioniz_sphere") else 0
        self.ion = self.nscal == 1 and p.par_geti("problem", "nradplanes") > 0       # main.c:546
        nlev = len(U0_levels) if U0_levels is not None else p.par_geti("job", "num_domains")   # None: problem() data
        self.gamma = p.par_getd("problem", "gamma")
        self.cour_no = p.par_getd("time", "cour_no")
        self.tlim = p.par_getd("time", "tlim")
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        if cuts == "balanced":
            cuts = balanced_cuts(p, nlev, self.world)
        lay = smr_slab_layout(p, nlev, self.world, cuts)
        use_comm = dist.get_backend() == "nccl"      # as in Run: NCCL inside the library on GPUs, torch.distributed under gloo
        self.use_comm = use_comm
        self.cuts = lay["cuts"]
        dx0 = lay["dx0"]
        self.dom_nx = [L["nx"] for L in lay["levels"]]
        self.dom_dx = [L["dx"] for L in lay["levels"]]
        self.levels = []          # per level: None or dict(G, bc, lo, hi, ka, kb, disp, nx, dom, userwork, reduce)
        self.ovs = []             # ovs[n]: overlap of my level n+1 slab on my level n slab, or None
        for n, LV in enumerate(lay["levels"]):
            holders = LV["holders"]
            group = None if len(holders) == self.world else dist.new_group(holders)      # collective: every rank calls it
            S = LV["slabs"][self.rank]
            if n > 0:
                o = None
                if S is not None:
                    o = engine.Overlap()
                    for a in range(3):
                        o.cs[a], o.ce[a], o.cdisp[a] = S["ov"]["cs"][a], S["ov"]["ce"][a], S["ov"]["cdisp"][a]
                    for a in range(6):
                        o.face[a] = S["ov"]["face"][a]
                self.ovs.append(o)
            if S is None:
                self.levels.append(None)
                continue
            G = engine.Grid(S["nx"], LV["dx"], S["xmin"], self.gamma, self.nscal, device=device, torch_memory=not use_comm)
            if use_comm and len(holders) > 1:        # this level's communicator inside the library (csrc/comm.cu)
                G.set_comm(engine.Comm(dist, device, ranks=holders, group=group), S["lo"], S["hi"])
            k0 = S["ka"] - S["dom"][0]
            U0 = None if U0_levels is None else {v: np.ascontiguousarray(U0_levels[n][v][k0:k0 + S["nx"][2]]) for v in VARS[:5 + self.nscal]}
            userwork = _setup_problem_hooks(G, problem, p, S["nx"], S["xmin"], LV["dx"], U0)
            self.levels.append(dict(G=G, bc=S["bc"], lo=S["lo"], hi=S["hi"], ka=S["ka"], kb=S["kb"], disp=S["disp"], nx=S["nx"],
                                    dom=S["dom"], userwork=userwork,
                                    reduce=make_group_reduce(dist, group) if (len(holders) > 1 and not use_comm) else None))
        if self.nscal:
            ip = engine.ion_par_from(p, dx0)
            for L in self.levels:
                if L is not None:
                    L["G"].ion_init(ip)
        if use_comm:
            self.wcomm = engine.Comm(dist, device)
            self.all_reduce = lambda vals, op: self.wcomm.allreduce(vals, op)
        else:
            self.all_reduce = make_group_reduce(dist, None)
        self.time, self.dt, self.nstep = 0.0, 0.0, 0
        self.gdt = [0.0] * nlev
        self.rad_log = []
        # main.c:400-451
        self._restrict(0.0, init=True)
        self._bvals_all()
        if self.ion:
            for L in self.levels:
                if L is not None:
                    L["G"].bvals_ionrad(self.flux_i)
        self._prolongate()
        self.new_dt()

    # ---- pieces -------------------------------------------------------------------------------------
    def _bvals(self, n):
        L = self.levels[n]
        if L is not None:
            L["G"].bvals_mhd(L["bc"])
            if not self.use_comm:
                exchange_x3_with(L["G"].U_torch, L["lo"], L["hi"], self.dist)

    def _bvals_all(self):
        for n in range(len(self.levels)):
            self._bvals(n)

    def _restrict(self, dt_unused, init=False):
        for n in range(len(self.levels) - 2, -1, -1):
            if self.ovs[n] is not None:
                self.levels[n]["G"].smr_restrict_correct(self.levels[n + 1]["G"], self.ovs[n], 0.0 if init else self.gdt[n])

    def _prolongate(self):
        for n in range(len(self.levels) - 1):
            if self.ovs[n] is not None:
                self.levels[n]["G"].smr_prolongate(self.levels[n + 1]["G"], self.ovs[n])

    def new_dt(self):
        """new_dt.c:58-195 with the one-process running maximum of wave speeds over the levels (max_v is never reset,
        new_dt.c:53); each level's maximum is taken over all ranks first so the result equals the one-GPU run."""
        nlev = len(self.levels)
        mine = np.zeros(4 * ((3 * nlev + 3) // 4))
        for n, L in enumerate(self.levels):
            if L is not None:
                mine[3 * n:3 * n + 3] = L["G"].max_wave_speeds()
        allv = np.concatenate([self.all_reduce(mine[i:i + 4], 1) for i in range(0, len(mine), 4)])
        max_dti, max_v = 0.0, [0.0, 0.0, 0.0]
        for n in range(nlev):
            for a in range(3):
                if self.dom_nx[n][a] > 1:
                    v = float(allv[3 * n + a])
                    max_v[a] = max_v[a] if max_v[a] > v else v
                    x = max_v[a] / self.dom_dx[n][a]
                    max_dti = max_dti if max_dti > x else x
        dt = self.cour_no / max_dti if self.nstep == 0 else min(2.0 * self.dt, self.cour_no / max_dti)
        if self.time < self.tlim and (self.tlim - self.time) < dt:
            dt = self.tlim - self.time
        self.dt = dt
        self.gdt = [dt] * nlev
        return dt

    def step(self):
        """SmrRun.step with the x3 exchange behind every bvals_mhd and levels this rank does not hold skipped."""
        lv, nlev = self.levels, len(self.levels)
        if self.ion:
            tcoarse = 0.0
            for n in range(nlev):                               # main.c:546-556, coarse to fine
                L = lv[n]
                if L is not None:
                    if n == 0:
                        self.gdt[0], ni, nt, nc = L["G"].ion_radtransfer_3d(self.flux_i, self.time, self.cour_no, self.gdt[0], L["reduce"])
                        tcoarse = L["G"].ion_coarse_time
                    else:
                        lv[n - 1]["G"].smr_ion_prolong(L["G"], self.ovs[n - 1])
                        self.gdt[n], ni, nt, nc = L["G"].ion_radtransfer_3d_fine(tcoarse, self.gdt[n], L["reduce"])
                    rec = np.array([float(ni), float(nt), float(nc), self.gdt[n]])
                else:
                    rec = np.zeros(4)
                # every rank records every level's line (ionrad_3d.c:1041) and the Mesh dt it leaves behind (:1033)
                rec = self.all_reduce(rec, 1)
                self.gdt[n] = float(rec[3])
                self.rad_log.append((int(rec[0]), int(rec[1]), int(rec[2]), float(rec[3])))
                self.dt = self.gdt[n]
                self._bvals(n)
            for n in range(nlev - 2, -1, -1):                   # ionradRestrictCorrect, finest pair first
                if self.ovs[n] is not None:
                    lv[n]["G"].smr_ion_restrict(lv[n + 1]["G"], self.ovs[n])
            # ionradRestrictCorrect rewrites E and s of parent cells AFTER that level's bvals_mhd.  On one GPU the cells on
            # either side of a cut plane are plain active cells; here the neighbour's copies in my ghost layers must follow
            # (exchange only: physical-boundary ghosts stay as the reference leaves them, set before the restriction).
            for n in range(nlev - 1):
                if lv[n] is not None:
                    if self.use_comm:
                        lv[n]["G"].exchange_x3()
                    else:
                        exchange_x3_with(lv[n]["G"].U_torch, lv[n]["lo"], lv[n]["hi"], self.dist)
        for n in range(nlev):
            if lv[n] is not None:
                lv[n]["G"].integrate_3d_ctu(self.gdt[n])
        self._restrict(None)
        for L in lv:                                            # Userwork_in_loop(&Mesh)
            if L is not None and L["userwork"] is not None:
                L["userwork"]()
        self.nstep += 1
        self.time += self.dt
        self.new_dt()
        self._bvals_all()
        self._prolongate()

    def run(self, nlim):
        while self.time < self.tlim and self.nstep < nlim:
            self.step()

    def gather_level(self, n, what="U"):
        """The whole level-n Domain on every rank: dict var -> [Nx3][Nx2][Nx1] (what="U"), or the EdgeFlux array
        [Nx3+1][Nx2+1][Nx1+1] (what="edgeflux"; each rank contributes its ray rows, the top rank the last edge row)."""
        import torch
        L = self.levels[n]
        nx = self.dom_nx[n]
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        nv = 5 + self.nscal
        if what == "U":
            full = torch.zeros((nv, nx[2], nx[1], nx[0]), dtype=torch.float64, device=dev)
            if L is not None:
                a = L["G"].download_active()
                k0 = L["ka"] - L["dom"][0]
                full[:, k0:k0 + L["nx"][2]] = torch.from_numpy(np.stack([a[v] for v in VARS[:nv]])).to(dev)
            self.dist.all_reduce(full)                          # disjoint slabs: a sum of zeros and one copy is exact
            out = full.cpu().numpy()
            return {v: out[i] for i, v in enumerate(VARS[:nv])}
        full = torch.zeros((nx[2] + 1, nx[1] + 1, nx[0] + 1), dtype=torch.float64, device=dev)
        if L is not None:
            e = torch.from_numpy(L["G"].download_edgeflux()).to(dev)
            k0 = L["ka"] - L["dom"][0]
            top = (L["kb"] == L["dom"][1])
            full[k0:k0 + L["nx"][2] + (1 if top else 0)] = e[:L["nx"][2] + (1 if top else 0)]
        self.dist.all_reduce(full)
        return full.cpu().numpy()
