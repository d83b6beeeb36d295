"""ctypes binding of csrc/libathena_b200.so (include/athena_b200.h) -- the only way Python reaches the GPU here.

There is no fallback: if the shared library is missing or no CUDA device is visible, construction raises.
PyTorch is used for plumbing only (device memory that NCCL can address, streams): no torch op touches the state.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AB2_LIB") or os.path.join(_HERE, "csrc", "libathena_b200.so")   # AB2_LIB: tuning variants of the same library
NG = 4
VARS = ("d", "M1", "M2", "M3", "E", "s")
dp = C.POINTER(C.c_double)

GRAV_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_double)
REDUCE_FN = C.CFUNCTYPE(None, dp, C.c_int, C.c_int, C.c_void_p)


class EngineError(RuntimeError):
    """What the C shim would hand to ath_error() (utils.c:118)."""


class GridDesc(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int), ("nscal", C.c_int),
                ("dx1", C.c_double), ("dx2", C.c_double), ("dx3", C.c_double),
                ("x1min", C.c_double), ("x2min", C.c_double), ("x3min", C.c_double),
                ("gamma", C.c_double), ("device", C.c_int), ("stream", C.c_void_p),
                ("use_external_stream", C.c_int), ("external_U", C.c_void_p * 6)]


class IonPar(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("sigma_ph", "m_H", "mu", "e_gamma", "alpha_C", "k_B", "time_unit", "max_de_iter",
                 "max_de_therm_iter", "max_dx_iter", "max_de_step", "max_de_therm_step", "max_dx_step",
                 "tfloor", "tceil")] + [("maxiter", C.c_int), ("min_area", C.c_double), ("d_nlo", C.c_double)]


class Overlap(C.Structure):
    """ab2_overlap: GridOvrlpS as plain ints (include/athena_b200.h)."""
    _fields_ = [("cs", C.c_int * 3), ("ce", C.c_int * 3), ("face", C.c_int * 6), ("cdisp", C.c_int * 3)]


def make_overlap(child_nx, child_disp, root_nx, level=1, parent_disp=(0, 0, 0)) -> Overlap:
    """CGrid.ijks/ijke and the myFlx != NULL pattern init_grid.c:395-660 produces for a single-Grid parent Domain with
    displacement parent_disp (parent-level cells; 0 for the root) and a single-Grid child Domain at `level` with
    displacement child_disp (child-level cells from the root origin)."""
    o = Overlap()
    ref = 2 ** level
    for a in range(3):
        o.cs[a] = child_disp[a] // 2 - parent_disp[a] + NG
        o.ce[a] = (child_disp[a] + child_nx[a]) // 2 - parent_disp[a] + NG - 1
        o.cdisp[a] = child_disp[a]
        o.face[2 * a] = 1 if child_disp[a] != 0 else 0
        o.face[2 * a + 1] = 1 if (child_disp[a] + child_nx[a]) // ref != root_nx[a] else 0
    return o


EXPORTS = ["ab2_last_error", "ab2_version", "ab2_device_count", "ab2_grid_create", "ab2_grid_destroy",
           "ab2_bytes_allocated", "ab2_kernel_launches", "ab2_synchronize", "ab2_upload_cons_aos",
           "ab2_download_cons_aos", "ab2_upload_cons_soa", "ab2_download_cons_soa", "ab2_upload_edgeflux",
           "ab2_download_edgeflux", "ab2_device_ptr", "ab2_bvals_mhd", "ab2_max_wave_speeds", "ab2_new_dt",
           "ab2_set_static_grav_pot", "ab2_integrate_3d_ctu", "ab2_set_fixed_cells", "ab2_apply_fixed_cells",
           "ab2_ion_init", "ab2_ion_derived", "ab2_bvals_ionrad", "ab2_ion_entering_flux", "ab2_ion_radtransfer_3d",
           "ab2_profile_enable", "ab2_profile_kernels", "ab2_profile_kernel_name", "ab2_profile_read",
           "ab2_smr_restrict_correct", "ab2_smr_ion_restrict", "ab2_smr_prolongate", "ab2_smr_ion_prolong",
           "ab2_ion_radtransfer_3d_fine", "ab2_ion_coarse_time", "ab2_set_ion_memo", "ab2_ion_memo_stats",
           "ab2_set_riemann_solver", "ab2_set_reconstruction",
           "ab2_comm_get_unique_id", "ab2_comm_init_rank", "ab2_comm_destroy", "ab2_comm_rank", "ab2_comm_size",
           "ab2_comm_exchanges", "ab2_comm_reductions", "ab2_comm_allreduce", "ab2_comm_bcast", "ab2_grid_set_comm",
           "ab2_exchange_x3", "ab2_set_scratch_planes", "ab2_scratch_planes", "ab2_set_bc_table", "ab2_pin_host", "ab2_unpin_host",
           "ab2_output_vtk_field", "ab2_output_active_var"]
# libathena_b200_units.so (include/athena_b200_units.h): function-level test hooks, not part of the bound interface
UNITS_LIB_PATH = os.path.join(_HERE, "csrc", "libathena_b200_units.so")
UNIT_EXPORTS = ["ab2_unit_roe_flux", "ab2_unit_convert", "ab2_unit_lr_states", "ab2_unit_ieee", "ab2_unit_libm",
                "ab2_unit_hllc_flux", "ab2_unit_lr_states_hll", "ab2_unit_lr_states_ppm"]


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libathena_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["bash", os.path.join(_HERE, "csrc", "build.sh")], capture_output=True, text=True)
    if r.returncode != 0:
        raise EngineError("nvcc build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ab2_last_error.restype = C.c_char_p
    L.ab2_version.restype = C.c_char_p
    L.ab2_grid_create.argtypes = [C.POINTER(GridDesc)]
    L.ab2_grid_create.restype = vp
    L.ab2_grid_destroy.argtypes = [vp]
    L.ab2_grid_destroy.restype = None
    L.ab2_bytes_allocated.argtypes = [vp]
    L.ab2_bytes_allocated.restype = C.c_longlong
    L.ab2_kernel_launches.argtypes = [vp]
    L.ab2_kernel_launches.restype = C.c_longlong
    L.ab2_synchronize.argtypes = [vp]
    L.ab2_upload_cons_aos.argtypes = [vp, dp]
    L.ab2_download_cons_aos.argtypes = [vp, dp]
    L.ab2_upload_cons_soa.argtypes = [vp, C.POINTER(dp)]
    L.ab2_download_cons_soa.argtypes = [vp, C.POINTER(dp)]
    L.ab2_upload_edgeflux.argtypes = [vp, dp]
    L.ab2_download_edgeflux.argtypes = [vp, dp]
    L.ab2_device_ptr.argtypes = [vp, C.c_int]
    L.ab2_device_ptr.restype = vp
    L.ab2_bvals_mhd.argtypes = [vp, C.POINTER(C.c_int)]
    L.ab2_max_wave_speeds.argtypes = [vp, dp]
    L.ab2_new_dt.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, vp, vp, dp]
    L.ab2_set_static_grav_pot.argtypes = [vp, vp]
    L.ab2_integrate_3d_ctu.argtypes = [vp, C.c_double]
    L.ab2_set_fixed_cells.argtypes = [vp, C.c_long, C.POINTER(C.c_long), dp]
    L.ab2_apply_fixed_cells.argtypes = [vp]
    L.ab2_ion_init.argtypes = [vp, C.POINTER(IonPar)]
    L.ab2_ion_derived.argtypes = [C.POINTER(IonPar), C.c_double, C.c_double, C.c_double]
    L.ab2_ion_derived.restype = None
    L.ab2_bvals_ionrad.argtypes = [vp, C.c_double]
    L.ab2_ion_entering_flux.argtypes = [C.c_double, C.c_double]
    L.ab2_ion_entering_flux.restype = C.c_double
    L.ab2_ion_radtransfer_3d.argtypes = [vp, C.c_double, C.c_double, C.c_double, dp,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp]
    op = C.POINTER(Overlap)
    L.ab2_smr_restrict_correct.argtypes = [vp, vp, op, C.c_double]
    L.ab2_smr_ion_restrict.argtypes = [vp, vp, op]
    L.ab2_smr_prolongate.argtypes = [vp, vp, op]
    L.ab2_smr_ion_prolong.argtypes = [vp, vp, op]
    L.ab2_ion_radtransfer_3d_fine.argtypes = [vp, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp]
    L.ab2_ion_coarse_time.argtypes = [vp]
    L.ab2_ion_coarse_time.restype = C.c_double
    L.ab2_set_ion_memo.argtypes = [vp, C.c_int]
    L.ab2_ion_memo_stats.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.ab2_profile_enable.argtypes = [vp, C.c_int]
    L.ab2_profile_kernel_name.argtypes = [C.c_int]
    L.ab2_profile_kernel_name.restype = C.c_char_p
    L.ab2_profile_read.argtypes = [vp, dp, C.POINTER(C.c_longlong), C.c_int]
    L.ab2_set_riemann_solver.argtypes = [vp, C.c_int]
    L.ab2_set_reconstruction.argtypes = [vp, C.c_int]
    L.ab2_comm_get_unique_id.argtypes = [C.c_char_p]
    L.ab2_comm_init_rank.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.ab2_comm_init_rank.restype = vp
    L.ab2_comm_destroy.argtypes = [vp]
    L.ab2_comm_destroy.restype = None
    for f in (L.ab2_comm_rank, L.ab2_comm_size):
        f.argtypes = [vp]
    for f in (L.ab2_comm_exchanges, L.ab2_comm_reductions):
        f.argtypes = [vp]
        f.restype = C.c_longlong
    L.ab2_comm_allreduce.argtypes = [vp, dp, C.c_int, C.c_int]
    L.ab2_comm_bcast.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ab2_grid_set_comm.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ab2_exchange_x3.argtypes = [vp]
    L.ab2_set_scratch_planes.argtypes = [vp, C.c_int]
    L.ab2_set_bc_table.argtypes = [vp, C.c_int, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_long), dp]
    L.ab2_pin_host.argtypes = [vp, C.c_longlong]
    L.ab2_unpin_host.argtypes = [vp]
    L.ab2_output_vtk_field.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.ab2_output_active_var.argtypes = [vp, C.c_int, dp]
    L.ab2_scratch_planes.argtypes = [vp]
    _lib = L
    return L


_units = None


def units():
    """libathena_b200_units.so: single routines of the kernels on host arrays (tests only)."""
    global _units
    if _units is not None:
        return _units
    lib()                     # the units library links against the product library
    if not os.path.exists(UNITS_LIB_PATH):
        raise EngineError(f"{UNITS_LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(UNITS_LIB_PATH)
    L.ab2_unit_ieee.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
    L.ab2_unit_libm.argtypes = [C.c_int, dp, dp, dp, dp]
    L.ab2_unit_roe_flux.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp, dp]
    L.ab2_unit_convert.argtypes = [C.c_int, dp, C.c_int, C.c_double, C.c_int, dp, dp]
    L.ab2_unit_lr_states.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
    L.ab2_unit_lr_states_hll.argtypes = L.ab2_unit_lr_states.argtypes
    L.ab2_unit_hllc_flux.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, C.c_double, dp]
    L.ab2_unit_lr_states_ppm.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp]
    _units = L
    return L


class Comm:
    """ab2_comm: one NCCL communicator inside the library (csrc/comm.cu).  `dist` (torch.distributed, any backend) only
    carries the 128-byte unique id from the group's first rank to the others -- the job MPI_Bcast does in the C driver."""

    def __init__(self, dist, device, ranks=None, group=None):
        L = lib()
        world_rank = dist.get_rank()
        self.ranks = list(range(dist.get_world_size())) if ranks is None else list(ranks)
        assert world_rank in self.ranks
        self.rank, self.size = self.ranks.index(world_rank), len(self.ranks)
        box = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            if L.ab2_comm_get_unique_id(buf) != 0:
                raise EngineError(L.ab2_last_error().decode())
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=self.ranks[0], group=group)
        self.h = L.ab2_comm_init_rank(self.size, self.rank, box[0], device)
        if not self.h:
            raise EngineError(L.ab2_last_error().decode())

    def local(self, world_rank):
        """rank of a world rank inside this communicator; None stays None (no neighbour) and maps to -1"""
        return -1 if world_rank is None else self.ranks.index(world_rank)

    def allreduce(self, vals, op):
        """op 0 MIN, 1 MAX, 2 SUM over the communicator; vals: up to 64 doubles.  Returns a new array."""
        a = np.ascontiguousarray(vals, dtype=np.float64).copy()
        if lib().ab2_comm_allreduce(self.h, a.ctypes.data_as(dp), len(a), op) != 0:
            raise EngineError(lib().ab2_last_error().decode())
        return a

    @property
    def counters(self):
        return {"exchanges": lib().ab2_comm_exchanges(self.h), "reductions": lib().ab2_comm_reductions(self.h)}

    def close(self):
        if getattr(self, "h", None):
            lib().ab2_comm_destroy(self.h)
            self.h = None


def ion_par_from(par, root_dx) -> IonPar:
    """ion_radtransfer_init_3d's par_getd calls (ionrad_3d.c:742-757) + ionrad.c:115-131."""
    p = IonPar()
    for n, _ in IonPar._fields_[:15]:
        setattr(p, n, par.par_getd("ionradiation", n))
    p.maxiter = int(par.par_getd("ionradiation", "maxiter"))
    lib().ab2_ion_derived(C.byref(p), *root_dx)
    return p


class Grid:
    """One reference GridS mirrored on one GPU."""

    def __init__(self, nx, dx, xmin, gamma, nscal, device=0, torch_memory=False):
        L = lib()
        self.nx = tuple(int(n) for n in nx)
        self.dx = tuple(float(x) for x in dx)
        self.xmin = tuple(float(x) for x in xmin)
        self.nscal = int(nscal)
        self.nvar = 5 + self.nscal
        self.shape = (self.nx[2] + 2 * NG, self.nx[1] + 2 * NG, self.nx[0] + 2 * NG)
        self.device = device
        desc = GridDesc(self.nx[0], self.nx[1], self.nx[2], self.nscal, *self.dx, *self.xmin, float(gamma), device, None, 0)
        self.U_torch = None
        if torch_memory:
            import torch
            torch.cuda.set_device(device)
            self.U_torch = torch.zeros((self.nvar,) + self.shape, dtype=torch.float64, device=f"cuda:{device}")
            desc.stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            desc.use_external_stream = 1
            for v in range(self.nvar):
                desc.external_U[v] = self.U_torch[v].data_ptr()
            torch.cuda.synchronize(device)
        self.h = L.ab2_grid_create(C.byref(desc))
        if not self.h:
            raise EngineError(L.ab2_last_error().decode())
        self._keep = []          # callbacks that must outlive the handle

    # -- helpers ---------------------------------------------------------------------------------
    def _ck(self, rc):
        err = getattr(self, "_cb_error", None)
        if err is not None:
            self._cb_error = None
            raise EngineError(f"reduction callback failed: {err!r}") from err
        if rc != 0:
            raise EngineError(lib().ab2_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().ab2_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def bytes_allocated(self):
        return lib().ab2_bytes_allocated(self.h)

    @property
    def kernel_launches(self):
        return lib().ab2_kernel_launches(self.h)

    def synchronize(self):
        self._ck(lib().ab2_synchronize(self.h))

    def set_bc_table(self, face, dst, src, coef):
        """A problem file's boundary function in table form (ab2_set_bc_table): dst [n], src [nvar][n] (-1: constant), coef [nvar][n]"""
        dst = np.ascontiguousarray(dst, dtype=np.int64)
        src = np.ascontiguousarray(src, dtype=np.int64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        assert src.shape == (self.nvar, len(dst)) and coef.shape == src.shape
        lp = C.POINTER(C.c_long)
        self._ck(lib().ab2_set_bc_table(self.h, int(face), len(dst), dst.ctypes.data_as(lp), src.ctypes.data_as(lp), coef.ctypes.data_as(dp)))

    def output_vtk_field(self, field, prim=False):
        """dump_vtk's bytes for one CELL_DATA field of the active zones: float32 big-endian, as a uint8 array"""
        n = self.nx[0] * self.nx[1] * self.nx[2] * (3 if field == 1 else 1)
        out = np.empty(n, dtype=np.float32)
        self._ck(lib().ab2_output_vtk_field(self.h, int(field), 1 if prim else 0, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out.view(np.uint8)

    def output_active_var(self, var):
        out = np.empty((self.nx[2], self.nx[1], self.nx[0]))
        self._ck(lib().ab2_output_active_var(self.h, int(var), out.ctypes.data_as(dp)))
        return out

    def set_scratch_planes(self, planes):
        """x3-chunked CTU scratch: `planes` active planes per chunk, 0 = whole Grid (results do not depend on it)"""
        self._ck(lib().ab2_set_scratch_planes(self.h, int(planes)))

    @property
    def scratch_planes(self):
        return lib().ab2_scratch_planes(self.h)

    def profile_enable(self, on=True):
        self._ck(lib().ab2_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, reset=True) -> dict:
        """{kernel name: (total ms, launches)} measured with CUDA events on the launching stream."""
        n = lib().ab2_profile_kernels()
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self._ck(lib().ab2_profile_read(self.h, ms, cnt, 1 if reset else 0))
        return {lib().ab2_profile_kernel_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}

    # -- coherence -------------------------------------------------------------------------------
    def upload(self, U: dict):
        """U: dict var -> float64 array [n3][n2][n1] incl. ghosts (SoA)."""
        arrs = [np.ascontiguousarray(U[v], dtype=np.float64) for v in VARS[:self.nvar]]
        for a in arrs:
            assert a.shape == self.shape, (a.shape, self.shape)
        ptrs = (dp * 6)(*[a.ctypes.data_as(dp) for a in arrs])
        self._ck(lib().ab2_upload_cons_soa(self.h, ptrs))

    def download(self) -> dict:
        out = {v: np.empty(self.shape) for v in VARS[:self.nvar]}
        ptrs = (dp * 6)(*[out[v].ctypes.data_as(dp) for v in VARS[:self.nvar]])
        self._ck(lib().ab2_download_cons_soa(self.h, ptrs))
        return out

    def upload_aos(self, U_aos: np.ndarray):
        a = np.ascontiguousarray(U_aos, dtype=np.float64)
        assert a.shape == self.shape + (self.nvar,)
        self._ck(lib().ab2_upload_cons_aos(self.h, a.ctypes.data_as(dp)))

    def download_aos(self) -> np.ndarray:
        a = np.empty(self.shape + (self.nvar,))
        self._ck(lib().ab2_download_cons_aos(self.h, a.ctypes.data_as(dp)))
        return a

    def upload_active(self, U: dict):
        full = {v: np.zeros(self.shape) for v in VARS[:self.nvar]}
        for v in full:
            full[v][NG:-NG, NG:-NG, NG:-NG] = U[v]
        self.upload(full)

    def download_active(self) -> dict:
        return {v: a[NG:-NG, NG:-NG, NG:-NG].copy() for v, a in self.download().items()}

    def download_edgeflux(self) -> np.ndarray:
        ef = np.empty((self.nx[2] + 1, self.nx[1] + 1, self.nx[0] + 1))
        self._ck(lib().ab2_download_edgeflux(self.h, ef.ctypes.data_as(dp)))
        return ef

    def upload_edgeflux(self, ef):
        a = np.ascontiguousarray(ef, dtype=np.float64)
        self._ck(lib().ab2_upload_edgeflux(self.h, a.ctypes.data_as(dp)))

    # -- hot path --------------------------------------------------------------------------------
    def bvals_mhd(self, bc):
        self._ck(lib().ab2_bvals_mhd(self.h, (C.c_int * 6)(*bc)))

    def max_wave_speeds(self):
        out = (C.c_double * 3)()
        self._ck(lib().ab2_max_wave_speeds(self.h, out))
        return tuple(out)

    def new_dt(self, cour_no, dt_old, nstep, time, tlim, reduce=None):
        dt = C.c_double(0.0)
        cb = self._reduce_cb(reduce)
        self._ck(lib().ab2_new_dt(self.h, cour_no, dt_old, nstep, time, tlim, cb, None, C.byref(dt)))
        return dt.value

    def set_static_grav_pot(self, fn):
        """fn: a C function pointer double(double,double,double) (ctypes function or CFUNCTYPE instance), or None."""
        if fn is None:
            self._ck(lib().ab2_set_static_grav_pot(self.h, None))
            return
        self._keep.append(fn)
        self._ck(lib().ab2_set_static_grav_pot(self.h, C.cast(fn, C.c_void_p)))

    def set_riemann_solver(self, kind):
        """0 = Roe + H-correction (rsolvers/roe.c, default), 1 = HLLC (rsolvers/hllc.c): the reference's --with-flux"""
        self._ck(lib().ab2_set_riemann_solver(self.h, int(kind)))

    def set_reconstruction(self, order):
        """2 = PLM (lr_states_plm.c, default), 3 = PPM (lr_states_ppm.c): the reference's --with-order"""
        self._ck(lib().ab2_set_reconstruction(self.h, int(order)))

    def integrate_3d_ctu(self, dt):
        self._ck(lib().ab2_integrate_3d_ctu(self.h, dt))

    def set_fixed_cells(self, cells, vals):
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert vals.shape == (6, len(cells))
        self._ck(lib().ab2_set_fixed_cells(self.h, len(cells), cells.ctypes.data_as(C.POINTER(C.c_long)), vals.ctypes.data_as(dp)))

    def apply_fixed_cells(self):
        self._ck(lib().ab2_apply_fixed_cells(self.h))

    def ion_init(self, ionpar: IonPar):
        self._ck(lib().ab2_ion_init(self.h, C.byref(ionpar)))

    def bvals_ionrad(self, flux_i):
        self._ck(lib().ab2_bvals_ionrad(self.h, flux_i))

    def ion_radtransfer_3d(self, flux_i, time, cour_no, dt, reduce=None):
        """Returns (new dt, niter, ntherm, nchem) -- the numbers of the reference's stderr line (ionrad_3d.c:1041)."""
        dtc = C.c_double(dt)
        ni, nc, nt = C.c_int(0), C.c_int(0), C.c_int(0)
        cb = self._reduce_cb(reduce)
        self._ck(lib().ab2_ion_radtransfer_3d(self.h, flux_i, time, cour_no, C.byref(dtc), C.byref(ni), C.byref(nc), C.byref(nt), cb, None))
        return dtc.value, ni.value, nt.value, nc.value

    def ion_radtransfer_3d_fine(self, tcoarse, dt, reduce=None):
        dtc = C.c_double(dt)
        ni, nc, nt = C.c_int(0), C.c_int(0), C.c_int(0)
        cb = self._reduce_cb(reduce)
        self._ck(lib().ab2_ion_radtransfer_3d_fine(self.h, tcoarse, C.byref(dtc), C.byref(ni), C.byref(nc), C.byref(nt), cb, None))
        return dtc.value, ni.value, nt.value, nc.value

    @property
    def ion_coarse_time(self):
        return lib().ab2_ion_coarse_time(self.h)

    def set_ion_memo(self, on):
        """Sparse sub-cycle on/off (include/athena_b200.h ab2_set_ion_memo): off re-evaluates every chunk in every iteration."""
        self._ck(lib().ab2_set_ion_memo(self.h, int(bool(on))))

    def ion_memo_stats(self):
        """(chunks re-used, rays with a chunk left to evaluate, chunks per pass) as the latest scan of the records found them."""
        out = (C.c_longlong * 3)()
        self._ck(lib().ab2_ion_memo_stats(self.h, out))
        return int(out[0]), int(out[1]), int(out[2])

    # -- more than one GPU ------------------------------------------------------------------------
    def set_comm(self, comm, lo=None, hi=None):
        """Attach to an engine.Comm with the x3 neighbours' WORLD ranks (None = physical or fine/coarse boundary):
        bvals_mhd then exchanges ghost slabs and the dt / sub-cycle reductions span the communicator, all inside the library."""
        self.comm = comm
        self._ck(lib().ab2_grid_set_comm(self.h, comm.h if comm else None, comm.local(lo) if comm else -1, comm.local(hi) if comm else -1))

    def exchange_x3(self):
        self._ck(lib().ab2_exchange_x3(self.h))

    # -- SMR: this Grid as the parent of `child` ---------------------------------------------------
    def smr_restrict_correct(self, child, o, dt_parent):
        self._ck(lib().ab2_smr_restrict_correct(self.h, child.h, C.byref(o), dt_parent))

    def smr_ion_restrict(self, child, o):
        self._ck(lib().ab2_smr_ion_restrict(self.h, child.h, C.byref(o)))

    def smr_prolongate(self, child, o):
        self._ck(lib().ab2_smr_prolongate(self.h, child.h, C.byref(o)))

    def smr_ion_prolong(self, child, o):
        self._ck(lib().ab2_smr_ion_prolong(self.h, child.h, C.byref(o)))

    def _reduce_cb(self, reduce):
        if reduce is None:
            return None
        if getattr(self, "_reduce_src", None) is reduce:
            return self._reduce_ptr

        def _cb(vals, n, op, ctx):
            arr = np.ctypeslib.as_array(vals, shape=(n,))
            try:
                arr[:] = reduce(arr.copy(), op)
            except BaseException as e:          # ctypes would print and swallow it: poison the values so the C loop stops,
                self._cb_error = e              # and re-raise from _ck() once the C call has returned
                arr[:] = np.nan
        cb = REDUCE_FN(_cb)
        self._keep.append(cb)
        self._reduce_src, self._reduce_ptr = reduce, C.cast(cb, C.c_void_p)
        return self._reduce_ptr
