/* athena_b200.h -- C ABI of libathena_b200.so: the B200-native engine for Atmospheric Athena's
 * per-timestep hot path (3D CTU hydro + ion-plane radiative transfer), FP64 CUDA for sm_100a.
 *
 * The reference has no plugin ABI: its "operator API" is link-time symbol substitution of
 *   new_dt(), bvals_mhd(), integrate_init()->integrate_3d_ctu(), ion_radtransfer_init()->ion_radtransfer_3d(),
 *   bvals_ionrad(), RestrictCorrect()/Prolongate()   (src/prototypes.h:75-77,147,224-228)
 * all taking MeshS* / DomainS*.  Those structs belong to the reference's headers, so this library takes
 * plain pointers and sizes instead; `atmospheric-athena_b200/shim/athena_shim.c` is the ~200-line C file a
 * maintainer compiles against the reference's own athena.h to bind the reference symbols to these entry
 * points (INTEGRATION.md shows the link line).  Every entry point below names the reference function it
 * replaces (paths relative to /root/reference/src).
 *
 * Conventions
 *  - One ab2_grid per reference GridS, owned by the caller through an opaque handle; one CUDA device and
 *    one stream per handle.  Not thread-safe per handle (neither is the reference, SURVEY 8b "Threading"); different
 *    handles may be driven from different host threads at the same time (bench.py's e2e leg does: error text is per thread).
 *  - Arrays with ghosts have extents [Nx3+8][Nx2+8][Nx1+8], x1 fastest, exactly GridS.U (athena.h:289-348,
 *    ath_array.c:78-120).  "aos" = the reference's ConsS layout, nvar = 5+nscal doubles per cell
 *    (d,M1,M2,M3,E[,s0]); "soa" = one array per variable (the device layout).
 *  - Every function returns 0 on success and a negative code on failure; ab2_last_error() then holds the
 *    message the shim passes to ath_error() (utils.c:118).  There is NO CPU fallback: without a CUDA device
 *    ab2_grid_create fails.
 */
#ifndef ATHENA_B200_H
#define ATHENA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define AB2_NGHOST 4

typedef struct ab2_grid ab2_grid;

typedef struct ab2_grid_desc {
  int nx1, nx2, nx3;              /* GridS.Nx[0..2]: active zones on this Grid                              */
  int nscal;                      /* NSCALARS: 0 (hydro only) or 1 (ion builds, s[0] = neutral density)      */
  double dx1, dx2, dx3;           /* GridS.dx1..3                                                            */
  double x1min, x2min, x3min;     /* GridS.MinX[0..2] (cc_pos.c)                                             */
  double gamma;                   /* globals.h: Gamma                                                        */
  int device;                     /* CUDA device ordinal                                                     */
  void* stream;                   /* cudaStream_t to run on when use_external_stream != 0 (NULL is then the    */
                                  /* legacy default stream, which is what torch hands out); otherwise ignored   */
  int use_external_stream;        /* 0: the library creates its own non-blocking stream                        */
  double* external_U[6];          /* optional caller-owned DEVICE arrays for d,M1,M2,M3,E,s (all or none):    */
                                  /* lets a host framework (torch / NCCL) exchange ghost slabs in place        */
} ab2_grid_desc;

/* <ionradiation> block as read by ion_radtransfer_init_3d (ionradiation/ionrad_3d.c:742-757), plus the two
 * values ion_radtransfer_init derives from the ROOT mesh spacing (ionradiation/ionrad.c:115-131). */
typedef struct ab2_ion_par {
  double sigma_ph, m_H, mu, e_gamma, alpha_C, k_B, time_unit;
  double max_de_iter, max_de_therm_iter, max_dx_iter;
  double max_de_step, max_de_therm_step, max_dx_step;
  double tfloor, tceil;
  int maxiter;
  double min_area, d_nlo;
} ab2_ion_par;

/* StaticGravPot: athena.h:503 GravPotFun_t, globals.h:24 */
typedef double (*ab2_grav_pot_fn)(double x1, double x2, double x3);
/* Combines scalars over the Grids of one Domain; op: 0 MIN, 1 MAX, 2 SUM, 4 = the four sub-cycle scalars of one ion
 * iteration in ONE exchange (n == 4: vals[0] SUM, vals[1] MAX, vals[2..3] MIN).  Stands where the reference calls
 * MPI_Allreduce (new_dt.c:175-179; ionrad_3d.c:275,399,554,672). */
typedef void (*ab2_reduce_fn)(double* vals, int n, int op, void* ctx);

/* ---- lifetime ------------------------------------------------------------------------------------ */
const char* ab2_last_error(void);
const char* ab2_version(void);
int  ab2_device_count(void);
/* integrate_init / integrate_init_3d (integrators/integrate.c:24-88, integrate_3d_ctu.c:3374-3492): allocates
 * the device mirror of U and all CTU scratch. */
ab2_grid* ab2_grid_create(const ab2_grid_desc* desc);
/* integrate_destruct (integrate_3d_ctu.c:3498-3550) */
void ab2_grid_destroy(ab2_grid* g);
/* The reference sizes its CTU scratch to the whole Grid (integrate_3d_ctu.c:3374-3472: ~0.6 kB per zone with a passive scalar).
 * Here it may hold only `planes` active x3 planes (+4 either side): ab2_integrate_3d_ctu then sweeps the Grid chunk by
 * chunk with identical results.  ab2_grid_create picks 0 (whole Grid) when that fits the device next to the Grid's other
 * arrays and a chunk size otherwise (AB2_CHUNK_PLANES overrides); this call changes it.  SMR Grids must stay at 0. */
int ab2_set_scratch_planes(ab2_grid* g, int planes);
int ab2_scratch_planes(const ab2_grid* g);
long long ab2_bytes_allocated(const ab2_grid* g);
long long ab2_kernel_launches(const ab2_grid* g);
int  ab2_synchronize(ab2_grid* g);

/* ---- coherence points: host U <-> device mirror ---------------------------------------- */
int ab2_upload_cons_aos(ab2_grid* g, const double* U_aos);      /* after problem() / restart_grids()          */
int ab2_download_cons_aos(ab2_grid* g, double* U_aos);          /* before a dump fires / Userwork_after_loop  */
int ab2_upload_cons_soa(ab2_grid* g, const double* const vars[6]);
int ab2_download_cons_soa(ab2_grid* g, double* const vars[6]);
int ab2_upload_edgeflux(ab2_grid* g, const double* edgeflux);   /* GridS.EdgeFlux [Nx3+1][Nx2+1][Nx1+1]        */
int ab2_download_edgeflux(ab2_grid* g, double* edgeflux);
double* ab2_device_ptr(ab2_grid* g, int var);                   /* device array of variable 0..5               */

/* ---- hot path -------------------------------------------------------------------------------------- */
/* bvals_mhd (bvals_mhd.c:174-508).  bc[6] = bc_ix1,bc_ox1,bc_ix2,bc_ox2,bc_ix3,bc_ox3 with the reference's flag
 * values (1 reflect, 2 outflow, 4 periodic, 5 conductor); a value <= 0 marks a face owned by a neighbouring Grid
 * (MPI in the reference, NCCL slab exchange here) or prolongated later, and is left untouched. */
int ab2_bvals_mhd(ab2_grid* g, const int bc[6]);
/* new_dt (new_dt.c:69-139): out[3] = max over active zones of |v_d| + sqrt(gamma P/d). */
int ab2_max_wave_speeds(ab2_grid* g, double out[3]);
/* new_dt (new_dt.c:32-198) for a single-Grid mesh: CFL, 2x growth limit, tlim clip.  reduce may be NULL. */
int ab2_new_dt(ab2_grid* g, double cour_no, double dt_old, int nstep, double time, double tlim,
               ab2_reduce_fn reduce, void* ctx, double* dt_out);
/* StaticGravPot enrolment: tabulates the user's potential on the host at exactly the arguments the reference
 * passes (integrate_3d_ctu.c:334-336,1172-1215,1468-1522,1878-1933,2746-2777) and uploads the tables. NULL clears;
 * a second call replaces the tables.  The callback runs on ONE host thread (problem-file potentials are written for a
 * serial code); AB2_GRAV_THREADS=n in the environment tabulates with n threads if the function is re-entrant. */
int ab2_set_static_grav_pot(ab2_grid* g, ab2_grav_pot_fn phi);
/* integrate_3d_ctu (integrators/integrate_3d_ctu.c:110-3368) with GridS.dt = dt. */
int ab2_integrate_3d_ctu(ab2_grid* g, double dt);
/* The reference fixes its Riemann solver and the matching form of the PLM characteristic tracing at configure time
 * (--with-flux, src/defs.h.in FLUX_DEF); the replacement library serves either build: kind 0 = Roe with H-correction
 * and HLLE fallback (rsolvers/roe.c, the shipped configuration, default), kind 1 = HLLC (rsolvers/hllc.c:50-261 with
 * lr_states_plm.c:257-293; no H-correction -- etah is defined in roe.c only), kind 2 = Roe for a build configured
 * WITHOUT --enable-h-correction (integrate_3d_ctu.c:2299-2426 compiled out, etah stays 0: the configuration of the
 * reference's linear-wave convergence deck).  Returns -1 for an unknown kind. */
int ab2_set_riemann_solver(ab2_grid* g, int kind);
/* The reference's --with-order (defs.h ACCURACY): 2 = SECOND_ORDER_CHAR, reconstruction/lr_states_plm.c (default);
 * 3 = THIRD_ORDER_CHAR, reconstruction/lr_states_ppm.c:67-651 (PPM in characteristic variables). */
int ab2_set_reconstruction(ab2_grid* g, int order);
/* Userwork_in_loop replay: a static list of cells (flat indices incl. ghosts) reset to fixed values
 * vals[6][n] every step, e.g. prob/ioniz_sphere.c:255-306. */
int ab2_set_fixed_cells(ab2_grid* g, long n, const long* cells, const double* vals);
int ab2_apply_fixed_cells(ab2_grid* g);
/* A problem file's boundary function (bvals_mhd_fun, bvals_mhd.c:917-947) in table form, so it need not run on the host:
 * for face = 0..5 (ix1, ox1, ix2, ox2, ix3, ox3) entry t sets, for every variable v,
 *   U[v][dst[t]] = src[v*n+t] >= 0 ? coef[v*n+t] * U[v][src[v*n+t]] : coef[v*n+t]
 * (flat cell indices incl. ghosts) -- constants (inflow) and signed copies (reflection, outflow, periodic wrap).  A face
 * whose table is set is applied by ab2_bvals_mhd when bc[face] == -2, in the reference's order (x1, x2, x3).  n = 0 clears. */
int ab2_set_bc_table(ab2_grid* g, int face, long n, const long* dst, const long* src, const double* coef);

/* ion_radtransfer_init (ionradiation/ionrad.c:64-149, ionrad_3d.c:739-814) */
int ab2_ion_init(ab2_grid* g, const ab2_ion_par* par);
/* ionrad.c:115-131: fills min_area and d_nlo from the root mesh spacing */
void ab2_ion_derived(ab2_ion_par* par, double dx1, double dx2, double dx3);
/* bvals_ionrad (ionradiation/bvals_ionrad.c:63-146, outflow_flux_ix1 :308-332) */
int ab2_bvals_ionrad(ab2_grid* g, double flux_i);
/* entering flux at the root upstream face at mesh time t (ionradplane_3d.c:265) */
double ab2_ion_entering_flux(double flux_i, double time);
/* ion_radtransfer_3d (ionradiation/ionrad_3d.c:862-1047) for a root-level Grid and one plane with dir = -1.
 * *dt is GridS.dt on entry and on exit (the routine may shorten it).  niter/nchem/ntherm are the numbers of the
 * reference's stderr line (:1041). */
int ab2_ion_radtransfer_3d(ab2_grid* g, double flux_i, double time, double cour_no, double* dt,
                           int* niter, int* nchem, int* ntherm, ab2_reduce_fn reduce, void* ctx);

/* ---- static mesh refinement: one parent Grid and one child Grid on the same device ------------------------ */
/* GridOvrlpS (athena.h:258-275) as plain ints.  cs/ce = CGrid.ijks/ijke on the PARENT (local indices incl. ghosts);
 * the child's own overlap is its whole active region (one Grid per Domain).  face[dim] = 1 where the reference
 * allocates myFlx[dim] (fine/coarse boundary interior to the root Domain, init_grid.c:508-510,624-626).
 * cdisp = DomainS.Disp of the child in child-level cells.  The tables come from init_grid.c unchanged. */
typedef struct ab2_overlap {
  int cs[3], ce[3];
  int face[6];
  int cdisp[3];
} ab2_overlap;
/* RestrictCorrect (smr.c:1207-2352): child -> parent 8-cell average, then U -= +-dt/dx (F_coarse - <F_fine>) on the
 * coarse cells outside the child faces, with the final fluxes of both Grids' last ab2_integrate_3d_ctu call. */
int ab2_smr_restrict_correct(ab2_grid* parent, ab2_grid* child, const ab2_overlap* o, double dt_parent);
/* ionradRestrictCorrect (smr.c:85-1205): E and s[0] only */
int ab2_smr_ion_restrict(ab2_grid* parent, ab2_grid* child, const ab2_overlap* o);
/* Prolongate (smr.c:2359-2923) + ProCon (:3068-3325): parent -> child ghost zones, MC-limited linear interpolation */
int ab2_smr_prolongate(ab2_grid* parent, ab2_grid* child, const ab2_overlap* o);
/* ionrad_prolong_snd + ionrad_prolong_rcv (ionradiation/ionrad_smr.c:345-459, 34-343): parent EdgeFlux plane at the
 * child's upstream x1 face -> child EdgeFlux[k][j][0] */
int ab2_smr_ion_prolong(ab2_grid* parent, ab2_grid* child, const ab2_overlap* o);
/* ion_radtransfer_3d on a level > 0 (ionrad_3d.c:919-1014, finegrid): sub-cycles to tcoarse; *dt returns GridS.dt */
int ab2_ion_radtransfer_3d_fine(ab2_grid* g, double tcoarse, double* dt, int* niter, int* nchem, int* ntherm,
                                ab2_reduce_fn reduce, void* ctx);
/* the reference's file-static tcoarse (ionrad_3d.c:47,1028): dt_done of the last root-level ab2_ion_radtransfer_3d */
double ab2_ion_coarse_time(const ab2_grid* g);
/* Sparse sub-cycle.  ion_radtransfer_3d re-evaluates every cell in every sub-cycle iteration (ionrad_3d.c:919-1014); the
 * library keeps one record per ray and 32-cell chunk and re-uses a chunk's rates, dt limits, check_range count and wave-speed
 * maxima while every input of that arithmetic is bitwise unchanged (cells at the neutral floor that ionization_update's own
 * condition :575 leaves alone, same incoming flux, sign states at their fixed point) -- results are bit-identical to the full
 * re-evaluation.  on = 0 turns the memo off (every chunk evaluated every iteration; also AB2_ION_SKIP=0 in the environment at
 * ab2_ion_init).  ab2_ion_memo_stats, from the latest scan of the records (queued behind every update of a call but the
 * last): out[0] = chunks re-used, out[1] = rays that still had a chunk to evaluate, out[2] = chunks per pass of this Grid. */
int ab2_set_ion_memo(ab2_grid* g, int on);
int ab2_ion_memo_stats(ab2_grid* g, long long out[3]);

/* ---- the device side of the driver's writers ---------------------------------------------------------- */
/* Page-lock a host array the driver owns (GridS.U) so the coherence copies run at PCIe speed straight into it. */
int ab2_pin_host(void* p, long long bytes);
int ab2_unpin_host(void* p);
/* One CELL_DATA field of dump_vtk (dump_vtk.c:164-312) for the active zones, the very bytes the reference writes: float32,
 * big-endian, k-j-i order.  field 0 density, 1 momentum / velocity (3 interleaved components), 2 total energy / pressure,
 * 3 scalar / specific scalar; prim != 0: primitive variables (Cons_to_Prim, convert_var.c:97-160).  dst: host floats. */
int ab2_output_vtk_field(ab2_grid* g, int field, int prim, float* dst);
/* Active zones of one conserved variable as FP64 in k-j-i order: a DENSITY / n-MOMENTUM / ENERGY / SCALAR block of
 * dump_restart (restart.c:548-760) without the ghost zones and the AoS transposition of the full mirror. */
int ab2_output_active_var(ab2_grid* g, int var, double* dst);

/* ---- per-kernel device timing (bench.py's roofline) ------------------------------------------------------ */
/* When enabled every kernel launch is bracketed by CUDA events on the launching stream; ab2_profile_read drains
 * them into ms[id] / counts[id] (id < ab2_profile_kernels(), names from ab2_profile_kernel_name). */
int ab2_profile_enable(ab2_grid* g, int on);
int ab2_profile_kernels(void);
const char* ab2_profile_kernel_name(int id);
int ab2_profile_read(ab2_grid* g, double* ms, long long* counts, int reset);

/* ---- more than one GPU: NCCL inside the library ------------------------------------------------------------
 * One process per GPU, one ab2_comm per set of Grids that exchange ghost zones and reduce together (the Grids of a
 * Domain cut into x3 slabs: NGrid_x1 = NGrid_x2 = 1, NGrid_x3 = ranks, rays run along x1 and are never cut).  The
 * driver creates the 128-byte id on one rank, distributes it the way it distributes everything else (MPI_Bcast in
 * main.c's MPI build; torch.distributed in the Python host) and every rank calls ab2_comm_init_rank.  libnccl is
 * dlopen'ed at the first of these calls. */
typedef struct ab2_comm ab2_comm;
#define AB2_UNIQUE_ID_BYTES 128
int ab2_comm_get_unique_id(char id[AB2_UNIQUE_ID_BYTES]);                       /* ncclGetUniqueId                      */
ab2_comm* ab2_comm_init_rank(int nranks, int rank, const char id[AB2_UNIQUE_ID_BYTES], int device);   /* MPI_Comm_create */
void ab2_comm_destroy(ab2_comm* c);
int ab2_comm_rank(const ab2_comm* c);
int ab2_comm_size(const ab2_comm* c);
long long ab2_comm_exchanges(const ab2_comm* c);
long long ab2_comm_reductions(const ab2_comm* c);
/* MPI_Allreduce of n <= 64 host doubles, op 0 MIN / 1 MAX / 2 SUM (new_dt.c:175-179 for meshes whose dt the driver
 * combines itself; main.c:725-734 zone-cycle totals) and MPI_Bcast of <= 512 raw bytes */
int ab2_comm_allreduce(ab2_comm* c, double* vals, int n, int op);
int ab2_comm_bcast(ab2_comm* c, void* data, int nbytes, int root);
/* Attach a Grid to its communicator with its x3 neighbours (GridS.lx3_id / rx3_id, init_grid.c:253-283; -1 where the
 * face is a physical boundary or a fine/coarse boundary).  From then on ab2_bvals_mhd also exchanges the 4 ghost layers
 * at those faces (bvals_mhd.c:423-446: ncclSend/ncclRecv of the contiguous SoA slabs, in place, one NCCL group),
 * ab2_max_wave_speeds / ab2_new_dt return the maximum over all ranks (new_dt.c:175-179) and the ion sub-cycle combines
 * its scalars over all ranks with one all-gather per iteration (ionrad_3d.c:275,399,554,672); the `reduce` callbacks
 * of those entry points are then passed NULL.  c = NULL detaches. */
int ab2_grid_set_comm(ab2_grid* g, ab2_comm* c, int lo_rank, int hi_rank);
/* the exchange alone (SMR across GPUs: parent cells rewritten by ionradRestrictCorrect after that level's bvals_mhd) */
int ab2_exchange_x3(ab2_grid* g);

#ifdef __cplusplus
}
#endif
#endif
