/* athena_shim.c -- reference-side binding: defines the reference's hot-path symbols on top of the C ABI of
 * libathena_b200.so, so the UNMODIFIED driver (main.c, par.c, init_mesh.c, init_grid.c, output*.c, dump_*.c,
 * restart.c, problem files) links against the B200 engine instead of
 *   new_dt.o bvals_mhd.o smr.o integrators/ *.o reconstruction/ *.o rsolvers/ *.o ionradiation/ *.o
 * (src/Makefile.in:10-120).  Compile this file with the reference's own headers on the include path and the same
 * defs.h as the rest of the build; link with
 *   -Wl,--wrap=Userwork_in_loop,--wrap=Userwork_after_loop,--wrap=dump_binary,--wrap=dump_history,
 *       --wrap=dump_tab_cons,--wrap=dump_tab_prim,--wrap=dump_vtk,--wrap=dump_restart,--wrap=output_pdf,
 *       --wrap=output_pgm,--wrap=output_ppm,--wrap=output_vtk,--wrap=output_tab   -lathena_b200
 * (shim/build_dropin.py does exactly this).
 *
 * Coherence: the device mirror is authoritative between coherence points.  Host U is uploaded at
 * the first hot-path call after problem()/restart_grids(), and downloaded (with EdgeFlux) only when an output
 * writer actually fires or Userwork_after_loop runs -- the --wrap hooks below.  Userwork_in_loop is replayed on
 * the device after a one-time discovery of the (static) cells it overwrites; if the discovery finds
 * state-dependent writes the shim falls back to a host round trip per step.
 *
 * Problem-file boundary functions (bvals_mhd_fun) are turned into device tables the same way when they are constants and
 * signed copies (inflow, reflection, outflow); dump_vtk gets its float32 big-endian fields straight from the device.
 * Both discoveries are re-validated against the real host function on a sparse schedule (steps 1, 4, 16, 64, ...)
 * and fall back to the host round trip on any mismatch; AB2_USERWORK=host / AB2_USER_BC=host / AB2_HOST_WRITERS=1 force
 * the host paths.
 *
 * Scope of this binding: one Grid per Domain per process.  Serial build: one GPU, any number of SMR levels (--enable-smr),
 * each child Grid with one parent Grid (the tables init_grid.c builds in CGrid[]/PGrid[] are used as they are).
 * --enable-mpi build: one process per GPU, the root Domain cut into x3 slabs (NGrid_x3 = ranks), NCCL inside the library.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "defs.h"
#include "athena.h"
#include "globals.h"
#include "prototypes.h"
#include "athena_b200.h"

#if defined(MHD) || !defined(ADIABATIC) || !defined(CARTESIAN) || defined(SPECIAL_RELATIVITY)
#error "athena_shim.c binds the adiabatic Cartesian hydro path only"
#endif
/* every other configure switch that would make the driver expect something this library does not compute */
#if defined(SINGLE_PREC)
#error "libathena_b200 is FP64: configure without --enable-single (pG->U and EdgeFlux are handed over as double*)"
#endif
#if !defined(CTU_INTEGRATOR)
#error "libathena_b200 replaces integrate_3d_ctu only: configure --with-integrator=ctu"
#endif
#if defined(FIRST_ORDER_FLUX_CORRECTION) || defined(SELF_GRAVITY) || defined(SHEARING_BOX) || defined(FARGO) || defined(PARTICLES) || defined(FEEDBACK)
#error "first-order flux correction, self-gravity, shearing box, FARGO and particles are not on the path this library replaces"
#endif
#if defined(RESISTIVITY) || defined(VISCOSITY) || defined(THERMAL_CONDUCTION) || defined(STELLAR_WIND)
#error "the explicit diffusion modules and the stellar-wind source are not on the path this library replaces"
#endif
#if (NSCALARS > 1)
#error "libathena_b200 carries at most one passive scalar (the neutral density of the ion build)"
#endif
_Static_assert(sizeof(Real) == 8, "Real must be double");
_Static_assert(sizeof(ConsS) == (5 + NSCALARS)*sizeof(double), "ConsS must be 5+NSCALARS packed doubles");
#if defined(MPI_PARALLEL) && defined(STATIC_MESH_REFINEMENT)
#error "MPI + SMR: the reference's equal-count decomposition of fine Domains is not bound (driver.SmrRunSlabs cuts all levels on common planes instead)"
#endif

#ifdef MPI_PARALLEL
/* More than one GPU: one process per GPU, the root Domain cut into x3 slabs (domain1/NGrid_x3 = ranks,
 * NGrid_x1 = NGrid_x2 = 1: rays run along x1 and are never cut).  MPI carries the control plane of the unmodified driver;
 * the hot path's exchanges and reductions are NCCL calls inside the library (ab2_grid_set_comm). */
static ab2_comm *the_comm = NULL;
static int my_rank = 0, n_ranks = 1;
static void comm_setup(void)
{
  char id[AB2_UNIQUE_ID_BYTES];
  int ndev;
  if (the_comm != NULL || n_ranks < 0) return;
  MPI_Comm_rank(MPI_COMM_WORLD, &my_rank);
  MPI_Comm_size(MPI_COMM_WORLD, &n_ranks);
  if (n_ranks == 1) { n_ranks = -1; return; }
  ndev = ab2_device_count();
  if (ndev < 1) ath_error("[athena_shim]: no CUDA device\n");
  memset(id, 0, sizeof id);
  if (my_rank == 0 && ab2_comm_get_unique_id(id) != 0) ath_error("[athena_shim]: %s\n", ab2_last_error());
  if (MPI_Bcast(id, AB2_UNIQUE_ID_BYTES, MPI_CHAR, 0, MPI_COMM_WORLD) != MPI_SUCCESS) ath_error("[athena_shim]: MPI_Bcast of the NCCL id failed\n");
  the_comm = ab2_comm_init_rank(n_ranks, my_rank, id, my_rank % ndev);
  if (the_comm == NULL) ath_error("[athena_shim]: %s\n", ab2_last_error());
}
#endif

typedef struct {
  GridS *pG;
  ab2_grid *h;
  int dev_valid;          /* device mirror holds the current state */
  int host_fresh;         /* host pG->U equals the device mirror */
} Mirror;

static Mirror mir[8];
static int nmir = 0;
static int ion_ready = 0;
static MeshS *the_mesh = NULL;     /* bvals_mhd(DomainS*) needs MeshS.BCFlag_*; recorded by bvals_mhd_init */

static void die(const char *where)
{
  ath_error("[%s]: %s\n", where, ab2_last_error());
}

static int nvar(void) { return 5 + NSCALARS; }

static Mirror *mirror_of(GridS *pG)
{
  int n;
  for (n=0; n<nmir; n++) if (mir[n].pG == pG) return &mir[n];
  if (nmir == 8) ath_error("[athena_shim]: too many Grids\n");
  {
    ab2_grid_desc d;
    memset(&d, 0, sizeof d);
    d.nx1 = pG->Nx[0]; d.nx2 = pG->Nx[1]; d.nx3 = pG->Nx[2];
    d.nscal = NSCALARS;
    d.dx1 = pG->dx1; d.dx2 = pG->dx2; d.dx3 = pG->dx3;
    d.x1min = pG->MinX[0]; d.x2min = pG->MinX[1]; d.x3min = pG->MinX[2];
    d.gamma = Gamma;
    d.device = 0;
#ifdef MPI_PARALLEL
    comm_setup();
    if (the_comm != NULL) d.device = my_rank % ab2_device_count();
#endif
    mir[nmir].pG = pG;
    mir[nmir].h = ab2_grid_create(&d);
    if (mir[nmir].h == NULL) die("ab2_grid_create");
#if defined(THIRD_ORDER_CHAR)    /* --with-order=3: reconstruction/lr_states_ppm.c */
    if (ab2_set_reconstruction(mir[nmir].h, 3) != 0) die("ab2_set_reconstruction");
#elif !defined(SECOND_ORDER_CHAR)
#error "libathena_b200 replaces the characteristic PLM and PPM reconstructions only (--with-order=2 or 3)"
#endif
#if defined(HLLC_FLUX)           /* the driver was configured --with-flux=hllc: rsolvers/hllc.c + lr_states_plm.c:257 */
    if (ab2_set_riemann_solver(mir[nmir].h, 1) != 0) die("ab2_set_riemann_solver");
#elif !defined(ROE_FLUX)
#error "libathena_b200 replaces the Roe and HLLC solvers only"
#elif !defined(H_CORRECTION)     /* Roe, configured without --enable-h-correction: etah is never set */
    if (ab2_set_riemann_solver(mir[nmir].h, 2) != 0) die("ab2_set_riemann_solver");
#endif
#ifdef MPI_PARALLEL
    if (the_comm != NULL) {
      /* GridS.lx3_id / rx3_id (init_grid.c:253-283, periodic wrap added by bvals_mhd_init below) are ranks of the Domain's
       * communicator; with one root Domain over all processes that is MPI_COMM_WORLD, the communicator NCCL was built on */
      if (pG->lx1_id >= 0 || pG->rx1_id >= 0 || pG->lx2_id >= 0 || pG->rx2_id >= 0)
        ath_error("[athena_shim]: decompose along x3 only (domain1/NGrid_x1 = NGrid_x2 = 1): the ion-plane rays run along x1\n");
      if (ab2_grid_set_comm(mir[nmir].h, the_comm, pG->lx3_id, pG->rx3_id) != 0) die("ab2_grid_set_comm");
    }
#endif
    /* page-lock the driver's arrays: coherence copies then run at PCIe speed straight into them */
    ab2_pin_host(&(pG->U[0][0][0]), (long long)sizeof(ConsS)*(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost));
    mir[nmir].dev_valid = 0; mir[nmir].host_fresh = 1;
  }
  return &mir[nmir++];
}

static void ensure_device(Mirror *m)
{
  if (m->dev_valid) return;
  if (ab2_upload_cons_aos(m->h, (const double*)&(m->pG->U[0][0][0])) != 0) die("ab2_upload_cons_aos");
#ifdef ION_RADIATION
  if (ion_ready && ab2_upload_edgeflux(m->h, (const double*)&(m->pG->EdgeFlux[0][0][0])) != 0) die("ab2_upload_edgeflux");
#endif
  m->dev_valid = 1; m->host_fresh = 1;
}

static void sync_host(Mirror *m)
{
  if (!m->dev_valid || m->host_fresh) return;
  if (ab2_download_cons_aos(m->h, (double*)&(m->pG->U[0][0][0])) != 0) die("ab2_download_cons_aos");
#ifdef ION_RADIATION
  if (ion_ready && ab2_download_edgeflux(m->h, (double*)&(m->pG->EdgeFlux[0][0][0])) != 0) die("ab2_download_edgeflux");
#endif
  m->host_fresh = 1;
}

static void sync_all_to_host(void) { int n; for (n=0; n<nmir; n++) sync_host(&mir[n]); }

static GridS *root_grid(MeshS *pM)
{
#ifndef STATIC_MESH_REFINEMENT
  if (pM->NLevels != 1 || pM->DomainsPerLevel[0] != 1) ath_error("[athena_shim]: this build has no SMR: unigrid meshes only\n");
#endif
  return pM->Domain[0][0].Grid;
}

/* loop over every Grid of the Mesh (one per Domain on this process) */
#define FOR_ALL_GRIDS(pM, pD_, pG_) \
  for (int nl_=0; nl_<(pM)->NLevels; nl_++) for (int nd_=0; nd_<(pM)->DomainsPerLevel[nl_]; nd_++) \
    if (((pD_) = &((pM)->Domain[nl_][nd_]), (pG_) = (pD_)->Grid) != NULL)

/* ---- integrators/integrate.c, reconstruction/ ------------------------------------------------------------- */
static void integrate_3d_ctu_b200(DomainS *pD)
{
  Mirror *m = mirror_of(pD->Grid);
  ensure_device(m);
  if (ab2_integrate_3d_ctu(m->h, pD->Grid->dt) != 0) die("integrate_3d_ctu");
  m->host_fresh = 0;
}

VDFun_t integrate_init(MeshS *pM)
{
  DomainS *pD; GridS *pG;
  (void)root_grid(pM);
  if (CourNo > 0.5) ath_error("[integrate_init]: CourNo=%e > 0.5 with 3D CTU integrator\n", CourNo);   /* integrate.c:66-68 */
  if (CoolingFunc != NULL) ath_error("[integrate_init]: CoolingFunc is not supported by the B200 engine\n");
  FOR_ALL_GRIDS(pM, pD, pG) {
    Mirror *m = mirror_of(pG);
    if (pG->Nx[0] <= 1 || pG->Nx[1] <= 1 || pG->Nx[2] <= 1) ath_error("[integrate_init]: the B200 engine binds the 3D integrator only\n");
    if (ab2_set_static_grav_pot(m->h, StaticGravPot) != 0) die("ab2_set_static_grav_pot");
  }
  return integrate_3d_ctu_b200;
}
void integrate_destruct(void) { int n; for (n=0; n<nmir; n++) { ab2_grid_destroy(mir[n].h); mir[n].h = NULL; } nmir = 0; }
void lr_states_init(MeshS *pM) { (void)pM; }
void lr_states_destruct(void) {}

/* ---- new_dt.c ------------------------------------------------------------------------------------------------ */
void new_dt(MeshS *pM)
{ /* new_dt.c:58-195: one dt for every level */
  DomainS *pD; GridS *pG;
  Real max_dti = 0.0, tlim, v[3], x;
  Real max_v[3] = {0.0, 0.0, 0.0};   /* new_dt.c:53: declared outside the Domain loops and never reset, so each Grid's
                                      * speed is the running maximum over the Grids before it (coarser levels) too */
  FOR_ALL_GRIDS(pM, pD, pG) {
    Mirror *m = mirror_of(pG);
    ensure_device(m);
    if (ab2_max_wave_speeds(m->h, v) != 0) die("new_dt");
    if (pG->Nx[0] > 1) { max_v[0] = MAX(max_v[0], v[0]); x = max_v[0]/pG->dx1; max_dti = MAX(max_dti, x); }
    if (pG->Nx[1] > 1) { max_v[1] = MAX(max_v[1], v[1]); x = max_v[1]/pG->dx2; max_dti = MAX(max_dti, x); }
    if (pG->Nx[2] > 1) { max_v[2] = MAX(max_v[2], v[2]); x = max_v[2]/pG->dx3; max_dti = MAX(max_dti, x); }
  }
  if (pM->nstep == 0) pM->dt = CourNo/max_dti;
  else pM->dt = MIN(2.0*pM->dt, CourNo/max_dti);
  tlim = par_getd("time","tlim");
  if ((pM->time < tlim) && ((tlim - pM->time) < pM->dt)) pM->dt = tlim - pM->time;
  FOR_ALL_GRIDS(pM, pD, pG) pG->dt = pM->dt;
}

/* ---- bvals_mhd.c ------------------------------------------------------------------------------------------- */
/* steps at which a discovered device table is checked against the real host function again: 1, 4, 16, 64, 256, ...
 * (a check costs a full download plus the host function: ~0.5 s at 256^3, i.e. 2 s over a 200-step run, nothing over 10^5 steps) */
static int revalidation_step(int nstep) { return nstep >= 1 && (nstep & (nstep-1)) == 0 && (nstep & 0x55555555) != 0; }

/* A problem-file boundary function in table form.  The function is run on two synthetic host states whose every value
 * encodes its own (cell, variable): a ghost value that decodes, in BOTH runs, to +-(another cell of the same variable)
 * is a signed copy; a value that is the same in both runs is a constant; anything else (arithmetic on the state, NaN)
 * makes the function "not tabular" and it keeps its host round trip.  Host U is saved and restored around the probes. */
typedef struct { long n; long *dst, *src; double *coef; } BcTable;
static BcTable bc_tab[8][6];
static int bc_mode[8][6];              /* 0 undiscovered, 1 device table, 2 host round trip */
static int bc_checked_step[8][6];

static void bc_free(BcTable *t) { free(t->dst); free(t->src); free(t->coef); t->dst = t->src = NULL; t->coef = NULL; t->n = 0; }

static int bc_derive(GridS *pG, VGFun_t fn, BcTable *out)
{
  const int nv = nvar();
  const long ncell = (long)(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost);
  const long nw = ncell*nv;
  double *U = (double*)&(pG->U[0][0][0]);
  double *save = (double*)malloc(sizeof(double)*nw), *r1 = (double*)malloc(sizeof(double)*nw);
  char *is_dst = (char*)calloc((size_t)ncell, 1);
  long c, w, n = 0, t;
  int v, ok = 1;
  if (!save || !r1 || !is_dst) ath_error("[bvals_mhd]: malloc failed\n");
  memcpy(save, U, sizeof(double)*nw);
  for (w=0; w<nw; w++) U[w] = (double)(w+1);
  (*fn)(pG);
  memcpy(r1, U, sizeof(double)*nw);
  for (w=0; w<nw; w++) U[w] = 1.0e12 + 3.0*(double)(w+1);
  (*fn)(pG);
  for (c=0; c<ncell; c++) {
    int changed = 0;
    for (v=0; v<nv; v++) if (r1[c*nv+v] != (double)(c*nv+v+1)) changed = 1;
    if (changed) { is_dst[c] = 1; n++; }
  }
  out->n = n;
  out->dst = (long*)malloc(sizeof(long)*(n+1)); out->src = (long*)malloc(sizeof(long)*(n+1)*nv); out->coef = (double*)malloc(sizeof(double)*(n+1)*nv);
  for (c=0, t=0; c<ncell && ok; c++) {
    if (!is_dst[c]) continue;
    out->dst[t] = c;
    for (v=0; v<nv; v++) {
      const double x1 = r1[c*nv+v], x2 = U[c*nv+v];
      if (x1 == x2) { out->src[v*n+t] = -1; out->coef[v*n+t] = x1; }                /* constant */
      else {
        const double a = fabs(x1);
        const long idx = (long)a - 1;
        const double sg = (x1 < 0.0) ? -1.0 : 1.0;
        if (!(a >= 1.0 && a <= (double)nw && (double)(idx+1) == a && idx % nv == v && x2 == sg*(1.0e12 + 3.0*(double)(idx+1)))) { ok = 0; break; }
        if (is_dst[idx/nv] && idx/nv != c) { ok = 0; break; }                        /* reads a cell the function also writes */
        out->src[v*n+t] = idx/nv; out->coef[v*n+t] = sg;
      }
    }
    t++;
  }
  memcpy(U, save, sizeof(double)*nw);
  free(save); free(r1); free(is_dst);
  if (!ok) bc_free(out);
  return ok;
}

static int bc_same(const BcTable *a, const BcTable *b)
{
  const int nv = nvar();
  return a->n == b->n && memcmp(a->dst, b->dst, sizeof(long)*a->n) == 0 && memcmp(a->src, b->src, sizeof(long)*a->n*nv) == 0
      && memcmp(a->coef, b->coef, sizeof(double)*a->n*nv) == 0;
}
void bvals_mhd_init(MeshS *pM)
{
  (void)root_grid(pM);
  the_mesh = pM;
#ifdef MPI_PARALLEL
  { /* bvals_mhd.c:768-774, 814-820: with periodic x3 the edge Grids of the Domain are each other's neighbours */
    DomainS *pD = &(pM->Domain[0][0]);
    GridS *pG = pD->Grid;
    int myL, myM, myN;
    if (pM->NLevels != 1 || pM->DomainsPerLevel[0] != 1) ath_error("[bvals_mhd_init]: the MPI binding drives one root Domain\n");
    if (pG != NULL && pD->NGrid[2] > 1) {
      get_myGridIndex(pD, myID_Comm_world, &myL, &myM, &myN);
      if (pM->BCFlag_ix3 == 4 && pG->lx3_id < 0) pG->lx3_id = pD->GData[pD->NGrid[2]-1][myM][myL].ID_Comm_Domain;
      if (pM->BCFlag_ox3 == 4 && pG->rx3_id < 0) pG->rx3_id = pD->GData[0][myM][myL].ID_Comm_Domain;
    }
  }
#endif
}

void bvals_mhd_fun(DomainS *pD, enum BCDirection dir, VGFun_t prob_bc)
{ /* bvals_mhd.c:917-947 */
  switch(dir){
  case left_x1:  pD->ix1_BCFun = prob_bc; break;
  case right_x1: pD->ox1_BCFun = prob_bc; break;
  case left_x2:  pD->ix2_BCFun = prob_bc; break;
  case right_x2: pD->ox2_BCFun = prob_bc; break;
  case left_x3:  pD->ix3_BCFun = prob_bc; break;
  case right_x3: pD->ox3_BCFun = prob_bc; break;
  default: ath_perr(-1,"[bvals_fun]: Unknown direction = %d\n",dir); exit(EXIT_FAILURE);
  }
}

void bvals_mhd(DomainS *pD)
{
  GridS *pG = pD->Grid;
  Mirror *m = mirror_of(pG);
  MeshS *pM = the_mesh;
  VGFun_t user[6];
  int flags[6], bc[6], dir, side, any_user = 0;
  if (pM == NULL) ath_error("[bvals_mhd]: bvals_mhd_init was not called\n");
  user[0] = pD->ix1_BCFun; user[1] = pD->ox1_BCFun; user[2] = pD->ix2_BCFun;
  user[3] = pD->ox2_BCFun; user[4] = pD->ix3_BCFun; user[5] = pD->ox3_BCFun;
  flags[0] = pM->BCFlag_ix1; flags[1] = pM->BCFlag_ox1; flags[2] = pM->BCFlag_ix2;
  flags[3] = pM->BCFlag_ox2; flags[4] = pM->BCFlag_ix3; flags[5] = pM->BCFlag_ox3;
  { /* Domain faces interior to the root are prolongated later, not BC'd (bvals_mhd.c:546-560 ProlongateLater) */
    int a, irefine = 1;
    for (a=0; a<pD->Level; a++) irefine *= 2;
    for (a=0; a<3; a++) {
      if (pD->Disp[a] != 0) flags[2*a] = 0;
      if ((pD->Disp[a] + pD->Nx[a])/irefine != pM->Nx[a]) flags[2*a+1] = 0;
    }
  }
#ifdef MPI_PARALLEL
  if (pG->lx3_id >= 0) flags[4] = 0;      /* faces shared with a neighbouring Grid: filled by the NCCL exchange inside ab2_bvals_mhd */
  if (pG->rx3_id >= 0) flags[5] = 0;
#endif
  for (dir=0; dir<6; dir++) if (user[dir] != NULL) any_user = 1;
  ensure_device(m);
  if (any_user) {
    /* problem-file boundary functions: in table form on the device where they are constants and signed copies */
    const int mi = (int)(m - mir);
    const char *force = getenv("AB2_USER_BC");
    int all_tab = 1;
    for (dir=0; dir<6; dir++) {
      BcTable t;
      if (user[dir] == NULL) continue;
      if (force != NULL && strcmp(force, "host") == 0) bc_mode[mi][dir] = 2;
      if (bc_mode[mi][dir] == 0 || (bc_mode[mi][dir] == 1 && revalidation_step(pM->nstep) && bc_checked_step[mi][dir] != pM->nstep)) {
        sync_host(m);                                  /* the probes run on (and restore) the host array */
        bc_checked_step[mi][dir] = pM->nstep;
        if (!bc_derive(pG, user[dir], &t)) {
          if (bc_mode[mi][dir] == 1) ath_pout(0,"[athena_b200]: boundary function of face %d is no longer tabular: host round trip from now on\n", dir);
          else ath_pout(0,"[athena_b200]: boundary function of face %d is not tabular (it computes on the state): host round trip\n", dir);
          bc_mode[mi][dir] = 2;
        } else if (bc_mode[mi][dir] == 1) {
          if (!bc_same(&t, &bc_tab[mi][dir])) {
            ath_pout(0,"[athena_b200]: boundary function of face %d changed its table (time-dependent?): host round trip from now on\n", dir);
            bc_mode[mi][dir] = 2;
          }
          bc_free(&t);
        } else {
          if (ab2_set_bc_table(m->h, dir, t.n, t.dst, t.src, t.coef) != 0) die("ab2_set_bc_table");
          bc_tab[mi][dir] = t;
          bc_mode[mi][dir] = 1;
          ath_pout(0,"[athena_b200]: boundary function of face %d runs on the device (%ld ghost cells in table form)\n", dir, t.n);
        }
      }
      if (bc_mode[mi][dir] != 1) all_tab = 0;
    }
    if (all_tab) {
      for (dir=0; dir<6; dir++) if (user[dir] != NULL) flags[dir] = -2;
      any_user = 0;
    }
  }
  if (!any_user) {
    if (ab2_bvals_mhd(m->h, flags) != 0) die("bvals_mhd");
    m->host_fresh = 0;
    return;
  }
  /* problem-file BC functions (bvals_mhd_fun) run on the host, direction by direction to keep the corner-fill order */
  for (dir=0; dir<3; dir++) {
    for (side=0; side<6; side++) bc[side] = 0;
    for (side=0; side<2; side++) if (user[2*dir+side] == NULL) bc[2*dir+side] = flags[2*dir+side];
    if (ab2_bvals_mhd(m->h, bc) != 0) die("bvals_mhd");
    m->host_fresh = 0;
    if (user[2*dir] != NULL || user[2*dir+1] != NULL) {
      sync_host(m);
      if (user[2*dir] != NULL) (*user[2*dir])(pG);
      if (user[2*dir+1] != NULL) (*user[2*dir+1])(pG);
      m->dev_valid = 0;
      ensure_device(m);
    }
  }
}

#ifdef STATIC_MESH_REFINEMENT
/* ---- smr.c --------------------------------------------------------------------------------------------------- */
static void overlap_from(const GridOvrlpS *pCO, const DomainS *pChildD, ab2_overlap *o)
{ /* the reference's own tables (init_grid.c:395-660), as plain ints */
  int a;
  for (a=0; a<3; a++) { o->cs[a] = pCO->ijks[a]; o->ce[a] = pCO->ijke[a]; o->cdisp[a] = pChildD->Disp[a]; }
  for (a=0; a<6; a++) o->face[a] = (pCO->myFlx[a] != NULL);
}

typedef int (*pair_fn)(ab2_grid*, ab2_grid*, const ab2_overlap*, double);
static int pair_prolong(ab2_grid *p, ab2_grid *c, const ab2_overlap *o, double dt) { (void)dt; return ab2_smr_prolongate(p, c, o); }
static int pair_ion_restrict(ab2_grid *p, ab2_grid *c, const ab2_overlap *o, double dt) { (void)dt; return ab2_smr_ion_restrict(p, c, o); }

/* apply fn to every (parent Grid, child Grid) pair of level nl -> nl+1 */
static void for_pairs(MeshS *pM, int nl, pair_fn fn, const char *who)
{
  int nd, ncg;
  for (nd=0; nd<pM->DomainsPerLevel[nl]; nd++) {
    GridS *pG = pM->Domain[nl][nd].Grid;
    if (pG == NULL) continue;
    for (ncg=0; ncg<pG->NCGrid; ncg++) {
      GridOvrlpS *pCO = &(pG->CGrid[ncg]);
      DomainS *pCD = &(pM->Domain[nl+1][pCO->DomN]);
      Mirror *mp = mirror_of(pG), *mc = mirror_of(pCD->Grid);
      ab2_overlap o;
      overlap_from(pCO, pCD, &o);
      ensure_device(mp); ensure_device(mc);
      if (fn(mp->h, mc->h, &o, pG->dt) != 0) die(who);
      mp->host_fresh = 0; mc->host_fresh = 0;
    }
  }
}

void SMR_init(MeshS *pM) { (void)pM; }
void RestrictCorrect(MeshS *pM)
{ /* smr.c:1233: from the finest level down, so a level is restricted after it received its children */
  int nl;
  for (nl=pM->NLevels-2; nl>=0; nl--) for_pairs(pM, nl, ab2_smr_restrict_correct, "RestrictCorrect");
}
void ionradRestrictCorrect(MeshS *pM)
{
  int nl;
  for (nl=pM->NLevels-2; nl>=0; nl--) for_pairs(pM, nl, pair_ion_restrict, "ionradRestrictCorrect");
}
void Prolongate(MeshS *pM)
{ /* smr.c:2386: from the root up */
  int nl;
  for (nl=0; nl<pM->NLevels-1; nl++) for_pairs(pM, nl, pair_prolong, "Prolongate");
}
#endif /* STATIC_MESH_REFINEMENT */

#ifdef ION_RADIATION
static Real tcoarse = 0.0;          /* ionrad_3d.c:47 */
/* ---- ionradiation/ ------------------------------------------------------------------------------------------ */
static void ion_radtransfer_3d_b200(DomainS *pD)
{
  GridS *pG = pD->Grid;
  MeshS *pM = pD->Mesh;
  Mirror *m = mirror_of(pG);
  int niter = 0, nchem = 0, ntherm = 0;
  double dt = pG->dt;
  if ((pM->radplanelist)->dir[0] != -1) ath_error("[ion_radtransfer_3d]: the B200 engine binds dir = -1 only (the one direction the reference's ray loop indexes correctly)\n");
  ensure_device(m);
  if (pD->Level == 0) {
    if (ab2_ion_radtransfer_3d(m->h, (pM->radplanelist)->flux_i, pM->time, CourNo, &dt, &niter, &nchem, &ntherm, NULL, NULL) != 0) die("ion_radtransfer_3d");
    tcoarse = ab2_ion_coarse_time(m->h);                       /* ionrad_3d.c:1028 */
  }
#ifdef STATIC_MESH_REFINEMENT
  else { /* ionrad_prolong_snd/rcv (ionrad_smr.c) from the parent Grid, then the fine-level sub-cycle to tcoarse */
    GridOvrlpS *pPO = &(pG->PGrid[0]);
    GridS *pPG = pM->Domain[pD->Level-1][pPO->DomN].Grid;
    Mirror *mp = mirror_of(pPG);
    ab2_overlap o;
    int n;
    for (n=0; n<pPG->NCGrid; n++) if (pPG->CGrid[n].DomN == pD->DomNumber) break;
    overlap_from(&(pPG->CGrid[n]), pD, &o);
    ensure_device(mp);
    if (ab2_smr_ion_prolong(mp->h, m->h, &o) != 0) die("ionrad_prolong");
    if (ab2_ion_radtransfer_3d_fine(m->h, tcoarse, &dt, &niter, &nchem, &ntherm, NULL, NULL) != 0) die("ion_radtransfer_3d");
  }
#endif
  m->host_fresh = 0;
  pG->dt = dt;
  pM->dt = pG->dt;                                             /* ionrad_3d.c:1033 */
  fprintf(stderr, "Radiation done in %d iterations: %d thermal, %d chemical; new dt = %e\n", niter, ntherm, nchem, pG->dt);   /* :1041 */
}

void ion_radtransfer_init_domain(MeshS *pM) { (void)pM; }

VDFun_t ion_radtransfer_init(MeshS *pM, int ires)
{
  DomainS *pD; GridS *pG;
  ab2_ion_par p;
  (void)ires;
  p.sigma_ph = par_getd("ionradiation", "sigma_ph");            /* ionrad_3d.c:742-757 */
  p.m_H = par_getd("ionradiation", "m_H");
  p.mu = par_getd("ionradiation", "mu");
  p.e_gamma = par_getd("ionradiation", "e_gamma");
  p.alpha_C = par_getd("ionradiation", "alpha_C");
  p.k_B = par_getd("ionradiation", "k_B");
  p.time_unit = par_getd("ionradiation", "time_unit");
  p.max_de_iter = par_getd("ionradiation", "max_de_iter");
  p.max_de_therm_iter = par_getd("ionradiation", "max_de_therm_iter");
  p.max_dx_iter = par_getd("ionradiation", "max_dx_iter");
  p.max_de_step = par_getd("ionradiation", "max_de_step");
  p.max_de_therm_step = par_getd("ionradiation", "max_de_therm_step");
  p.max_dx_step = par_getd("ionradiation", "max_dx_step");
  p.tfloor = par_getd("ionradiation", "tfloor");
  p.tceil = par_getd("ionradiation", "tceil");
  p.maxiter = par_getd("ionradiation", "maxiter");
  ab2_ion_derived(&p, pM->dx[0], pM->dx[1], pM->dx[2]);         /* ionrad.c:115-131 */
  (void)root_grid(pM);
  FOR_ALL_GRIDS(pM, pD, pG) {                                  /* one parameter set for every level (ionrad.c:94-135) */
    Mirror *m = mirror_of(pG);
    if (ab2_ion_init(m->h, &p) != 0) die("ion_radtransfer_init");
    if (m->dev_valid && ab2_upload_edgeflux(m->h, (const double*)&(pG->EdgeFlux[0][0][0])) != 0) die("ab2_upload_edgeflux");
  }
  ion_ready = 1;
  return ion_radtransfer_3d_b200;
}

void clear_coarse_time(void) {}
void set_coarse_time(void) {}

void add_radplane_3d(GridS *pGrid, int dir, Real flux)
{ /* ionradplane_3d.c:53-81 */
  MeshS *pMesh = pGrid->Mesh;
  (pMesh->radplanelist)->dir[0] = dir;
  (pMesh->radplanelist)->flux_i = flux;
}

void bvals_ionrad_init(MeshS *pM) { (void)pM; }

void bvals_ionrad(DomainS *pD)
{ /* bvals_ionrad.c:63-146 with outflow_flux_ix1 (:308-332): host array, read back on the first upload.  The BC
   * function pointers there are file-static and set by the FIRST Domain (:27-29,178), so every Domain gets the
   * root's outflow_flux_ix1. */
  GridS *pG = pD->Grid;
  MeshS *pM = pG->Mesh;
  int j, k;
  if ((pM->radplanelist)->dir[0] != -1) return;
  for (k=0; k<=pG->Nx[2]; k++)
    for (j=0; j<=pG->Nx[1]; j++)
      pG->EdgeFlux[k][j][0] = (pM->radplanelist)->flux_i;
}
#endif /* ION_RADIATION */

/* ---- --wrap hooks: coherence points ------------------------------------------------------------------------ */
#define WRAP_WRITER(name) \
  void __real_##name(MeshS *pM, OutputS *pOut); \
  void __wrap_##name(MeshS *pM, OutputS *pOut) { sync_all_to_host(); __real_##name(pM, pOut); }
WRAP_WRITER(dump_binary)
WRAP_WRITER(dump_history)
WRAP_WRITER(dump_tab_cons)
WRAP_WRITER(dump_tab_prim)
WRAP_WRITER(dump_restart)
WRAP_WRITER(output_pdf)
WRAP_WRITER(output_pgm)
WRAP_WRITER(output_ppm)
WRAP_WRITER(output_vtk)
WRAP_WRITER(output_tab)

/* dump_vtk (dump_vtk.c:28-330): the driver-side half -- file name, header, section labels, fwrite -- with the data half
 * done on the device: each CELL_DATA field arrives as the float32 big-endian bytes the reference would have produced from
 * the host mirror (ab2_output_vtk_field), 24 B/zone over PCIe instead of the 48 B/zone mirror, and no host conversion loop.
 * Builds that write ghost cells, and Grids whose device mirror is not current, use the reference's own writer. */
void __real_dump_vtk(MeshS *pM, OutputS *pOut);
void __wrap_dump_vtk(MeshS *pM, OutputS *pOut)
{
#ifdef WRITE_GHOST_CELLS
  sync_all_to_host(); __real_dump_vtk(pM, pOut);
#else
  int nl, nd, prim, fld;
  if (getenv("AB2_HOST_WRITERS") != NULL || (strcmp(pOut->out,"cons") != 0 && strcmp(pOut->out,"prim") != 0)) {
    sync_all_to_host(); __real_dump_vtk(pM, pOut); return;
  }
  prim = (strcmp(pOut->out,"prim") == 0);
  for (nl=0; nl<pM->NLevels; nl++) for (nd=0; nd<pM->DomainsPerLevel[nl]; nd++) {
    GridS *pG = pM->Domain[nl][nd].Grid;
    Mirror *m;
    FILE *fp;
    char *fname, *plev = NULL, *pdom = NULL, levstr[8], domstr[8];
    long nz;
    float *data;
    if (pG == NULL) continue;
    if (!((pOut->nlevel == -1 || pOut->nlevel == nl) && (pOut->ndomain == -1 || pOut->ndomain == nd))) continue;
    m = mirror_of(pG);
    ensure_device(m);
    nz = (long)pG->Nx[0]*pG->Nx[1]*pG->Nx[2];
    if (nl > 0) { plev = levstr; sprintf(plev, "lev%d", nl); }
    if (nd > 0) { pdom = domstr; sprintf(pdom, "dom%d", nd); }
    if ((fname = ath_fname(plev, pM->outfilename, plev, pdom, num_digit, pOut->num, NULL, "vtk")) == NULL)
      ath_error("[dump_vtk]: Error constructing filename\n");
    if ((fp = fopen(fname, "w")) == NULL) { ath_error("[dump_vtk]: Unable to open vtk dump file\n"); return; }
    if ((data = (float*)malloc(sizeof(float)*3*nz)) == NULL) { ath_error("[dump_vtk]: malloc failed for temporary array\n"); return; }
    fprintf(fp, "# vtk DataFile Version 2.0\n");
    fprintf(fp, "%s vars at time= %e, level= %i, domain= %i\n", prim ? "PRIMITIVE" : "CONSERVED", pG->time, nl, nd);
    fprintf(fp, "BINARY\n");
    fprintf(fp, "DATASET STRUCTURED_POINTS\n");
    fprintf(fp, "DIMENSIONS %d %d %d\n", pG->Nx[0]+1, pG->Nx[1]+1, pG->Nx[2]+1);
    fprintf(fp, "ORIGIN %e %e %e \n", pG->MinX[0], pG->MinX[1], pG->MinX[2]);
    fprintf(fp, "SPACING %e %e %e \n", pG->dx1, pG->dx2, pG->dx3);
    fprintf(fp, "CELL_DATA %d \n", (int)nz);
    for (fld=0; fld<3+NSCALARS; fld++) {
      if (fld == 0) fprintf(fp, "SCALARS density float\nLOOKUP_TABLE default\n");
      if (fld == 1) fprintf(fp, "\nVECTORS %s float\n", prim ? "velocity" : "momentum");
      if (fld == 2) fprintf(fp, "\nSCALARS %s float\nLOOKUP_TABLE default\n", prim ? "pressure" : "total_energy");
      if (fld == 3) fprintf(fp, "\nSCALARS %s[%d] float\nLOOKUP_TABLE default\n", prim ? "specific_scalar" : "scalar", 0);
      if (ab2_output_vtk_field(m->h, fld, prim, data) != 0) die("dump_vtk");
      fwrite(data, sizeof(float), (size_t)((fld == 1) ? 3*nz : nz), fp);
    }
    fclose(fp);
    free(data);
    free(fname);
  }
#endif
}

void __real_Userwork_after_loop(MeshS *pM);
void __wrap_Userwork_after_loop(MeshS *pM) { sync_all_to_host(); __real_Userwork_after_loop(pM); }

/* Userwork_in_loop (main.c:597): discover once which cells the problem file overwrites and with what; if the
 * writes do not depend on the state, replay them on the device from then on. */
void __real_Userwork_in_loop(MeshS *pM);
static long uw_n[8];               /* the replay tables as discovered, kept for the re-validation */
static long *uw_cells[8];
static double *uw_vals[8];

void __wrap_Userwork_in_loop(MeshS *pM)
{
  static int mode = 0;          /* 0 undiscovered, 1 device replay, 2 host round trip */
  const int nv = nvar();
  DomainS *pD; GridS *pG;
  {
    const char *force = getenv("AB2_USERWORK");
    if (force != NULL && strcmp(force, "host") == 0) mode = 2;
  }
  if (mode == 1 && revalidation_step(pM->nstep)) {
    /* The discovery below proves that the writes do not depend on the STATE at one step; it cannot see a dependence on
     * time, step number or the problem file's statics (a forcing that switches on later, a ramp).  So on a sparse schedule
     * the real function runs again on the current host state and must change exactly the table's cells to the table's
     * values; otherwise the replay is abandoned for the host round trip.  (This also lets a hook that only reads the
     * state -- diagnostics -- see fresh data on those steps.) */
    int g = 0, same = 1;
    double *before[8];
    sync_all_to_host();
    FOR_ALL_GRIDS(pM, pD, pG) {
      const long ncell = (long)(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost);
      before[g] = (double*)malloc(sizeof(double)*ncell*nv);
      if (!before[g]) ath_error("[Userwork_in_loop]: malloc failed\n");
      memcpy(before[g], (double*)&(pG->U[0][0][0]), sizeof(double)*ncell*nv);
      g++;
    }
    __real_Userwork_in_loop(pM);                              /* one call serves every Grid of the Mesh */
    g = 0;
    FOR_ALL_GRIDS(pM, pD, pG) {
      const long ncell = (long)(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost);
      double *U = (double*)&(pG->U[0][0][0]);
      long n; int v;
      /* every table cell must now hold the table's values; put those into `before` as well, so that ... */
      for (n=0; n<uw_n[g] && same; n++) for (v=0; v<nv; v++) {
        if (U[uw_cells[g][n]*nv+v] != uw_vals[g][(long)v*uw_n[g]+n]) { same = 0; break; }
        before[g][uw_cells[g][n]*nv+v] = U[uw_cells[g][n]*nv+v];
      }
      /* ... one memcmp shows that no cell outside the table changed */
      if (same && memcmp(U, before[g], sizeof(double)*ncell*nv) != 0) same = 0;
      free(before[g]);
      g++;
    }
    if (!same) {
      mode = 2;
      ath_pout(0,"[athena_b200]: Userwork_in_loop no longer matches its replay table (time-dependent?): host round trip from now on\n");
      FOR_ALL_GRIDS(pM, pD, pG) mirror_of(pG)->dev_valid = 0;     /* the host holds the true result */
      return;
    }
  }
  if (mode == 1) {
    FOR_ALL_GRIDS(pM, pD, pG) {
      Mirror *m = mirror_of(pG);
      if (ab2_apply_fixed_cells(m->h) != 0) die("Userwork_in_loop replay");
      m->host_fresh = 0;
    }
    return;
  }
  sync_all_to_host();
  if (mode == 2) {
    __real_Userwork_in_loop(pM);
    FOR_ALL_GRIDS(pM, pD, pG) mirror_of(pG)->dev_valid = 0;
    return;
  }
  {
    double *orig[8], *pert[8], *out2[8];
    long ncell[8], c, nfix_total = 0;
    int g = 0, ng, v, ok = 1;
    FOR_ALL_GRIDS(pM, pD, pG) {
      double *U = (double*)&(pG->U[0][0][0]);
      ncell[g] = (long)(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost);
      orig[g] = (double*)malloc(sizeof(double)*ncell[g]*nv); pert[g] = (double*)malloc(sizeof(double)*ncell[g]*nv);
      out2[g] = (double*)malloc(sizeof(double)*ncell[g]*nv);
      if (!orig[g] || !pert[g] || !out2[g]) ath_error("[Userwork_in_loop]: malloc failed\n");
      memcpy(orig[g], U, sizeof(double)*ncell[g]*nv);
      /* run 2 works on a perturbed state: every value moved, signs kept, zero momenta made non-zero */
      for (c=0; c<ncell[g]; c++) for (v=0; v<nv; v++)
        pert[g][c*nv+v] = orig[g][c*nv+v]*1.25 + ((v>=1 && v<=3) ? 1.0e-3*fabs(orig[g][c*nv]) : 0.0);
      memcpy(U, pert[g], sizeof(double)*ncell[g]*nv);
      g++;
    }
    ng = g;
    __real_Userwork_in_loop(pM);
    g = 0;
    FOR_ALL_GRIDS(pM, pD, pG) {
      double *U = (double*)&(pG->U[0][0][0]);
      memcpy(out2[g], U, sizeof(double)*ncell[g]*nv);
      memcpy(U, orig[g], sizeof(double)*ncell[g]*nv);
      g++;
    }
    __real_Userwork_in_loop(pM);                  /* run 1: the real state; its result is what the host keeps */
    g = 0;
    FOR_ALL_GRIDS(pM, pD, pG) {
      double *U = (double*)&(pG->U[0][0][0]);
      for (c=0; c<ncell[g]; c++) {
        int w2 = 0, w1 = 0;
        for (v=0; v<nv; v++) { if (out2[g][c*nv+v] != pert[g][c*nv+v]) w2 = 1; if (U[c*nv+v] != orig[g][c*nv+v]) w1 = 1; }
        if (w1 && !w2) ok = 0;
        if (w2) for (v=0; v<nv; v++) {
          if (out2[g][c*nv+v] == pert[g][c*nv+v]) ok = 0;                 /* partially written cell */
          if (out2[g][c*nv+v] != U[c*nv+v]) ok = 0;                       /* value depends on the state */
        }
      }
      g++;
    }
    if (ok) {
      g = 0;
      FOR_ALL_GRIDS(pM, pD, pG) {
        Mirror *m = mirror_of(pG);
        long nfix = 0, n, *cells; double *vals;
        for (c=0; c<ncell[g]; c++) { int w2 = 0; for (v=0; v<nv; v++) if (out2[g][c*nv+v] != pert[g][c*nv+v]) w2 = 1; nfix += w2; }
        cells = (long*)malloc(sizeof(long)*(nfix+1)); vals = (double*)calloc((size_t)6*(nfix+1), sizeof(double));
        nfix = 0;
        for (c=0; c<ncell[g]; c++) { int w2 = 0; for (v=0; v<nv; v++) if (out2[g][c*nv+v] != pert[g][c*nv+v]) w2 = 1; if (w2) cells[nfix++] = c; }
        for (n=0; n<nfix; n++) for (v=0; v<nv; v++) vals[(long)v*nfix+n] = out2[g][cells[n]*nv+v];
        if (ab2_set_fixed_cells(m->h, nfix, cells, vals) != 0) die("ab2_set_fixed_cells");
        uw_n[g] = nfix; uw_cells[g] = cells; uw_vals[g] = vals;        /* kept: the re-validation compares against them */
        nfix_total += nfix;
        g++;
      }
      mode = 1;
      ath_pout(0,"[athena_b200]: Userwork_in_loop replayed on the device (%ld fixed cells)\n", nfix_total);
    } else {
      mode = 2;
      ath_pout(0,"[athena_b200]: Userwork_in_loop depends on the state: host round trip every step\n");
    }
    for (g=0; g<ng; g++) { free(orig[g]); free(pert[g]); free(out2[g]); }
    FOR_ALL_GRIDS(pM, pD, pG) mirror_of(pG)->dev_valid = 0;   /* host holds the result of run 1: re-upload at the next call */
  }
}
