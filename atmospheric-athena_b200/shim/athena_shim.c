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
 * Scope of this binding: one Grid per process, root level only (unigrid).  SMR symbols are not provided yet.
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
#if defined(STATIC_MESH_REFINEMENT) || defined(MPI_PARALLEL)
#error "athena_shim.c: SMR / MPI builds are not bound yet (SURVEY 8 rows a24-a27, 8e use the Python host for now)"
#endif

typedef struct {
  GridS *pG;
  ab2_grid *h;
  int dev_valid;          /* device mirror holds the current state */
  int host_fresh;         /* host pG->U equals the device mirror */
  int userwork_mode;      /* 0 undiscovered, 1 device replay, 2 host round trip */
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
    mir[nmir].pG = pG;
    mir[nmir].h = ab2_grid_create(&d);
    if (mir[nmir].h == NULL) die("ab2_grid_create");
    mir[nmir].dev_valid = 0; mir[nmir].host_fresh = 1; mir[nmir].userwork_mode = 0;
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
  if (pM->NLevels != 1 || pM->DomainsPerLevel[0] != 1) ath_error("[athena_shim]: unigrid meshes only\n");
  return pM->Domain[0][0].Grid;
}

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
  GridS *pG = root_grid(pM);
  Mirror *m = mirror_of(pG);
  if (pG->Nx[0] <= 1 || pG->Nx[1] <= 1 || pG->Nx[2] <= 1) ath_error("[integrate_init]: the B200 engine binds the 3D integrator only\n");
  if (CourNo > 0.5) ath_error("[integrate_init]: CourNo=%e > 0.5 with 3D CTU integrator\n", CourNo);   /* integrate.c:66-68 */
  if (CoolingFunc != NULL) ath_error("[integrate_init]: CoolingFunc is not supported by the B200 engine\n");
  if (ab2_set_static_grav_pot(m->h, StaticGravPot) != 0) die("ab2_set_static_grav_pot");
  return integrate_3d_ctu_b200;
}
void integrate_destruct(void) { int n; for (n=0; n<nmir; n++) { ab2_grid_destroy(mir[n].h); mir[n].h = NULL; } nmir = 0; }
void lr_states_init(MeshS *pM) { (void)pM; }
void lr_states_destruct(void) {}

/* ---- new_dt.c ------------------------------------------------------------------------------------------------ */
void new_dt(MeshS *pM)
{
  GridS *pG = root_grid(pM);
  Mirror *m = mirror_of(pG);
  double dt;
  ensure_device(m);
  if (ab2_new_dt(m->h, CourNo, pM->dt, pM->nstep, pM->time, par_getd("time","tlim"), NULL, NULL, &dt) != 0) die("new_dt");
  pM->dt = dt;
  pG->dt = dt;
}

/* ---- bvals_mhd.c ------------------------------------------------------------------------------------------- */
void bvals_mhd_init(MeshS *pM) { (void)root_grid(pM); the_mesh = pM; }

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
  for (dir=0; dir<6; dir++) if (user[dir] != NULL) any_user = 1;
  ensure_device(m);
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

#ifdef ION_RADIATION
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
  if (ab2_ion_radtransfer_3d(m->h, (pM->radplanelist)->flux_i, pM->time, CourNo, &dt, &niter, &nchem, &ntherm, NULL, NULL) != 0) die("ion_radtransfer_3d");
  m->host_fresh = 0;
  pG->dt = dt;
  pM->dt = pG->dt;                                             /* ionrad_3d.c:1033 */
  fprintf(stderr, "Radiation done in %d iterations: %d thermal, %d chemical; new dt = %e\n", niter, ntherm, nchem, pG->dt);   /* :1041 */
}

void ion_radtransfer_init_domain(MeshS *pM) { (void)pM; }

VDFun_t ion_radtransfer_init(MeshS *pM, int ires)
{
  GridS *pG = root_grid(pM);
  Mirror *m = mirror_of(pG);
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
  if (ab2_ion_init(m->h, &p) != 0) die("ion_radtransfer_init");
  ion_ready = 1;
  if (m->dev_valid && ab2_upload_edgeflux(m->h, (const double*)&(pG->EdgeFlux[0][0][0])) != 0) die("ab2_upload_edgeflux");
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
{ /* bvals_ionrad.c:63-146 with outflow_flux_ix1 (:308-332): host array, read back on the first upload */
  GridS *pG = pD->Grid;
  MeshS *pM = pG->Mesh;
  int j, k;
  if ((pM->radplanelist)->dir[0] != -1 || pD->Disp[0] != 0) return;
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
WRAP_WRITER(dump_vtk)
WRAP_WRITER(dump_restart)
WRAP_WRITER(output_pdf)
WRAP_WRITER(output_pgm)
WRAP_WRITER(output_ppm)
WRAP_WRITER(output_vtk)
WRAP_WRITER(output_tab)

void __real_Userwork_after_loop(MeshS *pM);
void __wrap_Userwork_after_loop(MeshS *pM) { sync_all_to_host(); __real_Userwork_after_loop(pM); }

/* Userwork_in_loop (main.c:597): discover once which cells the problem file overwrites and with what; if the
 * writes do not depend on the state, replay them on the device from then on. */
void __real_Userwork_in_loop(MeshS *pM);
void __wrap_Userwork_in_loop(MeshS *pM)
{
  GridS *pG = root_grid(pM);
  Mirror *m = mirror_of(pG);
  const int nv = nvar();
  const long ncell = (long)(pG->Nx[0]+2*nghost)*(pG->Nx[1]+2*nghost)*(pG->Nx[2]+2*nghost);
  double *U = (double*)&(pG->U[0][0][0]);
  if (m->userwork_mode == 1) {
    if (ab2_apply_fixed_cells(m->h) != 0) die("Userwork_in_loop replay");
    m->host_fresh = 0;
    return;
  }
  sync_host(m);
  if (m->userwork_mode == 2) {
    __real_Userwork_in_loop(pM);
    m->dev_valid = 0;
    return;
  }
  {
    double *orig = (double*)malloc(sizeof(double)*ncell*nv), *pert = (double*)malloc(sizeof(double)*ncell*nv);
    double *out2 = (double*)malloc(sizeof(double)*ncell*nv);
    long c, nfix = 0, *cells; double *vals; int v, ok = 1;
    if (!orig || !pert || !out2) ath_error("[Userwork_in_loop]: malloc failed\n");
    memcpy(orig, U, sizeof(double)*ncell*nv);
    /* run 2 on a perturbed state (every value moved, signs kept, zeros made non-zero) */
    for (c=0; c<ncell; c++) for (v=0; v<nv; v++) pert[c*nv+v] = orig[c*nv+v]*1.25 + ((v>=1 && v<=3) ? 1.0e-3*fabs(orig[c*nv]) : 0.0);
    memcpy(U, pert, sizeof(double)*ncell*nv);
    __real_Userwork_in_loop(pM);
    memcpy(out2, U, sizeof(double)*ncell*nv);
    /* run 1 on the real state */
    memcpy(U, orig, sizeof(double)*ncell*nv);
    __real_Userwork_in_loop(pM);
    for (c=0; c<ncell; c++) {
      int w2 = 0, w1 = 0;
      for (v=0; v<nv; v++) { if (out2[c*nv+v] != pert[c*nv+v]) w2 = 1; if (U[c*nv+v] != orig[c*nv+v]) w1 = 1; }
      if (w2) { nfix++; for (v=0; v<nv; v++) if (out2[c*nv+v] != U[c*nv+v] && out2[c*nv+v] != pert[c*nv+v]) ok = 0; }
      if (w1 && !w2) ok = 0;
      if (w2) for (v=0; v<nv; v++) if (out2[c*nv+v] == pert[c*nv+v]) ok = 0;     /* partially written cell */
    }
    if (ok) {
      cells = (long*)malloc(sizeof(long)*(nfix+1)); vals = (double*)calloc((size_t)6*(nfix+1), sizeof(double));
      nfix = 0;
      for (c=0; c<ncell; c++) {
        int w2 = 0;
        for (v=0; v<nv; v++) if (out2[c*nv+v] != pert[c*nv+v]) w2 = 1;
        if (w2) { cells[nfix++] = c; }
      }
      { long n; for (n=0; n<nfix; n++) for (v=0; v<nv; v++) vals[(long)v*nfix+n] = out2[cells[n]*nv+v]; }
      if (ab2_set_fixed_cells(m->h, nfix, cells, vals) != 0) die("ab2_set_fixed_cells");
      free(cells); free(vals);
      m->userwork_mode = 1;
      ath_pout(0,"[athena_b200]: Userwork_in_loop replayed on the device (%ld fixed cells)\n", nfix);
    } else {
      m->userwork_mode = 2;
      ath_pout(0,"[athena_b200]: Userwork_in_loop depends on the state: host round trip every step\n");
    }
    free(orig); free(pert); free(out2);
    m->dev_valid = 0;          /* host now holds the result of run 1: re-upload at the next hot-path call */
  }
}
