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
 * Scope of this binding: one process, one GPU, one Grid per Domain; any number of SMR levels (--enable-smr), each
 * child Grid with one parent Grid (the tables init_grid.c builds in CGrid[]/PGrid[] are used as they are).
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
#if defined(MPI_PARALLEL)
#error "athena_shim.c: MPI builds are not bound (multi-GPU runs use the Python host with NCCL, INTEGRATION.md)"
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
  { /* Domain faces interior to the root are prolongated later, not BC'd (bvals_mhd.c:546-560 ProlongateLater) */
    int a, irefine = 1;
    for (a=0; a<pD->Level; a++) irefine *= 2;
    for (a=0; a<3; a++) {
      if (pD->Disp[a] != 0) flags[2*a] = 0;
      if ((pD->Disp[a] + pD->Nx[a])/irefine != pM->Nx[a]) flags[2*a+1] = 0;
    }
  }
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
  static int mode = 0;          /* 0 undiscovered, 1 device replay, 2 host round trip */
  const int nv = nvar();
  DomainS *pD; GridS *pG;
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
        free(cells); free(vals);
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
