/* athena_oracle.h -- CPU restatement (plain C) of the reference's per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (atmospheric-athena_b200/) never links, imports
 * or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this restatement against restart dumps
 * written by the reference's own serial build (oracle/_ref, built by oracle/build_ref.py from the
 * sources where they lie under /root/reference) -- committed as tests/golden/ *.npz together with the
 * generating script tests/golden/make_golden.py -- and against function-level vectors produced by the
 * reference's own lr_states()/fluxes() objects (tests/golden/units_*.npz).  Also pinned the same way: the SMR
 * functions on 2- and 3-level nests (the 3-level dumps come from the reference's MPI branch run as one rank against
 * oracle/mpi_stub -- its serial build segfaults there), the HLLC solver with the HLL-family PLM tracing, Roe without
 * H-correction, and PPM (lr_states_ppm.c) including the reference's work-array aliasing for the passive scalar.
 *
 * Data layout: SoA.  One double array per conserved variable, extents [n3][n2][n1] INCLUDING ghost
 * zones (n = Nx + 2*4 in every direction), i fastest.  Variable order: d, M1, M2, M3, E, s0.
 * With nscal = 0 the s pointer may be NULL.
 */
#ifndef ATHENA_ORACLE_H
#define ATHENA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NGHOST 4

typedef double (*OrcGravPot)(double x1, double x2, double x3);

typedef struct OrcGrid {
  int nx1, nx2, nx3;   /* active zones                                   */
  int nscal;           /* 0 (blast) or 1 (ion configs)                   */
  double dx1, dx2, dx3;
  double x1min, x2min, x3min; /* GridS.MinX (cc_pos.c)                   */
  double gamma;        /* ratio of specific heats (globals.h: Gamma)     */
  double *d, *M1, *M2, *M3, *E, *s; /* [n3][n2][n1] with ghosts          */
  OrcGravPot grav;     /* StaticGravPot (globals.h:24) or NULL           */
  double *Fsave[3];    /* optional: final x1/x2/x3 fluxes of the last orc_integrate_3d_ctu call, [cell][6] in
                        * PHYSICAL component order (d,M1,M2,M3,E,s) -- what step 12e (integrate_3d_ctu.c:3076-3364)
                        * copies into CGrid/PGrid[].myFlx for SMR.  NULL: not kept.                             */
} OrcGrid;

/* Overlap of ONE child Grid with ONE parent Grid (GridOvrlpS, athena.h:258-275; built by init_grid.c:300-1180),
 * single Grid per Domain: cs/ce = CGrid.ijks/ijke on the PARENT (local indices incl. ghosts); the child's own
 * overlap is its whole active region.  face[6]: 1 where the fine/coarse boundary is interior to the root Domain
 * (myFlx[dim] != NULL, init_grid.c:508-510, 624-626).  cdisp = child Domain Disp (child-level cells). */
typedef struct OrcOverlap {
  int cs[3], ce[3];
  int face[6];
  int cdisp[3];
} OrcOverlap;

typedef struct OrcIonPar {           /* <ionradiation> block, ionrad_3d.c:742-757 */
  double sigma_ph, m_H, mu, e_gamma, alpha_C, k_B, time_unit;
  double max_de_iter, max_de_therm_iter, max_dx_iter;
  double max_de_step, max_de_therm_step, max_dx_step;
  double tfloor, tceil;
  int maxiter;
  double min_area, d_nlo;            /* ionrad.c:115-131 */
} OrcIonPar;

/* bvals_mhd.c:174-508 with physical boundaries on every face.  bc[6] = ix1,ox1,ix2,ox2,ix3,ox3 with the
 * reference's flag values 1 reflect, 2 outflow, 4 periodic (5 conductor == reflect for hydro). */
void orc_bvals(OrcGrid *g, const int bc[6]);

/* new_dt.c:69-139: max over active zones of |v_d| + sqrt(gamma P / d); out[3]. */
void orc_max_speeds(const OrcGrid *g, double out[3]);
/* new_dt.c:155-185: the scalar dt logic. */
double orc_new_dt(const OrcGrid *g, double cour_no, double dt_old, int nstep, double time, double tlim);

/* integrate_3d_ctu.c:110-3368, hydro + H-correction (+ static potential) path. */
void orc_integrate_3d_ctu(OrcGrid *g, double dt);

/* pencil-level pieces, exposed for unit tests.  W/Wl/Wr are arrays of 6 doubles (d,Vx,Vy,Vz,P,r).
 * lr_states_plm.c:62-374: loops cells il-1..iu+1, writes Wl[i+1], Wr[i]. */
void orc_lr_states(const double (*W)[6], int nscal, double gamma, double dt, double dx, int il, int iu,
                   double (*Wl)[6], double (*Wr)[6]);
/* roe.c:59-336 (+ hlle.c:62-268 fallback); U = (d,Mx,My,Mz,E,s), W = (d,Vx,Vy,Vz,P,r). */
void orc_roe_flux(const double Ul[6], const double Ur[6], const double Wl[6], const double Wr[6],
                  int nscal, double gamma, double etah, double F[6]);
/* convert_var.c:389-461 */
void orc_cons_to_prim(const double U[6], int nscal, double gamma, double W[6]);
void orc_prim_to_cons(const double W[6], int nscal, double gamma, double U[6]);

/* bvals_ionrad.c:63-146 with outflow_flux_ix1 (:308-332): root Grid at the L-edge, dir = -1:
 * EdgeFlux[k][j][0] = flux_i over the whole (Nx3+1)(Nx2+1) upstream plane (called once, main.c:432). */
void orc_bvals_ionrad(const OrcGrid *g, double flux_i, double *edgeflux);

/* batch helpers for the unit tests: n independent interfaces / states, rows of 6 doubles */
/* rsolvers/hllc.c:50-261; orc_set_riemann_solver(1) makes orc_integrate_3d_ctu use it (the reference's
 * --with-flux=hllc build, configured without H-correction: etah lives in roe.c), 0 = Roe + H-correction (default),
 * 2 = Roe built without --enable-h-correction (etah stays 0). */
void orc_hllc_flux_batch(int n, const double *Ul, const double *Ur, const double *Wl, const double *Wr,
                         int nscal, double gamma, double *F);
void orc_set_riemann_solver(int kind);
/* lr_states(): 2 = reconstruction/lr_states_plm.c (default), 3 = reconstruction/lr_states_ppm.c:67-651 (--with-order=3).
 * PPM reads W over [il-3, iu+3]. */
void orc_set_reconstruction(int order);
void orc_roe_flux_batch(int n, const double *Ul, const double *Ur, const double *Wl, const double *Wr,
                        int nscal, double gamma, const double *etah, double *F);
void orc_prim_to_cons_batch(int n, const double *W, int nscal, double gamma, double *U);
void orc_cons_to_prim_batch(int n, const double *U, int nscal, double gamma, double *W);

/* ionrad.c:115-131 */
void orc_ion_derived(OrcIonPar *p, double dx1, double dx2, double dx3);
/* ionradplane_3d.c:265 (root level, dir=+-1): entering flux at mesh time t */
double orc_ion_entering_flux(double flux_i, double time);
/* ion_radtransfer_3d (ionrad_3d.c:862-1047), root level, one radiation plane with dir=-1.
 * edgeflux: [nx3+1][nx2+1][nx1+1].  *dt is pGrid->dt on entry and on exit.  Returns niter. */
int orc_ion_radtransfer(OrcGrid *g, const OrcIonPar *p, double flux_i, double time, double cour_no,
                        double *dt, double *edgeflux, int *nchem, int *ntherm);

/* as orc_ion_radtransfer, also returning tcoarse = dt_done (ionrad_3d.c:1028) for the finer levels */
int orc_ion_radtransfer_root(OrcGrid *g, const OrcIonPar *p, double flux_i, double time, double cour_no,
                             double *dt, double *edgeflux, int *nchem, int *ntherm, double *tcoarse);

/* ---- SMR (smr.c, ionrad_smr.c), one parent + one child on the same process --------------------------- */
/* RestrictCorrect (smr.c:1207-2352): restrict child solution + boundary fluxes, inject, flux-correct the parent.
 * Needs Fsave on both grids (zeros before the first step, as the reference's calloc'ed myFlx). */
void orc_restrict_correct(OrcGrid *parent, const OrcGrid *child, const OrcOverlap *o, double dt_parent);
/* ionradRestrictCorrect (smr.c:85-1205): E and s[0] only, no flux correction */
void orc_ion_restrict(OrcGrid *parent, const OrcGrid *child, const OrcOverlap *o);
/* Prolongate (smr.c:2359-2923) with ProCon (:3068-3325), mcd_slope (:3478-3495) */
void orc_prolongate(const OrcGrid *parent, OrcGrid *child, const OrcOverlap *o);
/* ionrad_prolong_snd + ionrad_prolong_rcv (ionrad_smr.c:345-459, 34-343, serial branch), dim = 0 */
void orc_ion_prolong(const OrcGrid *parent, const double *edgeflux_parent, const OrcGrid *child,
                     double *edgeflux_child, const OrcOverlap *o);
/* ion_radtransfer_3d on a fine level (ionrad_3d.c:919-1014 with finegrid = 1): sub-cycles to tcoarse with the
 * entering flux of each ray taken from edgeflux[k][j][0].  *dt returns GridS.dt = dt_done.  Returns niter. */
int orc_ion_radtransfer_fine(OrcGrid *g, const OrcIonPar *p, double tcoarse, double *dt, double *edgeflux,
                             int *nchem, int *ntherm);

#ifdef __cplusplus
}
#endif
#endif
