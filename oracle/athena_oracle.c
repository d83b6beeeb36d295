/* athena_oracle.c -- CPU restatement (plain C) of the reference's per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see athena_oracle.h).  Never linked into the product.
 *
 * Written from the algorithm in /root/reference/src (citations per function), for the configuration
 *   HYDRO ADIABATIC CARTESIAN CTU_INTEGRATOR ROE_FLUX H_CORRECTION SECOND_ORDER_CHAR DOUBLE_PREC
 *   NSCALARS in {0,1}, optional static potential, ION_RADIATION + ION_RADPLANE (dir = -1).
 * Operation order is kept identical to the reference wherever it is observable in floating point, so
 * that against the reference's own gcc -O3 (no FMA) build the hydro path agrees bit for bit.
 * Compile with:  gcc -O2 -ffp-contract=off -fPIC -shared athena_oracle.c -o libathena_oracle.so -lm
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "athena_oracle.h"

#define NG ORC_NGHOST
#define NWAVE 5
#define TINY_NUMBER 1.0e-20              /* defs.h.in:160 */
#define MIN(a,b) ( ((a) < (b)) ? (a) : (b) )   /* defs.h.in:146 */
#define MAX(a,b) ( ((a) > (b)) ? (a) : (b) )   /* defs.h.in:149 */
#define SIGN(a) ( ((a) < 0.) ? -1. : 1. )      /* defs.h.in:151 */
#define SQR(x) ( (x)*(x) )

typedef struct { double d, Mx, My, Mz, E, s; } C1;   /* Cons1DS, athena.h:148-166 */
typedef struct { double d, Vx, Vy, Vz, P, r; } P1;   /* Prim1DS, athena.h:172-188 */

/* ------------------------------------------------------------------------------------------------
 * convert_var.c:389-421 / 432-461 / 470-499
 * ---------------------------------------------------------------------------------------------- */
static P1 cons1d_to_prim1d(const C1 *U, int nscal, double Gamma_1)
{
  P1 W;
  double di = 1.0/U->d;
  W.d  = U->d;
  W.Vx = U->Mx*di;
  W.Vy = U->My*di;
  W.Vz = U->Mz*di;
  W.P = U->E - 0.5*(SQR(U->Mx)+SQR(U->My)+SQR(U->Mz))*di;
  W.P *= Gamma_1;
  W.P = MAX(W.P,TINY_NUMBER);
  W.r = nscal ? U->s*di : 0.0;
  return W;
}

static C1 prim1d_to_cons1d(const P1 *W, int nscal, double Gamma_1)
{
  C1 U;
  U.d  = W->d;
  U.Mx = W->d*W->Vx;
  U.My = W->d*W->Vy;
  U.Mz = W->d*W->Vz;
  U.E = W->P/Gamma_1 + 0.5*W->d*(SQR(W->Vx) + SQR(W->Vy) + SQR(W->Vz));
  U.s = nscal ? W->r*W->d : 0.0;
  return U;
}

static double cfast(const C1 *U, double Gamma, double Gamma_1)
{
  double p = Gamma_1*(U->E - 0.0 - 0.5*(SQR(U->Mx)+SQR(U->My)+SQR(U->Mz))/U->d);
  double asq = Gamma*p/U->d;
  return sqrt(asq);
}

void orc_cons_to_prim(const double U[6], int nscal, double gamma, double W[6])
{
  C1 u; P1 w;
  memcpy(&u, U, sizeof u);
  w = cons1d_to_prim1d(&u, nscal, gamma - 1.0);
  memcpy(W, &w, sizeof w);
}

void orc_prim_to_cons(const double W[6], int nscal, double gamma, double U[6])
{
  C1 u; P1 w;
  memcpy(&w, W, sizeof w);
  u = prim1d_to_cons1d(&w, nscal, gamma - 1.0);
  memcpy(U, &u, sizeof u);
}

/* ------------------------------------------------------------------------------------------------
 * esystem_prim.c:120-199
 * ---------------------------------------------------------------------------------------------- */
static void esys_prim_adb_hyd(double d, double v1, double rho_a2, double ev[NWAVE],
                              double rem[NWAVE][NWAVE], double lem[NWAVE][NWAVE])
{
  double asq = rho_a2/d;
  double a = sqrt(asq);
  ev[0] = v1 - a; ev[1] = v1; ev[2] = v1; ev[3] = v1; ev[4] = v1 + a;

  rem[0][0] = 1.0; rem[1][0] = -a/d; rem[4][0] = asq;
  rem[0][1] = 1.0;
  rem[2][2] = 1.0;
  rem[3][3] = 1.0;
  rem[0][4] = 1.0; rem[1][4] = -rem[1][0]; rem[4][4] = asq;

  lem[0][1] = -0.5*d/a; lem[0][4] = 0.5/asq;
  lem[1][0] = 1.0;      lem[1][4] = -1.0/asq;
  lem[2][2] = 1.0;
  lem[3][3] = 1.0;
  lem[4][1] = -lem[0][1]; lem[4][4] = lem[0][4];
}

/* ------------------------------------------------------------------------------------------------
 * esystem_roe.c:132-215
 * ---------------------------------------------------------------------------------------------- */
static void esys_roe_adb_hyd(double v1, double v2, double v3, double h, double Gamma_1, double ev[NWAVE],
                             double rem[NWAVE][NWAVE], double lem[NWAVE][NWAVE])
{
  double vsq, asq, a, na, qa;
  vsq = v1*v1 + v2*v2 + v3*v3;
  asq = Gamma_1*MAX((h-0.5*vsq), TINY_NUMBER);
  a = sqrt(asq);
  ev[0] = v1 - a; ev[1] = v1; ev[2] = v1; ev[3] = v1; ev[4] = v1 + a;
  if (rem == NULL || lem == NULL) return;

  rem[0][0] = 1.0; rem[1][0] = v1 - a; rem[2][0] = v2; rem[3][0] = v3; rem[4][0] = h - v1*a;
  rem[2][1] = 1.0; rem[4][1] = v2;
  rem[3][2] = 1.0; rem[4][2] = v3;
  rem[0][3] = 1.0; rem[1][3] = v1; rem[2][3] = v2; rem[3][3] = v3; rem[4][3] = 0.5*vsq;
  rem[0][4] = 1.0; rem[1][4] = v1 + a; rem[2][4] = v2; rem[3][4] = v3; rem[4][4] = h + v1*a;

  na = 0.5/asq;
  lem[0][0] = na*(0.5*Gamma_1*vsq + v1*a);
  lem[0][1] = -na*(Gamma_1*v1 + a);
  lem[0][2] = -na*Gamma_1*v2;
  lem[0][3] = -na*Gamma_1*v3;
  lem[0][4] = na*Gamma_1;
  lem[1][0] = -v2; lem[1][2] = 1.0;
  lem[2][0] = -v3; lem[2][3] = 1.0;
  qa = Gamma_1/asq;
  lem[3][0] = 1.0 - na*Gamma_1*vsq;
  lem[3][1] = qa*v1;
  lem[3][2] = qa*v2;
  lem[3][3] = qa*v3;
  lem[3][4] = -qa;
  lem[4][0] = na*(0.5*Gamma_1*vsq - v1*a);
  lem[4][1] = -na*(Gamma_1*v1 - a);
  lem[4][2] = lem[0][2];
  lem[4][3] = lem[0][3];
  lem[4][4] = lem[0][4];
}

/* ------------------------------------------------------------------------------------------------
 * lr_states_plm.c:62-374  (CARTESIAN: gamma_curv = 0 => opg = omg = beta = betai = 1)
 * ---------------------------------------------------------------------------------------------- */
/* which Riemann solver the build is configured with: 0 = roe.c with --enable-h-correction, 1 = hllc.c (--with-flux=hllc),
 * 2 = roe.c without H-correction (the integrator never sets etah, integrate_3d_ctu.c:2299-2426 compiled out).
 * It selects fluxes() AND the form of PLM steps 8-9 (lr_states_plm.c:257 vs :296). */
static int orc_riemann = 0;

static int orc_order = 2;      /* 2 = lr_states_plm.c (SECOND_ORDER_CHAR), 3 = lr_states_ppm.c (THIRD_ORDER_CHAR) */
static void lr_states_ppm(const P1 W[], int nscal, double Gamma, double dt, double dx, int il, int iu, P1 Wl[], P1 Wr[]);
static void lr_states_plm(const P1 W[], int nscal, double Gamma, double dt, double dx, int il, int iu, P1 Wl[], P1 Wr[]);
static void lr_states(const P1 W[], int nscal, double Gamma, double dt, double dx, int il, int iu,
                      P1 Wl[], P1 Wr[])
{
  if (orc_order == 3) lr_states_ppm(W, nscal, Gamma, dt, dx, il, iu, Wl, Wr);
  else lr_states_plm(W, nscal, Gamma, dt, dx, il, iu, Wl, Wr);
}

static void lr_states_plm(const P1 W[], int nscal, double Gamma, double dt, double dx, int il, int iu,
                          P1 Wl[], P1 Wr[])
{
  int i, n, m;
  const int NV = NWAVE + nscal;
  double lim_slope1, lim_slope2, qa, qx;
  double ev[NWAVE], rem[NWAVE][NWAVE], lem[NWAVE][NWAVE];
  double dWc[NWAVE+1], dWl[NWAVE+1], dWr[NWAVE+1], dWg[NWAVE+1];
  double dac[NWAVE+1], dal[NWAVE+1], dar[NWAVE+1], dag[NWAVE+1], da[NWAVE+1];
  double Wlv[NWAVE+1], Wrv[NWAVE+1], dW[NWAVE+1], dWm[NWAVE+1];
  double *pWl, *pWr;
  double qx1, qx2, C;
  const double opg = 1.0, omg = 1.0, beta = 1.0, betai = 1.0;
  const double dtodx = dt/dx;

  for (n=0; n<NWAVE; n++) for (m=0; m<NWAVE; m++) { rem[n][m] = 0.0; lem[n][m] = 0.0; }

  for (i=il-1; i<=iu+1; i++) {
    const double *pWm = (const double*)&W[i-1], *pW0 = (const double*)&W[i], *pWp = (const double*)&W[i+1];

    esys_prim_adb_hyd(W[i].d, W[i].Vx, Gamma*W[i].P, ev, rem, lem);            /* :105 */

    for (n=0; n<NV; n++) {                                                      /* :131-147 */
      dWc[n] = pWp[n] - pWm[n];
      dWl[n] = pW0[n] - pWm[n];
      dWr[n] = pWp[n] - pW0[n];
      if (dWl[n]*dWr[n] > 0.0) dWg[n] = 2.0*dWl[n]*dWr[n]/(dWl[n]+dWr[n]);
      else dWg[n] = 0.0;
    }

    for (n=0; n<NWAVE; n++) {                                                   /* :152-163 */
      dac[n] = lem[n][0]*dWc[0];
      dal[n] = lem[n][0]*dWl[0];
      dar[n] = lem[n][0]*dWr[0];
      dag[n] = lem[n][0]*dWg[0];
      for (m=1; m<NWAVE; m++) {
        dac[n] += lem[n][m]*dWc[m];
        dal[n] += lem[n][m]*dWl[m];
        dar[n] += lem[n][m]*dWr[m];
        dag[n] += lem[n][m]*dWg[m];
      }
    }
    for (n=NWAVE; n<NV; n++) { dac[n] = dWc[n]; dal[n] = dWl[n]; dar[n] = dWr[n]; dag[n] = dWg[n]; }

    for (n=0; n<NV; n++) {                                                      /* :180-187 */
      da[n] = 0.0;
      if (dal[n]*dar[n] > 0.0) {
        lim_slope1 = MIN(    fabs(dal[n]),fabs(dar[n]));
        lim_slope2 = MIN(0.5*fabs(dac[n]),fabs(dag[n]));
        da[n] = SIGN(dac[n])*MIN(2.0*lim_slope1,lim_slope2);
      }
    }

    for (n=0; n<NWAVE; n++) {                                                   /* :192-197 */
      dWm[n] = da[0]*rem[n][0];
      for (m=1; m<NWAVE; m++) dWm[n] += da[m]*rem[n][m];
    }
    for (n=NWAVE; n<NV; n++) dWm[n] = da[n];

    for (n=0; n<NV; n++) {                                                      /* :222-225 */
      Wlv[n] = pW0[n] - 0.5*dWm[n]*opg;
      Wrv[n] = pW0[n] + 0.5*dWm[n]*omg;
    }
    for (n=0; n<NV; n++) {                                                      /* :227-236 */
      C = Wrv[n] + beta*Wlv[n];
      Wlv[n] = MAX(MIN(pW0[n],pWm[n]),Wlv[n]);
      Wlv[n] = MIN(MAX(pW0[n],pWm[n]),Wlv[n]);
      Wrv[n] = C - beta*Wlv[n];
      Wrv[n] = MAX(MIN(pW0[n],pWp[n]),Wrv[n]);
      Wrv[n] = MIN(MAX(pW0[n],pWp[n]),Wrv[n]);
      Wlv[n] = (C - Wrv[n])*betai;
    }
    for (n=0; n<NV; n++) dW[n] = Wrv[n] - Wlv[n];

    pWl = (double*)&Wl[i+1];
    pWr = (double*)&Wr[i];

    if (orc_riemann == 1) {
      /* :257-293, the HLLE/HLLC/HLLD form of steps 8-9: start from the limited face values and subtract each wave
       * moving TOWARDS the interface.  The passive scalar is not touched by this branch at all: the reference leaves
       * Wl[i+1].r / Wr[i].r as they are in its malloc'ed pencils (integrate_3d_ctu.c:3438-3439) -- uninitialised
       * memory, zero in practice on a fresh heap.  Restated as zero; see DESIGN.md "HLLC". */
      for (n=0; n<NWAVE; n++) { pWl[n] = Wrv[n]; pWr[n] = Wlv[n]; }
      for (n=0; n<NWAVE; n++) {
        qa = 0.0;
        if (ev[n] > 0.) {
          qx = 0.5*dtodx*ev[n];
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*qx*dW[m];
          for (m=0; m<NWAVE; m++) pWl[m] -= qa*rem[m][n];
        }
        qa = 0.0;
        if (ev[n] < 0.) {
          qx = 0.5*dtodx*ev[n];
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*qx*dW[m];
          for (m=0; m<NWAVE; m++) pWr[m] -= qa*rem[m][n];
        }
      }
      Wl[i+1].r = 0.0; Wr[i].r = 0.0;
      continue;
    }

    qx = 0.5*MAX(ev[NWAVE-1],0.0)*dtodx;                                        /* :297-304 */
    for (n=0; n<NV; n++) pWl[n] = Wrv[n] - qx*dW[n];
    qx = -0.5*MIN(ev[0],0.0)*dtodx;                                             /* :306-313 */
    for (n=0; n<NV; n++) pWr[n] = Wlv[n] + qx*dW[n];

    for (n=0; n<NWAVE; n++) {                                                   /* :322-339 */
      if (ev[n] >= 0.0) {
        qa  = 0.0;
        qx1 = 0.5*dtodx*ev[NWAVE-1];
        qx2 = 0.5*dtodx*ev[n];
        qx = qx1 - qx2;
        for (m=0; m<NWAVE; m++) qa += lem[n][m]*qx*dW[m];
        for (m=0; m<NWAVE; m++) pWl[m] += qa*rem[m][n];
      }
    }
    for (n=0; n<NWAVE; n++) {                                                   /* :341-358 */
      if (ev[n] <= 0.0) {
        qa = 0.0;
        qx1 = -0.5*dtodx*ev[0];
        qx2 = -0.5*dtodx*ev[n];
        qx = -qx1 + qx2;
        for (m=0; m<NWAVE; m++) qa += lem[n][m]*qx*dW[m];
        for (m=0; m<NWAVE; m++) pWr[m] += qa*rem[m][n];
      }
    }
    for (n=NWAVE; n<NV; n++) {                                                  /* :361-367 */
      if (W[i].Vx > 0.) pWl[n] += 0.5*dtodx*(ev[NWAVE-1]-W[i].Vx)*dW[n];
      else if (W[i].Vx < 0.) pWr[n] += 0.5*dtodx*(ev[0]-W[i].Vx)*dW[n];
    }
    if (!nscal) { Wl[i+1].r = 0.0; Wr[i].r = 0.0; }
  }
}

void orc_lr_states(const double (*W)[6], int nscal, double gamma, double dt, double dx, int il, int iu,
                   double (*Wl)[6], double (*Wr)[6])
{
  lr_states((const P1*)W, nscal, gamma, dt, dx, il, iu, (P1*)Wl, (P1*)Wr);
}

/* ------------------------------------------------------------------------------------------------
 * lr_states_ppm.c:67-651 (THIRD_ORDER_CHAR, the reference's --with-order=3; HYDRO, ADIABATIC, CARTESIAN, CTU)
 * ---------------------------------------------------------------------------------------------- */
#define FOUR_3RDS 1.333333333333333          /* defs.h.in:159 (sic: 16 digits) */
#define TWO_3RDS  0.6666666666666667         /* defs.h.in:158 */

/* Steps 1-5 (== 8-12): eigensystem of cell i and its monotonised characteristic slopes, back in primitive variables */
static void ppm_tvd_slope(const P1 W[], int i, int nscal, double Gamma, double ev[NWAVE],
                          double rem[NWAVE][NWAVE], double lem[NWAVE][NWAVE], double dWm[NWAVE+1])
{
  const int NV = NWAVE + nscal;
  const double *pWm = (const double*)&W[i-1], *pW0 = (const double*)&W[i], *pWp = (const double*)&W[i+1];
  double dWc[NWAVE+1], dWl[NWAVE+1], dWr[NWAVE+1], dWg[NWAVE+1];
  double dac[NWAVE+1], dal[NWAVE+1], dar[NWAVE+1], dag[NWAVE+1], da[NWAVE+1];
  double lim_slope1, lim_slope2;
  int n, m;
  esys_prim_adb_hyd(W[i].d, W[i].Vx, Gamma*W[i].P, ev, rem, lem);
  for (n=0; n<NV; n++) {                                                        /* :130-150 */
    dWc[n] = pWp[n] - pWm[n];
    dWl[n] = pW0[n] - pWm[n];
    dWr[n] = pWp[n] - pW0[n];
    if (dWl[n]*dWr[n] > 0.0) dWg[n] = 2.0*dWl[n]*dWr[n]/(dWl[n]+dWr[n]);
    else dWg[n] = 0.0;
  }
  for (n=0; n<NWAVE; n++) {                                                     /* :156-167 */
    dac[n] = lem[n][0]*dWc[0];
    dal[n] = lem[n][0]*dWl[0];
    dar[n] = lem[n][0]*dWr[0];
    dag[n] = lem[n][0]*dWg[0];
    for (m=1; m<NWAVE; m++) {
      dac[n] += lem[n][m]*dWc[m];
      dal[n] += lem[n][m]*dWl[m];
      dar[n] += lem[n][m]*dWr[m];
      dag[n] += lem[n][m]*dWg[m];
    }
  }
  for (n=NWAVE; n<NV; n++) { dac[n] = dWc[n]; dal[n] = dWl[n]; dar[n] = dWr[n]; dag[n] = dWg[n]; }
  for (n=0; n<NV; n++) {                                                        /* :184-191 */
    da[n] = 0.0;
    if (dal[n]*dar[n] > 0.0) {
      lim_slope1 = MIN(    fabs(dal[n]),fabs(dar[n]));
      lim_slope2 = MIN(0.5*fabs(dac[n]),fabs(dag[n]));
      da[n] = SIGN(dac[n])*MIN(2.0*lim_slope1,lim_slope2);
    }
  }
  for (n=0; n<NWAVE; n++) {                                                     /* :196-201 */
    dWm[n] = da[0]*rem[n][0];
    for (m=1; m<NWAVE; m++) dWm[n] += da[m]*rem[n][m];
  }
  for (n=NWAVE; n<NV; n++) dWm[n] = da[n];
}

static void lr_states_ppm(const P1 W[], int nscal, double Gamma, double dt, double dx, int il, int iu,
                          P1 Wl[], P1 Wr[])
{
  const int NV = NWAVE + nscal;
  const double dtodx = dt/dx;
  double ev[NWAVE], rem[NWAVE][NWAVE], lem[NWAVE][NWAVE];
  double ev_ip1[NWAVE], rem_ip1[NWAVE][NWAVE], lem_ip1[NWAVE][NWAVE];
  double Wlv[NWAVE+1], Wrv[NWAVE+1], dW[NWAVE+1], W6[NWAVE+1];
  /* The reference allocates both work arrays as nmax x NWAVE (lr_states_ppm.c:697-701) but indexes them up to
   * NWAVE+NSCALARS: in the contiguous block of calloc_2d_array, entry [i][NWAVE] IS entry [i+1][0].  With a passive
   * scalar its slope and face value therefore alias the next cell's density entries; the net effect per cell i is
   *   Wlv[scalar] = density face value at i+1/2,   Wrv[scalar] = .5(s[i+1]+s[i]) - (ds(i+1) - dd(i+1))/6,
   * which the monotonicity step mostly flattens to first order.  Reproduced by using the same flat NWAVE-stride layout. */
  double *dWm_flat = calloc((size_t)(iu+5)*NWAVE, sizeof(double));
  double *Wim1h_flat = calloc((size_t)(iu+5)*NWAVE, sizeof(double));
#define dWm(i) (dWm_flat + (size_t)(i)*NWAVE)
#define Wim1h(i) (Wim1h_flat + (size_t)(i)*NWAVE)
  double qa, qb, qc, qx1, qx2;
  double *pWl, *pWr;
  int i, n, m;

  for (n=0; n<NWAVE; n++) for (m=0; m<NWAVE; m++) { rem[n][m] = lem[n][m] = rem_ip1[n][m] = lem_ip1[n][m] = 0.0; }
  for (i=il-2; i<=il-1; i++) ppm_tvd_slope(W, i, nscal, Gamma, ev, rem, lem, dWm(i));       /* :107-253 */
  {                                                                                         /* :259-271 */
    const double *p1 = (const double*)&W[il-1], *p2 = (const double*)&W[il-2];
    for (n=0; n<NV; n++) Wim1h(il-1)[n] = .5*(p1[n]+p2[n]) - (dWm(il-1)[n]-dWm(il-2)[n])/6.;
  }
  for (i=il-1; i<=iu+1; i++) {
    const double *pWm = (const double*)&W[i-1], *pW0 = (const double*)&W[i], *pWp = (const double*)&W[i+1];
    ppm_tvd_slope(W, i+1, nscal, Gamma, ev_ip1, rem_ip1, lem_ip1, dWm(i+1));                /* :287-427 */
    for (n=0; n<NV; n++) Wim1h(i+1)[n] = 0.5*(pWp[n]+pW0[n]) - (dWm(i+1)[n]-dWm(i)[n])/6.0; /* :434-446 */

    for (n=0; n<NV; n++) { Wlv[n] = Wim1h(i)[n]; Wrv[n] = Wim1h(i+1)[n]; }                  /* :451-454 */
    for (n=0; n<NV; n++) {                                                                  /* :465-477, gamma_curv = 0 */
      qa = (Wrv[n]-pW0[n])*(pW0[n]-Wlv[n]);
      qb = Wrv[n]-Wlv[n];
      qc = 6.0*(pW0[n] - 0.5*(Wlv[n]*(1.0-0.0) + Wrv[n]*(1.0+0.0)));
      if (qa <= 0.0) {
        Wlv[n] = pW0[n];
        Wrv[n] = pW0[n];
      } else if ((qb*qc) > (qb*qb)) {
        Wlv[n] = (6.0*pW0[n] - Wrv[n]*(4.0+3.0*0.0))/(2.0-3.0*0.0);
      } else if ((qb*qc) < -(qb*qb)) {
        Wrv[n] = (6.0*pW0[n] - Wlv[n]*(4.0-3.0*0.0))/(2.0+3.0*0.0);
      }
    }
    for (n=0; n<NV; n++) {                                                                  /* :479-484 */
      Wlv[n] = MAX(MIN(pW0[n],pWm[n]),Wlv[n]);
      Wlv[n] = MIN(MAX(pW0[n],pWm[n]),Wlv[n]);
      Wrv[n] = MAX(MIN(pW0[n],pWp[n]),Wrv[n]);
      Wrv[n] = MIN(MAX(pW0[n],pWp[n]),Wrv[n]);
    }
    for (n=0; n<NV; n++) {                                                                  /* :489-492 */
      dW[n] = Wrv[n] - Wlv[n];
      W6[n] = 6.0*(pW0[n] - 0.5*(Wlv[n]*(1.0-0.0) + Wrv[n]*(1.0+0.0)));
    }
    pWl = (double*)&Wl[i+1];
    pWr = (double*)&Wr[i];

    if (orc_riemann == 1) {                                                                 /* :510-553 HLL-family form */
      for (n=0; n<NWAVE; n++) { pWl[n] = Wrv[n]; pWr[n] = Wlv[n]; }
      for (n=0; n<NWAVE; n++) {
        qa = 0.0;
        if (ev[n] > 0.0) {
          qx1 = 0.5*dtodx*ev[n];
          qb  = qx1;
          qc  = FOUR_3RDS*SQR(qx1);
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*(qb*(dW[m]-W6[m]) + qc*W6[m]);
          for (m=0; m<NWAVE; m++) pWl[m] -= qa*rem[m][n];
        }
        qa = 0.0;
        if (ev[n] < 0.0) {
          qx2 = 0.5*dtodx*ev[n];
          qb  = qx2;
          qc  = FOUR_3RDS*SQR(qx2);
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*(qb*(dW[m]+W6[m]) + qc*W6[m]);
          for (m=0; m<NWAVE; m++) pWr[m] -= qa*rem[m][n];
        }
      }
      Wl[i+1].r = 0.0; Wr[i].r = 0.0;            /* scalar never written by this branch, see lr_states() above */
    } else {
      qx1 = 0.5*MAX(ev[NWAVE-1],0.0)*dtodx;                                                 /* :557-577, qxx = 0 */
      for (n=0; n<NV; n++)
        pWl[n] = Wrv[n] - qx1*(dW[n] - (1.0-FOUR_3RDS*qx1)*W6[n]) + 0.0*(dW[n] - (1.0-2.0*qx1)*W6[n]);
      qx2 = -0.5*MIN(ev[0],0.0)*dtodx;
      for (n=0; n<NV; n++)
        pWr[n] = Wlv[n] + qx2*(dW[n] + (1.0-FOUR_3RDS*qx2)*W6[n]) + 0.0*(dW[n] + (1.0-2.0*qx2)*W6[n]);

      for (n=0; n<NWAVE; n++) {                                                             /* :586-606 */
        if (ev[n] >= 0.0) {
          qa  = 0.0;
          qx1 = 0.5*dtodx*ev[NWAVE-1];
          qx2 = 0.5*dtodx*ev[n];
          qb  = qx1 - qx2;
          qc  = FOUR_3RDS*(SQR(qx1) - SQR(qx2));
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*(qb*(dW[m]-W6[m]) + qc*W6[m]);
          for (m=0; m<NWAVE; m++) pWl[m] += qa*rem[m][n];
        }
      }
      for (n=0; n<NWAVE; n++) {                                                             /* :608-628 */
        if (ev[n] <= 0.0) {
          qa = 0.0;
          qx1 = 0.5*dtodx*ev[0];
          qx2 = 0.5*dtodx*ev[n];
          qb  = qx1 - qx2;
          qc  = FOUR_3RDS*(SQR(qx1) - SQR(qx2));
          for (m=0; m<NWAVE; m++) qa += lem[n][m]*(qb*(dW[m]+W6[m]) + qc*W6[m]);
          for (m=0; m<NWAVE; m++) pWr[m] += qa*rem[m][n];
        }
      }
      for (n=NWAVE; n<NV; n++) {                                                            /* :631-641 (its dW[m], m == NWAVE, is dW[n]) */
        if (W[i].Vx > 0.) {
          qb = 0.5*dtodx*(ev[NWAVE-1]-W[i].Vx);
          qc = 0.5*dtodx*dtodx*TWO_3RDS*(SQR(ev[NWAVE-1]) - SQR(W[i].Vx));
          pWl[n] += qb*(dW[NWAVE]-W6[NWAVE]) + qc*W6[NWAVE];
        } else if (W[i].Vx < 0.) {
          qb = 0.5*dtodx*(ev[0]-W[i].Vx);
          qc = 0.5*dtodx*dtodx*TWO_3RDS*(ev[0]*ev[0] - W[i].Vx*W[i].Vx);
          pWr[n] += qb*(dW[NWAVE]+W6[NWAVE]) + qc*W6[NWAVE];
        }
      }
      if (!nscal) { Wl[i+1].r = 0.0; Wr[i].r = 0.0; }
    }
    for (m=0; m<NWAVE; m++) {                                                               /* :648-654 */
      ev[m] = ev_ip1[m];
      for (n=0; n<NWAVE; n++) { rem[m][n] = rem_ip1[m][n]; lem[m][n] = lem_ip1[m][n]; }
    }
  }
  free(dWm_flat); free(Wim1h_flat);
#undef dWm
#undef Wim1h
}

/* ------------------------------------------------------------------------------------------------
 * hlle.c:62-268 (HYDRO, ADIABATIC)
 * ---------------------------------------------------------------------------------------------- */
static void flux_hlle(const C1 Ul, const C1 Ur, const P1 Wl, const P1 Wr, int nscal,
                      double Gamma, double Gamma_1, C1 *pFlux)
{
  double sqrtdl, sqrtdr, isdlpdr, v1roe, v2roe, v3roe, pbl=0.0, pbr=0.0;
  double asq, vaxsq=0.0, qsq, cfsq, cfl, cfr, bp, bm, ct2=0.0, tmp, hroe;
  double ev[NWAVE], al, ar;
  double *pFl, *pFr, *pF;
  C1 Fl, Fr;
  int n;

  sqrtdl = sqrt(Wl.d);
  sqrtdr = sqrt(Wr.d);
  isdlpdr = 1.0/(sqrtdl + sqrtdr);
  v1roe = (sqrtdl*Wl.Vx + sqrtdr*Wr.Vx)*isdlpdr;
  v2roe = (sqrtdl*Wl.Vy + sqrtdr*Wr.Vy)*isdlpdr;
  v3roe = (sqrtdl*Wl.Vz + sqrtdr*Wr.Vz)*isdlpdr;
  hroe  = ((Ul.E + Wl.P + pbl)/sqrtdl + (Ur.E + Wr.P + pbr)/sqrtdr)*isdlpdr;
  esys_roe_adb_hyd(v1roe, v2roe, v3roe, hroe, Gamma_1, ev, NULL, NULL);

  asq = Gamma*Wl.P/Wl.d;
  qsq = vaxsq + ct2 + asq;
  tmp = vaxsq + ct2 - asq;
  cfsq = 0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2));
  cfl = sqrt(cfsq);

  asq = Gamma*Wr.P/Wr.d;
  qsq = vaxsq + ct2 + asq;
  tmp = vaxsq + ct2 - asq;
  cfsq = 0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2));
  cfr = sqrt(cfsq);

  ar = MAX(ev[NWAVE-1],(Wr.Vx + cfr));
  al = MIN(ev[0]      ,(Wl.Vx - cfl));
  bp = MAX(ar, 0.0);
  bm = MIN(al, 0.0);

  Fl.d  = Ul.Mx - bm*Ul.d;
  Fr.d  = Ur.Mx - bp*Ur.d;
  Fl.Mx = Ul.Mx*(Wl.Vx - bm);
  Fr.Mx = Ur.Mx*(Wr.Vx - bp);
  Fl.My = Ul.My*(Wl.Vx - bm);
  Fr.My = Ur.My*(Wr.Vx - bp);
  Fl.Mz = Ul.Mz*(Wl.Vx - bm);
  Fr.Mz = Ur.Mz*(Wr.Vx - bp);
  Fl.Mx += Wl.P;
  Fr.Mx += Wr.P;
  Fl.E  = Ul.E*(Wl.Vx - bm) + Wl.P*Wl.Vx;
  Fr.E  = Ur.E*(Wr.Vx - bp) + Wr.P*Wr.Vx;
  Fl.s = Fl.d*Wl.r;
  Fr.s = Fr.d*Wr.r;

  pFl = (double*)&Fl; pFr = (double*)&Fr; pF = (double*)pFlux;
  tmp = 0.5*(bp + bm)/(bp - bm);
  for (n=0; n<(NWAVE+nscal); n++) pF[n] = 0.5*(pFl[n] + pFr[n]) + (pFl[n] - pFr[n])*tmp;
  if (!nscal) pFlux->s = 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * hllc.c:50-261 (HYDRO, ADIABATIC, CARTESIAN) -- the reference's --with-flux=hllc build
 * ---------------------------------------------------------------------------------------------- */
static void fluxes_hllc(const C1 Ul, const C1 Ur, const P1 Wl, const P1 Wr, int nscal,
                        double Gamma, double Gamma_1, C1 *pFlux)
{
  double sqrtdl, sqrtdr, isdlpdr, v1roe, v2roe, v3roe, hroe;
  double ev[NWAVE];
  double *pFl, *pFr, *pF;
  C1 Fl, Fr;
  int n;
  double cfl, cfr, bp, bm, tmp;
  double al, ar, am, cp;
  double tl, tr, dl, dr, sl, sm, sr;

  sqrtdl = sqrt(Wl.d);                                                          /* :78-86 */
  sqrtdr = sqrt(Wr.d);
  isdlpdr = 1.0/(sqrtdl + sqrtdr);
  v1roe = (sqrtdl*Wl.Vx + sqrtdr*Wr.Vx)*isdlpdr;
  v2roe = (sqrtdl*Wl.Vy + sqrtdr*Wr.Vy)*isdlpdr;
  v3roe = (sqrtdl*Wl.Vz + sqrtdr*Wr.Vz)*isdlpdr;
  hroe = ((Ul.E + Wl.P)/sqrtdl + (Ur.E + Wr.P)/sqrtdr)*isdlpdr;                 /* :94 */
  esys_roe_adb_hyd(v1roe, v2roe, v3roe, hroe, Gamma_1, ev, NULL, NULL);

  cfl = sqrt(Gamma*Wl.P/Wl.d);                                                  /* :114-115 */
  cfr = sqrt(Gamma*Wr.P/Wr.d);
  ar = MAX(ev[NWAVE-1],(Wr.Vx + cfr));
  al = MIN(ev[0]      ,(Wl.Vx - cfl));
  bp = ar > 0.0 ? ar : 0.0;
  bm = al < 0.0 ? al : 0.0;

  tl = Wl.P + (Wl.Vx - al)*Ul.Mx;                                               /* :132-148 */
  tr = Wr.P + (Wr.Vx - ar)*Ur.Mx;
  dl =   Ul.Mx - Ul.d*al;
  dr = -(Ur.Mx - Ur.d*ar);
  tmp = 1.0/(dl + dr);
  am = (tl - tr)*tmp;
  cp = (dl*tr + dr*tl)*tmp;
  cp = cp > 0.0 ? cp : 0.0;

  Fl.d  = Ul.Mx - bm*Ul.d;                                                      /* :154-176 */
  Fr.d  = Ur.Mx - bp*Ur.d;
  Fl.Mx = Ul.Mx*(Wl.Vx - bm);
  Fr.Mx = Ur.Mx*(Wr.Vx - bp);
  Fl.My = Ul.My*(Wl.Vx - bm);
  Fr.My = Ur.My*(Wr.Vx - bp);
  Fl.Mz = Ul.Mz*(Wl.Vx - bm);
  Fr.Mz = Ur.Mz*(Wr.Vx - bp);
  Fl.Mx += Wl.P;
  Fr.Mx += Wr.P;
  Fl.E  = Ul.E*(Wl.Vx - bm) + Wl.P*Wl.Vx;
  Fr.E  = Ur.E*(Wr.Vx - bp) + Wr.P*Wr.Vx;

  if (am >= 0.0) {                                                              /* :189-198 */
    sl =  am/(am - bm);
    sr = 0.0;
    sm = -bm/(am - bm);
  } else {
    sl =  0.0;
    sr = -am/(bp - am);
    sm =  bp/(bp - am);
  }
  pFl = (double*)&Fl; pFr = (double*)&Fr; pF = (double*)pFlux;
  for (n=0; n<NWAVE; n++) pF[n] = sl*pFl[n] + sr*pFr[n];                        /* :206 */
  pFlux->Mx += sm*cp;
  pFlux->E  += sm*cp*am;
  pFlux->s = 0.0;
  if (nscal) {                                                                  /* :215-221 */
    if (pFlux->d >= 0.0) pFlux->s = pFlux->d*Wl.r;
    else pFlux->s = pFlux->d*Wr.r;
  }
}

void orc_set_riemann_solver(int kind) { orc_riemann = kind; }
void orc_set_reconstruction(int order) { orc_order = order; }

static void fluxes_roe(const C1 Ul, const C1 Ur, const P1 Wl, const P1 Wr, int nscal,
                       double Gamma, double Gamma_1, double etah, C1 *pFlux);
static void fluxes(const C1 Ul, const C1 Ur, const P1 Wl, const P1 Wr, int nscal,
                   double Gamma, double Gamma_1, double etah, C1 *pFlux)
{
  if (orc_riemann == 1) fluxes_hllc(Ul, Ur, Wl, Wr, nscal, Gamma, Gamma_1, pFlux);
  else fluxes_roe(Ul, Ur, Wl, Wr, nscal, Gamma, Gamma_1, (orc_riemann == 2) ? 0.0 : etah, pFlux);   /* 2: built without H_CORRECTION, etah stays 0 (roe.c:33) */
}

/* ------------------------------------------------------------------------------------------------
 * roe.c:59-336  (TEST_INTERMEDIATE_STATES is defined at roe.c:45)
 * ---------------------------------------------------------------------------------------------- */
static void fluxes_roe(const C1 Ul, const C1 Ur, const P1 Wl, const P1 Wr, int nscal,
                   double Gamma, double Gamma_1, double etah, C1 *pFlux)
{
  double sqrtdl, sqrtdr, isdlpdr, v1roe, v2roe, v3roe, pbl=0.0, pbr=0.0, hroe;
  double coeff[NWAVE], ev[NWAVE], rem[NWAVE][NWAVE], lem[NWAVE][NWAVE], dU[NWAVE], a[NWAVE];
  double u_inter[NWAVE], p_inter=0.0;
  const double *pUl, *pUr;
  double *pFl, *pFr, *pF;
  C1 Fl, Fr;
  int n, m, hlle_flag;

  for (n=0; n<NWAVE; n++) for (m=0; m<NWAVE; m++) { rem[n][m] = 0.0; lem[n][m] = 0.0; }

  sqrtdl = sqrt(Wl.d);
  sqrtdr = sqrt(Wr.d);
  isdlpdr = 1.0/(sqrtdl + sqrtdr);
  v1roe = (sqrtdl*Wl.Vx + sqrtdr*Wr.Vx)*isdlpdr;
  v2roe = (sqrtdl*Wl.Vy + sqrtdr*Wr.Vy)*isdlpdr;
  v3roe = (sqrtdl*Wl.Vz + sqrtdr*Wr.Vz)*isdlpdr;
  hroe  = ((Ul.E + Wl.P + pbl)/sqrtdl + (Ur.E + Wr.P + pbr)/sqrtdr)*isdlpdr;

  esys_roe_adb_hyd(v1roe, v2roe, v3roe, hroe, Gamma_1, ev, rem, lem);

  Fl.d  = Ul.Mx;          Fr.d  = Ur.Mx;
  Fl.Mx = Ul.Mx*Wl.Vx;    Fr.Mx = Ur.Mx*Wr.Vx;
  Fl.My = Ul.Mx*Wl.Vy;    Fr.My = Ur.Mx*Wr.Vy;
  Fl.Mz = Ul.Mx*Wl.Vz;    Fr.Mz = Ur.Mx*Wr.Vz;
  Fl.Mx += Wl.P;          Fr.Mx += Wr.P;
  Fl.E  = (Ul.E + Wl.P)*Wl.Vx;
  Fr.E  = (Ur.E + Wr.P)*Wr.Vx;
  Fl.s = nscal ? Fl.d*Wl.r : 0.0;
  Fr.s = nscal ? Fr.d*Wr.r : 0.0;

  if (ev[0] >= 0.0) { *pFlux = Fl; return; }                                    /* :215-224 */
  if (ev[NWAVE-1] <= 0.0) { *pFlux = Fr; return; }                              /* :226-235 */

  pUr = (const double*)&Ur; pUl = (const double*)&Ul;
  for (n=0; n<NWAVE; n++) dU[n] = pUr[n] - pUl[n];
  for (n=0; n<NWAVE; n++) {
    a[n] = 0.0;
    for (m=0; m<NWAVE; m++) a[n] += lem[n][m]*dU[m];
  }

  hlle_flag = 0;                                                                /* :256-284 */
  for (n=0; n<NWAVE; n++) u_inter[n] = pUl[n];
  for (n=0; n<NWAVE-1; n++) {
    for (m=0; m<NWAVE; m++) u_inter[m] += a[n]*rem[m][n];
    if (ev[n+1] > ev[n]) {
      if (u_inter[0] <= 0.0) { hlle_flag=1; break; }
      p_inter = u_inter[4] - 0.5*(SQR(u_inter[1])+SQR(u_inter[2])+SQR(u_inter[3]))/u_inter[0];
      if (p_inter < 0.0) { hlle_flag=2; break; }
    }
  }
  if (hlle_flag != 0) { flux_hlle(Ul,Ur,Wl,Wr,nscal,Gamma,Gamma_1,pFlux); return; }

  pFl = (double*)&Fl; pFr = (double*)&Fr; pF = (double*)pFlux;
  for (m=0; m<NWAVE; m++) coeff[m] = 0.5*MAX(fabs(ev[m]),etah)*a[m];            /* :295-300 */
  for (n=0; n<NWAVE; n++) {
    pF[n] = 0.5*(pFl[n] + pFr[n]);
    for (m=0; m<NWAVE; m++) pF[n] -= coeff[m]*rem[n][m];
  }
  pFlux->s = 0.0;
  if (nscal) {                                                                  /* :316-320 */
    if (pFlux->d >= 0.0) pFlux->s = pFlux->d*Wl.r;
    else pFlux->s = pFlux->d*Wr.r;
  }
}

void orc_roe_flux(const double Ul[6], const double Ur[6], const double Wl[6], const double Wr[6],
                  int nscal, double gamma, double etah, double F[6])
{
  C1 ul, ur, f; P1 wl, wr;
  memcpy(&ul, Ul, sizeof ul); memcpy(&ur, Ur, sizeof ur);
  memcpy(&wl, Wl, sizeof wl); memcpy(&wr, Wr, sizeof wr);
  fluxes_roe(ul, ur, wl, wr, nscal, gamma, gamma - 1.0, etah, &f);
  memcpy(F, &f, sizeof f);
}

void orc_hllc_flux_batch(int n, const double *Ul, const double *Ur, const double *Wl, const double *Wr,
                         int nscal, double gamma, double *F)
{
  int t;
  for (t=0; t<n; t++) {
    C1 ul, ur, f; P1 wl, wr;
    memcpy(&ul, Ul+6*t, sizeof ul); memcpy(&ur, Ur+6*t, sizeof ur);
    memcpy(&wl, Wl+6*t, sizeof wl); memcpy(&wr, Wr+6*t, sizeof wr);
    fluxes_hllc(ul, ur, wl, wr, nscal, gamma, gamma - 1.0, &f);
    memcpy(F+6*t, &f, sizeof f);
  }
}

void orc_roe_flux_batch(int n, const double *Ul, const double *Ur, const double *Wl, const double *Wr,
                        int nscal, double gamma, const double *etah, double *F)
{
  int t;
  for (t=0; t<n; t++) orc_roe_flux(Ul+6*t, Ur+6*t, Wl+6*t, Wr+6*t, nscal, gamma, etah[t], F+6*t);
}
void orc_prim_to_cons_batch(int n, const double *W, int nscal, double gamma, double *U)
{
  int t;
  for (t=0; t<n; t++) orc_prim_to_cons(W+6*t, nscal, gamma, U+6*t);
}
void orc_cons_to_prim_batch(int n, const double *U, int nscal, double gamma, double *W)
{
  int t;
  for (t=0; t<n; t++) orc_cons_to_prim(U+6*t, nscal, gamma, W+6*t);
}

/* ------------------------------------------------------------------------------------------------
 * Grid helpers
 * ---------------------------------------------------------------------------------------------- */
#define N1(g) ((g)->nx1 + 2*NG)
#define N2(g) ((g)->nx2 + 2*NG)
#define N3(g) ((g)->nx3 + 2*NG)
#define IDX(g,k,j,i) (((size_t)(k)*N2(g) + (size_t)(j))*N1(g) + (size_t)(i))

/* cc_pos.c: cc_pos() */
static void cc_pos(const OrcGrid *g, int i, int j, int k, double *x1, double *x2, double *x3)
{
  *x1 = g->x1min + ((double)(i - NG) + 0.5)*g->dx1;
  *x2 = g->x2min + ((double)(j - NG) + 0.5)*g->dx2;
  *x3 = g->x3min + ((double)(k - NG) + 0.5)*g->dx3;
}

/* ------------------------------------------------------------------------------------------------
 * bvals_mhd.c: reflect :959-1313, outflow :1319-1631, periodic :1637-1949; order x1,x2,x3 (:170-172)
 * ---------------------------------------------------------------------------------------------- */
static void copy_cell(OrcGrid *g, size_t dst, size_t src, int negate /* 0 none, 1..3 = M1..M3 */)
{
  g->d[dst] = g->d[src];
  g->M1[dst] = (negate==1) ? -g->M1[src] : g->M1[src];
  g->M2[dst] = (negate==2) ? -g->M2[src] : g->M2[src];
  g->M3[dst] = (negate==3) ? -g->M3[src] : g->M3[src];
  g->E[dst] = g->E[src];
  if (g->nscal) g->s[dst] = g->s[src];
}

void orc_bvals(OrcGrid *g, const int bc[6])
{
  const int is=NG, ie=g->nx1+NG-1, js=NG, je=g->nx2+NG-1, ks=NG, ke=g->nx3+NG-1;
  int i,j,k,side;
  /* x1 faces: active j,k only */
  for (side=0; side<2; side++) {
    int flag = bc[side];
    for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=1; i<=NG; i++) {
      int dst = side ? ie+i : is-i, src;
      if (flag==1 || flag==5) src = side ? ie-(i-1) : is+(i-1);
      else if (flag==2)       src = side ? ie : is;
      else if (flag==4)       src = side ? is+(i-1) : ie-(i-1);
      else continue;
      copy_cell(g, IDX(g,k,j,dst), IDX(g,k,j,src), (flag==1||flag==5) ? 1 : 0);
    }
  }
  /* x2 faces: all i incl. ghosts, active k */
  for (side=0; side<2; side++) {
    int flag = bc[2+side];
    for (k=ks; k<=ke; k++) for (j=1; j<=NG; j++) for (i=is-NG; i<=ie+NG; i++) {
      int dst = side ? je+j : js-j, src;
      if (flag==1 || flag==5) src = side ? je-(j-1) : js+(j-1);
      else if (flag==2)       src = side ? je : js;
      else if (flag==4)       src = side ? js+(j-1) : je-(j-1);
      else continue;
      copy_cell(g, IDX(g,k,dst,i), IDX(g,k,src,i), (flag==1||flag==5) ? 2 : 0);
    }
  }
  /* x3 faces: all i,j incl. ghosts */
  for (side=0; side<2; side++) {
    int flag = bc[4+side];
    for (k=1; k<=NG; k++) for (j=js-NG; j<=je+NG; j++) for (i=is-NG; i<=ie+NG; i++) {
      int dst = side ? ke+k : ks-k, src;
      if (flag==1 || flag==5) src = side ? ke-(k-1) : ks+(k-1);
      else if (flag==2)       src = side ? ke : ks;
      else if (flag==4)       src = side ? ks+(k-1) : ke-(k-1);
      else continue;
      copy_cell(g, IDX(g,dst,j,i), IDX(g,src,j,i), (flag==1||flag==5) ? 3 : 0);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * new_dt.c:69-139 and :155-185
 * ---------------------------------------------------------------------------------------------- */
void orc_max_speeds(const OrcGrid *g, double out[3])
{
  const int is=NG, ie=g->nx1+NG-1, js=NG, je=g->nx2+NG-1, ks=NG, ke=g->nx3+NG-1;
  const double Gamma = g->gamma, Gamma_1 = g->gamma - 1.0;
  double max_v1=0.0, max_v2=0.0, max_v3=0.0;
  int i,j,k;
  for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=is; i<=ie; i++) {
    size_t c = IDX(g,k,j,i);
    double di = 1.0/g->d[c];
    double v1 = g->M1[c]*di, v2 = g->M2[c]*di, v3 = g->M3[c]*di;
    double qsq = v1*v1 + v2*v2 + v3*v3;
    double p = MAX(Gamma_1*(g->E[c] - 0.5*g->d[c]*qsq), TINY_NUMBER);
    double asq = Gamma*p*di;
    if (g->nx1 > 1) max_v1 = MAX(max_v1,fabs(v1)+sqrt(asq));
    if (g->nx2 > 1) max_v2 = MAX(max_v2,fabs(v2)+sqrt(asq));
    if (g->nx3 > 1) max_v3 = MAX(max_v3,fabs(v3)+sqrt(asq));
  }
  out[0] = max_v1; out[1] = max_v2; out[2] = max_v3;
}

double orc_new_dt(const OrcGrid *g, double cour_no, double dt_old, int nstep, double time, double tlim)
{
  double v[3], max_dti = 0.0, dt;
  orc_max_speeds(g, v);
  if (g->nx1 > 1) max_dti = MAX(max_dti, v[0]/g->dx1);
  if (g->nx2 > 1) max_dti = MAX(max_dti, v[1]/g->dx2);
  if (g->nx3 > 1) max_dti = MAX(max_dti, v[2]/g->dx3);
  if (nstep == 0) dt = cour_no/max_dti;
  else dt = MIN(2.0*dt_old, cour_no/max_dti);
  if ((time < tlim) && ((tlim - time) < dt)) dt = tlim - time;
  return dt;
}

/* ------------------------------------------------------------------------------------------------
 * integrate_3d_ctu.c:110-3368
 * ---------------------------------------------------------------------------------------------- */
void orc_integrate_3d_ctu(OrcGrid *g, double dt)
{
  const int n1=N1(g), n2=N2(g), n3=N3(g);
  const int is=NG, ie=g->nx1+NG-1, js=NG, je=g->nx2+NG-1, ks=NG, ke=g->nx3+NG-1;
  const int il=is-2, iu=ie+2, jl=js-2, ju=je+2, kl=ks-2, ku=ke+2;
  const int nscal = g->nscal;
  const double Gamma = g->gamma, Gamma_1 = g->gamma - 1.0;
  const double dtodx1=dt/g->dx1, dtodx2=dt/g->dx2, dtodx3=dt/g->dx3;
  const double q1 = 0.5*dtodx1, q2 = 0.5*dtodx2, q3 = 0.5*dtodx3;
  const double dx1=g->dx1, dx2=g->dx2, dx3=g->dx3;
  const size_t ncell = (size_t)n1*n2*n3;
  int nmax = n1 > n2 ? n1 : n2; if (n3 > nmax) nmax = n3;
  int i,j,k;
  double x1,x2,x3,phicl,phicr,phifc,phil,phir,phic,cfr,cfl,lambdar,lambdal,etah;
  OrcGravPot Phi = g->grav;

  C1 *Ul1 = calloc(ncell,sizeof(C1)), *Ur1 = calloc(ncell,sizeof(C1));
  C1 *Ul2 = calloc(ncell,sizeof(C1)), *Ur2 = calloc(ncell,sizeof(C1));
  C1 *Ul3 = calloc(ncell,sizeof(C1)), *Ur3 = calloc(ncell,sizeof(C1));
  C1 *F1 = calloc(ncell,sizeof(C1)), *F2 = calloc(ncell,sizeof(C1)), *F3 = calloc(ncell,sizeof(C1));
  double *eta1 = calloc(ncell,sizeof(double)), *eta2 = calloc(ncell,sizeof(double));
  double *eta3 = calloc(ncell,sizeof(double)), *dhalf = calloc(ncell,sizeof(double));
  C1 *U1d = malloc(nmax*sizeof(C1));
  P1 *W = malloc(nmax*sizeof(P1)), *Wl = malloc(nmax*sizeof(P1)), *Wr = malloc(nmax*sizeof(P1));
#define X(k,j,i) IDX(g,k,j,i)
#define SV(c) (nscal ? g->s[c] : 0.0)

  /*=== STEP 1: x1 L/R states and fluxes (:202-526) ===*/
  for (k=kl; k<=ku; k++) for (j=jl; j<=ju; j++) {
    for (i=is-NG; i<=ie+NG; i++) {
      size_t c = X(k,j,i);
      U1d[i].d = g->d[c]; U1d[i].Mx = g->M1[c]; U1d[i].My = g->M2[c]; U1d[i].Mz = g->M3[c];
      U1d[i].E = g->E[c]; U1d[i].s = SV(c);
    }
    for (i=is-NG; i<=ie+NG; i++) W[i] = cons1d_to_prim1d(&U1d[i],nscal,Gamma_1);
    lr_states(W,nscal,Gamma,dt,dx1,il+1,iu-1,Wl,Wr);
    if (Phi != NULL) {
      for (i=il+1; i<=iu; i++) {
        cc_pos(g,i,j,k,&x1,&x2,&x3);
        phicr = Phi( x1         ,x2,x3);
        phicl = Phi((x1-    dx1),x2,x3);
        phifc = Phi((x1-0.5*dx1),x2,x3);
        Wl[i].Vx -= dtodx1*(phifc - phicl);
        Wr[i].Vx -= dtodx1*(phicr - phifc);
      }
    }
    for (i=il+1; i<=iu; i++) {
      size_t c = X(k,j,i);
      Ul1[c] = prim1d_to_cons1d(&Wl[i],nscal,Gamma_1);
      Ur1[c] = prim1d_to_cons1d(&Wr[i],nscal,Gamma_1);
      fluxes(Ul1[c],Ur1[c],Wl[i],Wr[i],nscal,Gamma,Gamma_1,0.0,&F1[c]);
    }
  }

  /*=== STEP 2: x2 (:535-714), (Mx,My,Mz) = (M2,M3,M1) ===*/
  for (k=kl; k<=ku; k++) for (i=il; i<=iu; i++) {
    for (j=js-NG; j<=je+NG; j++) {
      size_t c = X(k,j,i);
      U1d[j].d = g->d[c]; U1d[j].Mx = g->M2[c]; U1d[j].My = g->M3[c]; U1d[j].Mz = g->M1[c];
      U1d[j].E = g->E[c]; U1d[j].s = SV(c);
    }
    for (j=js-NG; j<=je+NG; j++) W[j] = cons1d_to_prim1d(&U1d[j],nscal,Gamma_1);
    lr_states(W,nscal,Gamma,dt,dx2,jl+1,ju-1,Wl,Wr);
    if (Phi != NULL) {
      for (j=jl+1; j<=ju; j++) {
        cc_pos(g,i,j,k,&x1,&x2,&x3);
        phicr = Phi(x1, x2         ,x3);
        phicl = Phi(x1,(x2-    dx2),x3);
        phifc = Phi(x1,(x2-0.5*dx2),x3);
        Wl[j].Vx -= dtodx2*(phifc - phicl);
        Wr[j].Vx -= dtodx2*(phicr - phifc);
      }
    }
    for (j=jl+1; j<=ju; j++) {
      size_t c = X(k,j,i);
      Ul2[c] = prim1d_to_cons1d(&Wl[j],nscal,Gamma_1);
      Ur2[c] = prim1d_to_cons1d(&Wr[j],nscal,Gamma_1);
      fluxes(Ul2[c],Ur2[c],Wl[j],Wr[j],nscal,Gamma,Gamma_1,0.0,&F2[c]);
    }
  }

  /*=== STEP 3: x3 (:723-897), (Mx,My,Mz) = (M3,M1,M2) ===*/
  for (j=jl; j<=ju; j++) for (i=il; i<=iu; i++) {
    for (k=ks-NG; k<=ke+NG; k++) {
      size_t c = X(k,j,i);
      U1d[k].d = g->d[c]; U1d[k].Mx = g->M3[c]; U1d[k].My = g->M1[c]; U1d[k].Mz = g->M2[c];
      U1d[k].E = g->E[c]; U1d[k].s = SV(c);
    }
    for (k=ks-NG; k<=ke+NG; k++) W[k] = cons1d_to_prim1d(&U1d[k],nscal,Gamma_1);
    lr_states(W,nscal,Gamma,dt,dx3,kl+1,ku-1,Wl,Wr);
    if (Phi != NULL) {
      for (k=kl+1; k<=ku; k++) {
        cc_pos(g,i,j,k,&x1,&x2,&x3);
        phicr = Phi(x1,x2, x3         );
        phicl = Phi(x1,x2,(x3-    dx3));
        phifc = Phi(x1,x2,(x3-0.5*dx3));
        Wl[k].Vx -= dtodx3*(phifc - phicl);
        Wr[k].Vx -= dtodx3*(phicr - phifc);
      }
    }
    for (k=kl+1; k<=ku; k++) {
      size_t c = X(k,j,i);
      Ul3[c] = prim1d_to_cons1d(&Wl[k],nscal,Gamma_1);
      Ur3[c] = prim1d_to_cons1d(&Wr[k],nscal,Gamma_1);
      fluxes(Ul3[c],Ur3[c],Wl[k],Wr[k],nscal,Gamma,Gamma_1,0.0,&F3[c]);
    }
  }

/* U.(X,Y,Z) -= q*(Fp.(a,b,c) - Fm.(a,b,c)) with the component permutations of SURVEY App. A.2 */
#define CORR(U,q,Fp,Fm,a,b,c) do { \
    (U).d  -= (q)*((Fp).d  - (Fm).d ); \
    (U).Mx -= (q)*((Fp).a  - (Fm).a ); \
    (U).My -= (q)*((Fp).b  - (Fm).b ); \
    (U).Mz -= (q)*((Fp).c  - (Fm).c ); \
    (U).E  -= (q)*((Fp).E  - (Fm).E ); \
    if (nscal) (U).s -= (q)*((Fp).s - (Fm).s); } while (0)

  /*=== STEP 5: correct x1 faces (:978-1065), gravity (:1167-1219) ===*/
  for (k=kl+1; k<=ku-1; k++) for (j=jl+1; j<=ju-1; j++) for (i=il+1; i<=iu; i++) {
    size_t c = X(k,j,i);
    CORR(Ul1[c],q2,F2[X(k,j+1,i-1)],F2[X(k,j,i-1)],Mz,Mx,My);
    CORR(Ur1[c],q2,F2[X(k,j+1,i  )],F2[X(k,j,i  )],Mz,Mx,My);
    CORR(Ul1[c],q3,F3[X(k+1,j,i-1)],F3[X(k,j,i-1)],My,Mz,Mx);
    CORR(Ur1[c],q3,F3[X(k+1,j,i  )],F3[X(k,j,i  )],My,Mz,Mx);
  }
  if (Phi != NULL) {
    for (k=kl+1; k<=ku-1; k++) for (j=jl+1; j<=ju-1; j++) for (i=il+1; i<=iu; i++) {
      size_t c = X(k,j,i);
      cc_pos(g,i,j,k,&x1,&x2,&x3);
      phic = Phi(x1, x2         ,x3);
      phir = Phi(x1,(x2+0.5*dx2),x3);
      phil = Phi(x1,(x2-0.5*dx2),x3);
      Ur1[c].My -= q2*(phir-phil)*g->d[c];
      Ur1[c].E -= q2*(F2[X(k,j  ,i)].d*(phic - phil) + F2[X(k,j+1,i)].d*(phir - phic));
      phir = Phi(x1,x2,(x3+0.5*dx3));
      phil = Phi(x1,x2,(x3-0.5*dx3));
      Ur1[c].Mz -= q3*(phir-phil)*g->d[c];
      Ur1[c].E -= q3*(F3[X(k  ,j,i)].d*(phic - phil) + F3[X(k+1,j,i)].d*(phir - phic));

      phic = Phi((x1-dx1), x2         ,x3);
      phir = Phi((x1-dx1),(x2+0.5*dx2),x3);
      phil = Phi((x1-dx1),(x2-0.5*dx2),x3);
      Ul1[c].My -= q2*(phir-phil)*g->d[X(k,j,i-1)];
      Ul1[c].E -= q2*(F2[X(k,j  ,i-1)].d*(phic - phil) + F2[X(k,j+1,i-1)].d*(phir - phic));
      phir = Phi((x1-dx1),x2,(x3+0.5*dx3));
      phil = Phi((x1-dx1),x2,(x3-0.5*dx3));
      Ul1[c].Mz -= q3*(phir-phil)*g->d[X(k,j,i-1)];
      Ul1[c].E -= q3*(F3[X(k  ,j,i-1)].d*(phic - phil) + F3[X(k+1,j,i-1)].d*(phir - phic));
    }
  }

  /*=== STEP 6: correct x2 faces (:1282-1366), gravity (:1463-1526) ===*/
  for (k=kl+1; k<=ku-1; k++) for (j=jl+1; j<=ju; j++) for (i=il+1; i<=iu-1; i++) {
    size_t c = X(k,j,i);
    CORR(Ul2[c],q1,F1[X(k,j-1,i+1)],F1[X(k,j-1,i)],My,Mz,Mx);
    CORR(Ur2[c],q1,F1[X(k,j  ,i+1)],F1[X(k,j  ,i)],My,Mz,Mx);
    CORR(Ul2[c],q3,F3[X(k+1,j-1,i)],F3[X(k,j-1,i)],Mz,Mx,My);
    CORR(Ur2[c],q3,F3[X(k+1,j  ,i)],F3[X(k,j  ,i)],Mz,Mx,My);
  }
  if (Phi != NULL) {
    for (k=kl+1; k<=ku-1; k++) for (j=jl+1; j<=ju; j++) for (i=il+1; i<=iu-1; i++) {
      size_t c = X(k,j,i);
      cc_pos(g,i,j,k,&x1,&x2,&x3);
      phic = Phi((x1        ),x2,x3);
      phir = Phi((x1+0.5*dx1),x2,x3);
      phil = Phi((x1-0.5*dx1),x2,x3);
      Ur2[c].Mz -= q1*(phir-phil)*g->d[c];
      Ur2[c].E -= q1*(F1[X(k,j,i  )].d*(phic - phil) + F1[X(k,j,i+1)].d*(phir - phic));
      phir = Phi(x1,x2,(x3+0.5*dx3));
      phil = Phi(x1,x2,(x3-0.5*dx3));
      Ur2[c].My -= q3*(phir-phil)*g->d[c];
      Ur2[c].E -= q3*(F3[X(k  ,j,i)].d*(phic - phil) + F3[X(k+1,j,i)].d*(phir - phic));

      phic = Phi((x1        ),(x2-dx2),x3);
      phir = Phi((x1+0.5*dx1),(x2-dx2),x3);
      phil = Phi((x1-0.5*dx1),(x2-dx2),x3);
      Ul2[c].Mz -= q1*(phir-phil)*g->d[X(k,j-1,i)];
      Ul2[c].E -= q1*(F1[X(k,j-1,i  )].d*(phic - phil) + F1[X(k,j-1,i+1)].d*(phir - phic));
      phir = Phi(x1,(x2-dx2),(x3+0.5*dx3));
      phil = Phi(x1,(x2-dx2),(x3-0.5*dx3));
      Ul2[c].My -= q3*(phir-phil)*g->d[X(k,j-1,i)];
      Ul2[c].E -= q3*(F3[X(k  ,j-1,i)].d*(phic - phil) + F3[X(k+1,j-1,i)].d*(phir - phic));
    }
  }

  /*=== STEP 7: correct x3 faces (:1691-1776), gravity (:1873-1937) ===*/
  for (k=kl+1; k<=ku; k++) for (j=jl+1; j<=ju-1; j++) for (i=il+1; i<=iu-1; i++) {
    size_t c = X(k,j,i);
    CORR(Ul3[c],q1,F1[X(k-1,j,i+1)],F1[X(k-1,j,i)],Mz,Mx,My);
    CORR(Ur3[c],q1,F1[X(k  ,j,i+1)],F1[X(k  ,j,i)],Mz,Mx,My);
    CORR(Ul3[c],q2,F2[X(k-1,j+1,i)],F2[X(k-1,j,i)],My,Mz,Mx);
    CORR(Ur3[c],q2,F2[X(k  ,j+1,i)],F2[X(k  ,j,i)],My,Mz,Mx);
  }
  if (Phi != NULL) {
    for (k=kl+1; k<=ku; k++) for (j=jl+1; j<=ju-1; j++) for (i=il+1; i<=iu-1; i++) {
      size_t c = X(k,j,i);
      cc_pos(g,i,j,k,&x1,&x2,&x3);
      phic = Phi((x1        ),x2,x3);
      phir = Phi((x1+0.5*dx1),x2,x3);
      phil = Phi((x1-0.5*dx1),x2,x3);
      Ur3[c].My -= q1*(phir-phil)*g->d[c];
      Ur3[c].E -= q1*(F1[X(k,j,i  )].d*(phic - phil) + F1[X(k,j,i+1)].d*(phir - phic));
      phir = Phi(x1,(x2+0.5*dx2),x3);
      phil = Phi(x1,(x2-0.5*dx2),x3);
      Ur3[c].Mz -= q2*(phir-phil)*g->d[c];
      Ur3[c].E -= q2*(F2[X(k,j  ,i)].d*(phic - phil) + F2[X(k,j+1,i)].d*(phir - phic));

      phic = Phi((x1        ),x2,(x3-dx3));
      phir = Phi((x1+0.5*dx1),x2,(x3-dx3));
      phil = Phi((x1-0.5*dx1),x2,(x3-dx3));
      Ul3[c].My -= q1*(phir-phil)*g->d[X(k-1,j,i)];
      Ul3[c].E -= q1*(F1[X(k-1,j,i  )].d*(phic - phil) + F1[X(k-1,j,i+1)].d*(phir - phic));
      phir = Phi(x1,(x2+0.5*dx2),(x3-dx3));
      phil = Phi(x1,(x2-0.5*dx2),(x3-dx3));
      Ul3[c].Mz -= q2*(phir-phil)*g->d[X(k-1,j,i)];
      Ul3[c].E -= q2*(F2[X(k-1,j  ,i)].d*(phic - phil) + F2[X(k-1,j+1,i)].d*(phir - phic));
    }
  }

  /*=== STEP 8a: d^{n+1/2} (:2104-2125) ===*/
  if (Phi != NULL) {
    for (k=kl+1; k<=ku-1; k++) for (j=jl+1; j<=ju-1; j++) for (i=il+1; i<=iu-1; i++) {
      size_t c = X(k,j,i);
      dhalf[c] = g->d[c]
        - q1*(F1[X(k,j,i+1)].d - F1[c].d)
        - q2*(F2[X(k,j+1,i)].d - F2[c].d)
        - q3*(F3[X(k+1,j,i)].d - F3[c].d);
    }
  }

  /*=== STEP 9a: eta (:2300-2343) ===*/
  for (k=ks-1; k<=ke+1; k++) for (j=js-1; j<=je+1; j++) for (i=is-1; i<=ie+2; i++) {
    size_t c = X(k,j,i);
    cfr = cfast(&Ur1[c],Gamma,Gamma_1);
    cfl = cfast(&Ul1[c],Gamma,Gamma_1);
    lambdar = Ur1[c].Mx/Ur1[c].d + cfr;
    lambdal = Ul1[c].Mx/Ul1[c].d - cfl;
    eta1[c] = 0.5*fabs(lambdar - lambdal);
  }
  for (k=ks-1; k<=ke+1; k++) for (j=js-1; j<=je+2; j++) for (i=is-1; i<=ie+1; i++) {
    size_t c = X(k,j,i);
    cfr = cfast(&Ur2[c],Gamma,Gamma_1);
    cfl = cfast(&Ul2[c],Gamma,Gamma_1);
    lambdar = Ur2[c].Mx/Ur2[c].d + cfr;
    lambdal = Ul2[c].Mx/Ul2[c].d - cfl;
    eta2[c] = 0.5*fabs(lambdar - lambdal);
  }
  for (k=ks-1; k<=ke+2; k++) for (j=js-1; j<=je+1; j++) for (i=is-1; i<=ie+1; i++) {
    size_t c = X(k,j,i);
    cfr = cfast(&Ur3[c],Gamma,Gamma_1);
    cfl = cfast(&Ul3[c],Gamma,Gamma_1);
    lambdar = Ur3[c].Mx/Ur3[c].d + cfr;
    lambdal = Ul3[c].Mx/Ul3[c].d - cfl;
    eta3[c] = 0.5*fabs(lambdar - lambdal);
  }

  /*=== STEP 9b-d: final fluxes (:2350-2437) ===*/
  for (k=ks-1; k<=ke+1; k++) for (j=js-1; j<=je+1; j++) for (i=is; i<=ie+1; i++) {
    size_t c = X(k,j,i);
    P1 wl, wr;
    etah = MAX(eta2[X(k,j,i-1)],eta2[X(k,j,i)]);
    etah = MAX(etah,eta2[X(k,j+1,i-1)]);
    etah = MAX(etah,eta2[X(k,j+1,i  )]);
    etah = MAX(etah,eta3[X(k  ,j,i-1)]);
    etah = MAX(etah,eta3[X(k  ,j,i  )]);
    etah = MAX(etah,eta3[X(k+1,j,i-1)]);
    etah = MAX(etah,eta3[X(k+1,j,i  )]);
    etah = MAX(etah,eta1[X(k  ,j,i  )]);
    wl = cons1d_to_prim1d(&Ul1[c],nscal,Gamma_1);
    wr = cons1d_to_prim1d(&Ur1[c],nscal,Gamma_1);
    fluxes(Ul1[c],Ur1[c],wl,wr,nscal,Gamma,Gamma_1,etah,&F1[c]);
  }
  for (k=ks-1; k<=ke+1; k++) for (j=js; j<=je+1; j++) for (i=is-1; i<=ie+1; i++) {
    size_t c = X(k,j,i);
    P1 wl, wr;
    etah = MAX(eta1[X(k,j-1,i)],eta1[X(k,j,i)]);
    etah = MAX(etah,eta1[X(k,j-1,i+1)]);
    etah = MAX(etah,eta1[X(k,j  ,i+1)]);
    etah = MAX(etah,eta3[X(k  ,j-1,i)]);
    etah = MAX(etah,eta3[X(k  ,j  ,i)]);
    etah = MAX(etah,eta3[X(k+1,j-1,i)]);
    etah = MAX(etah,eta3[X(k+1,j  ,i)]);
    etah = MAX(etah,eta2[X(k  ,j  ,i)]);
    wl = cons1d_to_prim1d(&Ul2[c],nscal,Gamma_1);
    wr = cons1d_to_prim1d(&Ur2[c],nscal,Gamma_1);
    fluxes(Ul2[c],Ur2[c],wl,wr,nscal,Gamma,Gamma_1,etah,&F2[c]);
  }
  for (k=ks; k<=ke+1; k++) for (j=js-1; j<=je+1; j++) for (i=is-1; i<=ie+1; i++) {
    size_t c = X(k,j,i);
    P1 wl, wr;
    etah = MAX(eta1[X(k-1,j,i)],eta1[X(k,j,i)]);
    etah = MAX(etah,eta1[X(k-1,j,i+1)]);
    etah = MAX(etah,eta1[X(k  ,j,i+1)]);
    etah = MAX(etah,eta2[X(k-1,j  ,i)]);
    etah = MAX(etah,eta2[X(k  ,j  ,i)]);
    etah = MAX(etah,eta2[X(k-1,j+1,i)]);
    etah = MAX(etah,eta2[X(k  ,j+1,i)]);
    etah = MAX(etah,eta3[X(k  ,j  ,i)]);
    wl = cons1d_to_prim1d(&Ul3[c],nscal,Gamma_1);
    wr = cons1d_to_prim1d(&Ur3[c],nscal,Gamma_1);
    fluxes(Ul3[c],Ur3[c],wl,wr,nscal,Gamma,Gamma_1,etah,&F3[c]);
  }

  /*=== STEP 12e (SMR): keep the final fluxes, physical component order (:3076-3364) ===*/
  {
    int dir;
    for (dir=0; dir<3; dir++) if (g->Fsave[dir] != NULL) {
      const C1 *F = (dir == 0) ? F1 : (dir == 1) ? F2 : F3;
      size_t c;
      for (c=0; c<ncell; c++) {
        double *o = g->Fsave[dir] + 6*c;
        o[0] = F[c].d; o[4] = F[c].E; o[5] = F[c].s;
        if (dir == 0) { o[1] = F[c].Mx; o[2] = F[c].My; o[3] = F[c].Mz; }
        if (dir == 1) { o[1] = F[c].Mz; o[2] = F[c].Mx; o[3] = F[c].My; }
        if (dir == 2) { o[1] = F[c].My; o[2] = F[c].Mz; o[3] = F[c].Mx; }
      }
    }
  }

  /*=== STEP 11a: static gravity, full step (:2741-2782) ===*/
  if (Phi != NULL) {
    for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=is; i<=ie; i++) {
      size_t c = X(k,j,i);
      cc_pos(g,i,j,k,&x1,&x2,&x3);
      phic = Phi((x1        ),x2,x3);
      phir = Phi((x1+0.5*dx1),x2,x3);
      phil = Phi((x1-0.5*dx1),x2,x3);
      g->M1[c] -= dtodx1*(phir-phil)*dhalf[c];
      g->E[c] -= dtodx1*(F1[c].d*(phic - phil) + F1[X(k,j,i+1)].d*(phir - phic));
      phir = Phi(x1,(x2+0.5*dx2),x3);
      phil = Phi(x1,(x2-0.5*dx2),x3);
      g->M2[c] -= dtodx2*(phir-phil)*dhalf[c];
      g->E[c] -= dtodx2*(F2[c].d*(phic - phil) + F2[X(k,j+1,i)].d*(phir - phic));
      phir = Phi(x1,x2,(x3+0.5*dx3));
      phil = Phi(x1,x2,(x3-0.5*dx3));
      g->M3[c] -= dtodx3*(phir-phil)*dhalf[c];
      g->E[c] -= dtodx3*(F3[c].d*(phic - phil) + F3[X(k+1,j,i)].d*(phir - phic));
    }
  }

  /*=== STEP 12: conservative update, three passes (:2981-3050) ===*/
  for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=is; i<=ie; i++) {
    size_t c = X(k,j,i), p = X(k,j,i+1);
    g->d[c]  -= dtodx1*(F1[p].d  - F1[c].d );
    g->M1[c] -= dtodx1*(F1[p].Mx - F1[c].Mx);
    g->M2[c] -= dtodx1*(F1[p].My - F1[c].My);
    g->M3[c] -= dtodx1*(F1[p].Mz - F1[c].Mz);
    g->E[c]  -= dtodx1*(F1[p].E  - F1[c].E );
    if (nscal) g->s[c] -= dtodx1*(F1[p].s - F1[c].s);
  }
  for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=is; i<=ie; i++) {
    size_t c = X(k,j,i), p = X(k,j+1,i);
    g->d[c]  -= dtodx2*(F2[p].d  - F2[c].d );
    g->M1[c] -= dtodx2*(F2[p].Mz - F2[c].Mz);
    g->M2[c] -= dtodx2*(F2[p].Mx - F2[c].Mx);
    g->M3[c] -= dtodx2*(F2[p].My - F2[c].My);
    g->E[c]  -= dtodx2*(F2[p].E  - F2[c].E );
    if (nscal) g->s[c] -= dtodx2*(F2[p].s - F2[c].s);
  }
  for (k=ks; k<=ke; k++) for (j=js; j<=je; j++) for (i=is; i<=ie; i++) {
    size_t c = X(k,j,i), p = X(k+1,j,i);
    g->d[c]  -= dtodx3*(F3[p].d  - F3[c].d );
    g->M1[c] -= dtodx3*(F3[p].My - F3[c].My);
    g->M2[c] -= dtodx3*(F3[p].Mz - F3[c].Mz);
    g->M3[c] -= dtodx3*(F3[p].Mx - F3[c].Mx);
    g->E[c]  -= dtodx3*(F3[p].E  - F3[c].E );
    if (nscal) g->s[c] -= dtodx3*(F3[p].s - F3[c].s);
  }

#undef X
#undef SV
#undef CORR
  free(Ul1); free(Ur1); free(Ul2); free(Ur2); free(Ul3); free(Ur3);
  free(F1); free(F2); free(F3); free(eta1); free(eta2); free(eta3); free(dhalf);
  free(U1d); free(W); free(Wl); free(Wr);
}

/* ================================================================================================
 * Ion radiative transfer, ionradiation/
 * ============================================================================================== */
#define MINFLUXFRAC 1.0e-3     /* ionrad.h:26 */
#define MINOPTDEPTH 1.0e-4     /* ionrad.h:29 */
#define IONFRACFLOOR 1.0e-4    /* ionrad.h:31 */
#define CION 8.0e5             /* ionrad.h:36 */
#define MAXCELLCOUNT 20        /* ionrad.h:38 */
#define LARGE DBL_MAX          /* ionrad.h:16 */
#define KB 1.38e-16            /* ionrad_chemistry.c:43 */
#define MAXSIGNCOUNT 4         /* ionrad_3d.c:286 */
#define DAMPFACTOR 0.5         /* ionrad_3d.c:287 */

/* ionrad_chemistry.c:111-126, :137-150, :350-353 */
static double recomb_rate_coef(double T) { return 2.59e-13*pow(T/1.0e4, -0.7); }
static double recomb_cool_rate_coef(double T) { if (T < 100.0) return 0.0; return 6.11e-10*pow(T,-0.89)*KB*T; }
static double lya_cool_rate(double nh, double nhplus, double T) { return -7.5e-19*nhplus*nh*exp(-118348/T); }

void orc_ion_derived(OrcIonPar *p, double dx1, double dx2, double dx3)   /* ionrad.c:115-131 */
{
  double area1 = dx1*dx2, area2 = dx1*dx3, area3 = dx2*dx3, maxdx;
  if (area1 < area2) { if (area1 < area3) p->min_area = area1; else p->min_area = area3; }
  else { if (area2 < area3) p->min_area = area2; else p->min_area = area3; }
  maxdx = dx1 > dx2 ? dx1 : dx2;
  maxdx = maxdx > dx3 ? maxdx : dx2;          /* sic: ionrad.c:129 */
  p->d_nlo = MINOPTDEPTH * p->m_H / (p->sigma_ph * maxdx);
}

double orc_ion_entering_flux(double flux_i, double time)                  /* ionradplane_3d.c:265 */
{
  return flux_i*(5.*(erf((time - 1.2e5)/8e4)+1)+0.1);
}

void orc_bvals_ionrad(const OrcGrid *g, double flux_i, double *edgeflux)   /* bvals_ionrad.c:308-332 */
{
  int j,k;
  for (k=0; k<=g->nx3; k++) for (j=0; j<=g->nx2; j++)
    edgeflux[((size_t)k*(g->nx2+1) + (size_t)j)*(g->nx1+1)] = flux_i;
}

typedef struct {
  OrcGrid *g; const OrcIonPar *p;
  double *ph_rate, *edot, *nHdot, *e_init, *e_th_init, *x_init;
  int *last_sign, *sign_count;
  int is,ie,js,je,ks,ke;
  double Gamma_1;
} IonCtx;

#define IONLOOP(c) for (k=c->ks; k<=c->ke; k++) for (j=c->js; j<=c->je; j++) for (i=c->is; i<=c->ie; i++)
#define KE(g,x) (0.5 * ((g)->M1[x]*(g)->M1[x] + (g)->M2[x]*(g)->M2[x] + (g)->M3[x]*(g)->M3[x]) / (g)->d[x])

static void apply_temp_floor(IonCtx *c)                                    /* ionrad_3d.c:70-136 */
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p; int i,j,k;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double n_H = g->s[x] / p->m_H;
    double n_Hplus = (g->d[x] - g->s[x]) / p->m_H;
    double n_e = n_Hplus + g->d[x] * p->alpha_C / (14.0 * p->m_H);
    double xi = n_e / (n_H + n_Hplus);
    double ke = KE(g,x);
    double e_thermal = g->E[x] - ke;
    double e_sp = e_thermal / g->d[x];
    double T = c->Gamma_1 * e_sp * (xi*0.5*p->m_H+(1.0-xi)*p->mu)/ p->k_B;
    if (T < p->tfloor) {
      e_sp = p->tfloor * p->k_B / ((xi*0.5*p->m_H+(1.0-xi)*p->mu) * c->Gamma_1);
      g->E[x] = KE(g,x) + e_sp * g->d[x];
    }
    if ((T > p->tceil) && (p->tceil > 0)) {
      e_sp = p->tceil * p->k_B / ((xi*0.5*p->m_H+(1.0-xi)*p->mu) * c->Gamma_1);
      g->E[x] = KE(g,x) + e_sp * g->d[x];
    }
  }
}

static void apply_neutral_floor(IonCtx *c)                                 /* :140-157 */
{
  OrcGrid *g = c->g; int i,j,k;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double d_nlim = g->d[x]*IONFRACFLOOR;
    d_nlim = d_nlim < c->p->d_nlo ? d_nlim : c->p->d_nlo;
    if (g->s[x] < d_nlim) g->s[x] = d_nlim;
    else if (g->s[x] > g->d[x]) g->s[x] = g->d[x];
  }
}

static void save_energy_and_x(IonCtx *c)                                   /* :162-201 */
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p; int i,j,k;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double e_thermal = g->E[x] - KE(g,x);
    double n_H = g->s[x] / p->m_H;
    double n_Hplus = (g->d[x] - g->s[x]) / p->m_H;
    double n_e = n_Hplus + g->d[x] * p->alpha_C / (14.0 * p->m_H);
    c->e_init[x] = g->E[x];
    c->e_th_init[x] = e_thermal;
    c->x_init[x] = n_e / (n_H + n_Hplus);
    c->last_sign[x] = 0;
    c->sign_count[x] = 0;
  }
}

static int check_range(IonCtx *c)                                          /* :206-283 */
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p; int i,j,k;
  long cellcount = 0;
  double e_thermal = 0.0;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double n_H = g->s[x] / p->m_H, n_Hplus, n_e, xi;
    if (c->ph_rate[x] / (p->min_area * n_H) > 2.0*CION) continue;
    if (p->max_de_therm_step > 0) e_thermal = g->E[x] - KE(g,x);
    if ((e_thermal / c->e_th_init[x] >= 1 + p->max_de_therm_step) ||
        (c->e_th_init[x] / e_thermal >= 1 + p->max_de_therm_step)) { cellcount++; continue; }
    if (p->max_de_step > 0) {
      if ((g->E[x] / c->e_init[x] >= 1 + p->max_de_step) ||
          (c->e_init[x] / g->E[x] >= 1 + p->max_de_step)) { cellcount++; continue; }
    }
    if (p->max_dx_step > 0) {
      n_Hplus = (g->d[x] - g->s[x]) / p->m_H;
      n_e = n_Hplus + g->d[x] * p->alpha_C / (14.0 * p->m_H);
      xi = n_e / (n_H + n_Hplus);
      if ((xi / c->x_init[x] >= 1 + p->max_dx_step) ||
          (c->x_init[x] / xi >= 1 + p->max_dx_step)) { cellcount++; continue; }
    }
  }
  return cellcount > MAXCELLCOUNT;
}

static double compute_chem_rates(IonCtx *c)                                /* :288-409 */
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p; int i,j,k,n;
  double dt_chem_min = LARGE;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double n_H = g->s[x] / p->m_H;
    double n_Hplus = (g->d[x] - g->s[x]) / p->m_H;
    double n_e = n_Hplus + g->d[x] * p->alpha_C / (14.0 * p->m_H);
    double xi = n_e / (n_H + n_Hplus);
    double e_sp = (g->E[x] - KE(g,x)) / g->d[x];
    double T = c->Gamma_1 * e_sp * (xi*0.5*p->m_H+(1.0-xi)*p->mu)/ p->k_B;
    double d_nlim, dt_chem1, dt_chem2, dt_chem;
    if (T < p->tfloor) T = p->tfloor;
    c->nHdot[x] = recomb_rate_coef(T) * p->time_unit * n_e * n_Hplus - c->ph_rate[x] * n_H;
    if (c->nHdot[x] < 0.0) {
      if (c->last_sign[x] == 1) c->sign_count[x]++;
      else if (c->sign_count[x] > 0) c->sign_count[x]--;
      c->last_sign[x] = -1;
    } else if (c->nHdot[x] > 0.0) {
      if (c->last_sign[x] == -1) c->sign_count[x]++;
      else if (c->sign_count[x] > 0) c->sign_count[x]--;
      c->last_sign[x] = 1;
    } else {
      c->sign_count[x] = c->last_sign[x] = 0;
    }
    for (n=MAXSIGNCOUNT; n<c->sign_count[x]; n++) {
      c->edot[x] *= DAMPFACTOR;
      c->nHdot[x] *= DAMPFACTOR;
    }
    d_nlim = g->d[x]*IONFRACFLOOR;
    d_nlim = d_nlim < p->d_nlo ? d_nlim : p->d_nlo;
    if (c->nHdot[x] == 0.0) {
      dt_chem1 = dt_chem2 = LARGE;
    } else if (c->nHdot[x] > 0.0) {
      dt_chem1 = p->max_dx_iter / (1+p->max_dx_iter) * n_e / c->nHdot[x];
      dt_chem2 = p->max_dx_iter * n_H / c->nHdot[x];
    } else if (g->s[x] > 1.0001*d_nlim) {
      dt_chem1 = -p->max_dx_iter * n_e / c->nHdot[x];
      dt_chem2 = -p->max_dx_iter / (1+p->max_dx_iter) * n_H / c->nHdot[x];
    } else {
      dt_chem1 = dt_chem2 = LARGE;
    }
    dt_chem = (dt_chem1 < dt_chem2) ? dt_chem1 : dt_chem2;
    if (dt_chem < dt_chem_min) dt_chem_min = dt_chem;
    if (dt_chem < 0) { fprintf(stderr,"[orc compute_chem_rates]: dt_chem = %e < 0\n", dt_chem); abort(); }
  }
  return dt_chem_min;
}

static double compute_therm_rates(IonCtx *c)                               /* :414-562 */
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p; int i,j,k;
  double dt_therm_min = LARGE;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double n_H = g->s[x] / p->m_H;
    double n_Hplus = (g->d[x] - g->s[x]) / p->m_H;
    double n_e = n_Hplus + g->d[x] * p->alpha_C / (14.0 * p->m_H);
    double xi = n_e / (n_H + n_Hplus);
    double e_thermal = g->E[x] - KE(g,x);
    double e_sp = e_thermal / g->d[x];
    double T = c->Gamma_1 * e_sp * (xi*0.5*p->m_H+(1.0-xi)*p->mu)/ p->k_B;
    double d_nlim, dt_therm1, dt_therm2, dt_therm, e_sp_min, e_th_min, e_min;
    if (T < p->tfloor) { c->edot[x] = 0.0; continue; }
    d_nlim = g->d[x]*IONFRACFLOOR;
    d_nlim = d_nlim < p->d_nlo ? d_nlim : p->d_nlo;
    if ((c->nHdot[x] < 0) && (g->s[x] < 1.0001*d_nlim)) { c->edot[x] = 0.0; continue; }
    c->edot[x] = c->ph_rate[x] * p->e_gamma * n_H
      - recomb_cool_rate_coef(T) * p->time_unit * n_Hplus * n_e
      + lya_cool_rate(n_H, n_e, T) * p->time_unit;
    if (c->edot[x] == 0.0) {
      dt_therm1 = dt_therm2 = LARGE;
    } else if (c->edot[x] > 0.0) {
      dt_therm1 = p->max_de_iter * g->E[x] / c->edot[x];
      dt_therm2 = p->max_de_therm_iter * e_thermal / c->edot[x];
    } else {
      e_sp_min = p->tfloor * p->k_B / ((xi*0.5*p->m_H+(1.0-xi)*p->mu) * c->Gamma_1);
      e_th_min = e_sp_min * g->d[x];
      e_min = KE(g,x) + e_th_min;
      if ((e_thermal/(1.0+p->max_de_therm_iter) < e_th_min) &&
          (g->E[x]/(1.0+p->max_de_iter) < e_min)) continue;
      dt_therm1 = -p->max_de_iter / (1+p->max_de_iter) * g->E[x] / c->edot[x];
      dt_therm2 = -p->max_de_therm_iter / (1+p->max_de_therm_iter) * e_thermal / c->edot[x];
    }
    dt_therm = (dt_therm1 < dt_therm2) ? dt_therm1 : dt_therm2;
    if (dt_therm < dt_therm_min) dt_therm_min = dt_therm;
  }
  return dt_therm_min;
}

static void ionization_update(IonCtx *c, double dt)                        /* :565-590 */
{
  OrcGrid *g = c->g; int i,j,k;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double d_nlim = g->d[x]*IONFRACFLOOR;
    d_nlim = d_nlim < c->p->d_nlo ? d_nlim : c->p->d_nlo;
    if ((c->nHdot[x] > 0) || (g->s[x] > 1.0001*d_nlim)) {
      g->E[x] += c->edot[x] * dt;
      g->s[x] += c->nHdot[x] * dt * c->p->m_H;
    }
  }
}

static double compute_dt_hydro(IonCtx *c, double cour_no)                  /* :593-678 */
{
  OrcGrid *g = c->g; int i,j,k;
  const double Gamma = g->gamma, Gamma_1 = c->Gamma_1;
  double max_dti = 0.0;
  IONLOOP(c) {
    size_t x = IDX(g,k,j,i);
    double di = 1.0/g->d[x];
    double v1 = g->M1[x]*di, v2 = g->M2[x]*di, v3 = g->M3[x]*di;
    double qsq = v1*v1 + v2*v2 + v3*v3;
    double pr = MAX(Gamma_1*(g->E[x] - 0.5*g->d[x]*qsq), TINY_NUMBER);
    double asq = Gamma*pr*di;
    if (g->nx1 > 1) max_dti = MAX(max_dti,(fabs(v1)+sqrt(asq))/g->dx1);
    if (g->nx2 > 1) max_dti = MAX(max_dti,(fabs(v2)+sqrt(asq))/g->dx2);
    if (g->nx3 > 1) max_dti = MAX(max_dti,(fabs(v3)+sqrt(asq))/g->dx3);
  }
  return cour_no/max_dti;
}

/* ionradplane_3d.c:88-407, dir = -1 branch (:253-322), root level */
static void get_ph_rate_plane(IonCtx *c, double flux0, double *edgeflux, int finegrid)
{
  OrcGrid *g = c->g; const OrcIonPar *p = c->p;
  const int e1 = g->nx1+1, e2 = g->nx2+1;
  const int s = c->is, e = c->ie, fixed = 0;
  const double cell_len = g->dx1;
  int i,j,k,ii;
  for (k=c->ks; k<=c->ke; k++) for (j=c->js; j<=c->je; j++) {
    double *ef = edgeflux + ((size_t)(k-c->ks)*e2 + (size_t)(j-c->js))*e1;
    double flux = finegrid ? ef[fixed] : flux0, flux_frac = 0.0;      /* ionradplane_3d.c:263-272 */
    for (i=s; i<=e; i++) {
      size_t x = IDX(g,k,j,i);
      double n_H, tau, etau, kph;
      ef[i-s] = flux;
      n_H = g->s[x] / p->m_H;
      tau = p->sigma_ph * n_H * g->dx1;
      etau = exp(-tau);
      kph = flux * (1.0-etau) / (n_H*cell_len);
      c->ph_rate[x] += kph;
      flux *= etau;
      flux_frac = flux / (ef[fixed] + 1e-12);
      if (flux_frac < MINFLUXFRAC) {
        for (ii=i; ii<=e; ii++) ef[ii-s+1] = 0.0;
        break;
      }
    }
    ef[e-s+1] = flux_frac < MINFLUXFRAC ? 0.0 : flux;
  }
}

int orc_ion_radtransfer(OrcGrid *g, const OrcIonPar *p, double flux_i, double time, double cour_no,
                        double *dt_io, double *edgeflux, int *nchem_out, int *ntherm_out)
{
  return orc_ion_radtransfer_root(g, p, flux_i, time, cour_no, dt_io, edgeflux, nchem_out, ntherm_out, NULL);
}

int orc_ion_radtransfer_root(OrcGrid *g, const OrcIonPar *p, double flux_i, double time, double cour_no,
                             double *dt_io, double *edgeflux, int *nchem_out, int *ntherm_out, double *tcoarse_out)
{
  IonCtx ctx, *c = &ctx;
  const size_t ncell = (size_t)N1(g)*N2(g)*N3(g);
  double dt_chem, dt_therm, dt_hydro, dt, dt_done, grid_dt = *dt_io;
  int niter = 0, hydro_done, nchem, ntherm, i, j, k;

  c->g = g; c->p = p; c->Gamma_1 = g->gamma - 1.0;
  c->is = NG; c->ie = g->nx1+NG-1; c->js = NG; c->je = g->nx2+NG-1; c->ks = NG; c->ke = g->nx3+NG-1;
  c->ph_rate = calloc(ncell,sizeof(double)); c->edot = calloc(ncell,sizeof(double));
  c->nHdot = calloc(ncell,sizeof(double)); c->e_init = calloc(ncell,sizeof(double));
  c->e_th_init = calloc(ncell,sizeof(double)); c->x_init = calloc(ncell,sizeof(double));
  c->last_sign = calloc(ncell,sizeof(int)); c->sign_count = calloc(ncell,sizeof(int));

  apply_temp_floor(c);                                                     /* :896 */
  apply_neutral_floor(c);                                                  /* :900 */
  save_energy_and_x(c);                                                    /* :905 */

  dt_done = 0.0; hydro_done = 0; nchem = ntherm = 0;
  while (!hydro_done) {                                                    /* :919, root level */
    IONLOOP(c) c->ph_rate[IDX(g,k,j,i)] = 0.0;                             /* :922 */
    get_ph_rate_plane(c, orc_ion_entering_flux(flux_i, time), edgeflux, 0);   /* :929 */
    dt_chem = compute_chem_rates(c);                                       /* :935 */
    dt_therm = compute_therm_rates(c);                                     /* :938 */
    if (dt_chem < dt_therm) nchem++; else ntherm++;
    dt = MIN(dt_therm, dt_chem);
    if (dt_done + dt > grid_dt) { dt = grid_dt - dt_done; hydro_done = 1; }
    ionization_update(c, dt);
    dt_done += dt;
    niter++;
    apply_temp_floor(c);
    apply_neutral_floor(c);
    if (check_range(c)) { grid_dt = dt_done; break; }
    if (hydro_done) break;
    dt_hydro = compute_dt_hydro(c, cour_no);
    if (dt_hydro < dt_done) { grid_dt = dt_done; break; }
  }
  if (niter == p->maxiter) grid_dt = dt_done;                              /* :1022 */
  if (tcoarse_out) *tcoarse_out = dt_done;                                 /* :1028 tcoarse = dt_done */

  free(c->ph_rate); free(c->edot); free(c->nHdot); free(c->e_init); free(c->e_th_init); free(c->x_init);
  free(c->last_sign); free(c->sign_count);
  *dt_io = grid_dt;
  if (nchem_out) *nchem_out = nchem;
  if (ntherm_out) *ntherm_out = ntherm;
  return niter;
}

int orc_ion_radtransfer_fine(OrcGrid *g, const OrcIonPar *p, double tcoarse, double *dt_io, double *edgeflux,
                             int *nchem_out, int *ntherm_out)
{
  IonCtx ctx, *c = &ctx;
  const size_t ncell = (size_t)N1(g)*N2(g)*N3(g);
  double dt_chem, dt_therm, dt, dt_done, grid_dt = *dt_io;
  int niter = 0, coarsetime_done = 0, nchem = 0, ntherm = 0, i, j, k;

  c->g = g; c->p = p; c->Gamma_1 = g->gamma - 1.0;
  c->is = NG; c->ie = g->nx1+NG-1; c->js = NG; c->je = g->nx2+NG-1; c->ks = NG; c->ke = g->nx3+NG-1;
  c->ph_rate = calloc(ncell,sizeof(double)); c->edot = calloc(ncell,sizeof(double));
  c->nHdot = calloc(ncell,sizeof(double)); c->e_init = calloc(ncell,sizeof(double));
  c->e_th_init = calloc(ncell,sizeof(double)); c->x_init = calloc(ncell,sizeof(double));
  c->last_sign = calloc(ncell,sizeof(int)); c->sign_count = calloc(ncell,sizeof(int));

  apply_temp_floor(c);
  apply_neutral_floor(c);
  save_energy_and_x(c);
  dt_done = 0.0;
  while (1) {                                                              /* :919 finegrid */
    IONLOOP(c) c->ph_rate[IDX(g,k,j,i)] = 0.0;
    get_ph_rate_plane(c, 0.0, edgeflux, 1);
    dt_chem = compute_chem_rates(c);
    dt_therm = compute_therm_rates(c);
    if (dt_chem < dt_therm) nchem++; else ntherm++;
    dt = MIN(dt_therm, dt_chem);
    if (dt_done + dt > tcoarse) { dt = tcoarse - dt_done; coarsetime_done = 1; }   /* :958-961 */
    ionization_update(c, dt);
    dt_done += dt;
    niter++;
    apply_temp_floor(c);
    apply_neutral_floor(c);
    if (coarsetime_done) { grid_dt = dt_done; break; }                     /* :1008-1012 */
  }
  free(c->ph_rate); free(c->edot); free(c->nHdot); free(c->e_init); free(c->e_th_init); free(c->x_init);
  free(c->last_sign); free(c->sign_count);
  *dt_io = grid_dt;
  if (nchem_out) *nchem_out = nchem;
  if (ntherm_out) *ntherm_out = ntherm;
  return niter;
}

/* ================================================================================================
 * SMR: smr.c, ionrad_smr.c (one parent Grid, one child Grid, same process)
 * ============================================================================================== */
static double *var_of(const OrcGrid *g, int v)
{
  switch (v) { case 0: return g->d; case 1: return g->M1; case 2: return g->M2; case 3: return g->M3; case 4: return g->E; default: return g->s; }
}

/* 8-cell conservative average of variable v at child cells (k,j,i)..(+1): smr.c:1874-1964 summation order */
static double restrict8(const OrcGrid *ch, int v, int k, int j, int i)
{
  const double *a = var_of(ch, v);
  double sum = a[IDX(ch,k,j,i)] + a[IDX(ch,k,j,i+1)];
  sum += a[IDX(ch,k,j+1,i)] + a[IDX(ch,k,j+1,i+1)];
  sum += a[IDX(ch,k+1,j,i)] + a[IDX(ch,k+1,j,i+1)] + a[IDX(ch,k+1,j+1,i)] + a[IDX(ch,k+1,j+1,i+1)];
  return sum*0.125;
}

static void restrict_inject(OrcGrid *pa, const OrcGrid *ch, const OrcOverlap *o, int v)
{
  int k,j,i,kk,jj,ii;
  double *dst = var_of(pa, v);
  for (k=o->cs[2], kk=NG; k<=o->ce[2]; k++, kk+=2)
    for (j=o->cs[1], jj=NG; j<=o->ce[1]; j++, jj+=2)
      for (i=o->cs[0], ii=NG; i<=o->ce[0]; i++, ii+=2)
        dst[IDX(pa,k,j,i)] = restrict8(ch, v, kk, jj, ii);
}

void orc_ion_restrict(OrcGrid *pa, const OrcGrid *ch, const OrcOverlap *o)        /* smr.c:85-1205 */
{
  restrict_inject(pa, ch, o, 4);
  restrict_inject(pa, ch, o, 5);
}

void orc_restrict_correct(OrcGrid *pa, const OrcGrid *ch, const OrcOverlap *o, double dt_parent)
{
  const int nvar = 5 + pa->nscal;
  int v, dim, a, b;
  for (v=0; v<nvar; v++) restrict_inject(pa, ch, o, v);                            /* Steps 3a, 1b */
  /* Step 2a with Step 3c: flux correction of the coarse cells just outside the child (smr.c:1497-1577, 2019-2227) */
  for (dim=0; dim<6; dim++) {
    const int d = dim/2, side = dim%2;
    const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;      /* the two transverse axes, ascending */
    const double dx = (d == 0) ? pa->dx1 : (d == 1) ? pa->dx2 : pa->dx3;
    const double q = side ? (dt_parent/dx) : -(dt_parent/dx);
    const int nch[3] = { ch->nx1, ch->nx2, ch->nx3 };
    if (!o->face[dim]) continue;
    for (b=o->cs[t2]; b<=o->ce[t2]; b++) for (a=o->cs[t1]; a<=o->ce[t1]; a++) {
      int pc[3], pf[3], cf[3];                   /* parent cell to correct, parent face cell, child face cell (first of 2x2) */
      size_t xc, xf, c00, c01, c10, c11;
      pc[t1] = a; pc[t2] = b; pc[d] = side ? o->ce[d]+1 : o->cs[d]-1;
      pf[t1] = a; pf[t2] = b; pf[d] = side ? o->ce[d]+1 : o->cs[d];
      cf[t1] = NG + 2*(a - o->cs[t1]); cf[t2] = NG + 2*(b - o->cs[t2]); cf[d] = side ? NG + nch[d] : NG;
      xc = IDX(pa,pc[2],pc[1],pc[0]); xf = IDX(pa,pf[2],pf[1],pf[0]);
      c00 = IDX(ch,cf[2],cf[1],cf[0]);
      { int e1[3] = {cf[0],cf[1],cf[2]}; e1[t1]++; c01 = IDX(ch,e1[2],e1[1],e1[0]); }
      { int e2[3] = {cf[0],cf[1],cf[2]}; e2[t2]++; c10 = IDX(ch,e2[2],e2[1],e2[0]); }
      { int e3[3] = {cf[0],cf[1],cf[2]}; e3[t1]++; e3[t2]++; c11 = IDX(ch,e3[2],e3[1],e3[0]); }
      for (v=0; v<nvar; v++) {
        const double *Fc = ch->Fsave[d], *Fp = pa->Fsave[d];
        /* (f[b][a] + f[b][a+1]) + (f[b+1][a] + f[b+1][a+1]), x 0.25: smr.c:2046-2097 and the x2/x3 analogues */
        double r = Fc[6*c00+v] + Fc[6*c01+v];
        r += Fc[6*c10+v] + Fc[6*c11+v];
        r *= 0.25;
        var_of(pa,v)[xc] -= q*(Fp[6*xf+v] - r);
      }
    }
  }
}

static double mcd_slope(double vl, double vc, double vr)                           /* smr.c:3478-3495 */
{
  double dvl = (vc - vl), dvr = (vr - vc), dv, dvm;
  if (dvl > 0.0 && dvr > 0.0) {
    dv = 2.0*(dvl < dvr ? dvl : dvr);
    dvm = 0.5*(dvl + dvr);
    return (dvm < dv ? dvm : dv);
  } else if (dvl < 0.0 && dvr < 0.0) {
    dv = 2.0*(dvl > dvr ? dvl : dvr);
    dvm = 0.5*(dvl + dvr);
    return (dvm > dv ? dvm : dv);
  }
  return 0.0;
}

static double pres_like(const OrcGrid *g, size_t x)                                /* smr.c:3209: E - KE (not x Gamma_1) */
{
  return g->E[x] - 0.5*(SQR(g->M1[x]) + SQR(g->M2[x]) + SQR(g->M3[x]))/g->d[x];
}

/* ProCon (smr.c:3068-3325) for the coarse cell (pk,pj,pi) of the parent, written to the 2x2x2 child cells at (ck,cj,ci) */
static void procon(const OrcGrid *pa, int pk, int pj, int pi, OrcGrid *ch, int ck, int cj, int ci)
{
  const size_t x0 = IDX(pa,pk,pj,pi), xim = IDX(pa,pk,pj,pi-1), xip = IDX(pa,pk,pj,pi+1);
  const size_t xjm = IDX(pa,pk,pj-1,pi), xjp = IDX(pa,pk,pj+1,pi), xkm = IDX(pa,pk-1,pj,pi), xkp = IDX(pa,pk+1,pj,pi);
  const int nvar = 5 + pa->nscal;
  int v, i, j, k;
  double PC[2][2][2][6];
  for (v=0; v<nvar; v++) {
    double dq1, dq2, dq3, base;
    if (v == 4) {
      const double Pi = pres_like(pa,x0);
      dq1 = mcd_slope(pres_like(pa,xim), Pi, pres_like(pa,xip));
      dq2 = mcd_slope(pres_like(pa,xjm), Pi, pres_like(pa,xjp));
      dq3 = mcd_slope(pres_like(pa,xkm), Pi, pres_like(pa,xkp));
      base = Pi;
    } else {
      const double *a = var_of(pa, v);
      dq1 = mcd_slope(a[xim], a[x0], a[xip]);
      dq2 = mcd_slope(a[xjm], a[x0], a[xjp]);
      dq3 = mcd_slope(a[xkm], a[x0], a[xkp]);
      base = a[x0];
    }
    for (k=0; k<2; k++) for (j=0; j<2; j++) for (i=0; i<2; i++)
      PC[k][j][i][v] = base + (0.5*i - 0.25)*dq1 + (0.5*j - 0.25)*dq2 + (0.5*k - 0.25)*dq3;
  }
  for (k=0; k<2; k++) for (j=0; j<2; j++) for (i=0; i<2; i++) {
    double *c = PC[k][j][i];
    c[4] += 0.5*(SQR(c[1]) + SQR(c[2]) + SQR(c[3]))/c[0];                          /* :3240 */
    for (v=0; v<nvar; v++) var_of(ch,v)[IDX(ch,ck+k,cj+j,ci+i)] = c[v];
  }
}

void orc_prolongate(const OrcGrid *pa, OrcGrid *ch, const OrcOverlap *o)           /* smr.c:2359-2923 */
{
  const int ps0[3] = { NG, NG, NG }, pe0[3] = { ch->nx1+NG-1, ch->nx2+NG-1, ch->nx3+NG-1 };   /* PGrid.ijks/ijke on the child */
  int dim, a;
  for (dim=0; dim<6; dim++) {
    int cs[3], ps[3], pe[3], k, j, i, kk, jj, ii;
    if (!o->face[dim]) continue;
    for (a=0; a<3; a++) {
      cs[a] = o->cs[a] - (NG/2) - 1;             /* parent slab start (:2442-2463) */
      ps[a] = ps0[a] - NG; pe[a] = pe0[a] + NG;  /* child ghost box (:2667-2688) */
    }
    if (dim%2 == 0) pe[dim/2] = ps0[dim/2] - 1;
    else { ps[dim/2] = pe0[dim/2] + 1; cs[dim/2] = o->ce[dim/2]; }
    for (k=ps[2], kk=1; k<=pe[2]; k+=2, kk++)
      for (j=ps[1], jj=1; j<=pe[1]; j+=2, jj++)
        for (i=ps[0], ii=1; i<=pe[0]; i+=2, ii++)
          procon(pa, cs[2]+kk, cs[1]+jj, cs[0]+ii, ch, k, j, i);
  }
}

void orc_ion_prolong(const OrcGrid *pa, const double *efp, const OrcGrid *ch, double *efc, const OrcOverlap *o)
{ /* ionrad_smr.c:345-459 (snd, dim 0) + :34-343 (rcv).  The serial branch (:282-337) forms the child index as
   * `k*2 - Disp` with k parent-LOCAL, which is only right below a root-level parent (and reads out of bounds below
   * that); the MPI branch (:142-211) maps child edge (kk,jj) to coarse edge (kk/2, jj/2) of the overlap.
   * Both are the same map for level 1; written here in the overlap-relative form so it also holds for deeper
   * levels, where only the MPI branch can be executed (oracle/mpi_stub) to pin it. */
  const int pe1 = pa->nx1+1, pe2 = pa->nx2+1, ce1 = ch->nx1+1, ce2 = ch->nx2+1;
  const int fixed_p = o->cs[0] - NG, fixed_c = 0;
  int k, j;
  if (!o->face[0]) return;
  for (k=o->cs[2]-NG; k<=o->ce[2]+1-NG; k++) for (j=o->cs[1]-NG; j<=o->ce[1]+1-NG; j++) {
    const double val = efp[((size_t)k*pe2 + j)*pe1 + fixed_p];
    const int ks = 2*(k - (o->cs[2]-NG)), js = 2*(j - (o->cs[1]-NG));
#define EFC(kk,jj) efc[((size_t)(kk)*ce2 + (jj))*ce1 + fixed_c]
    EFC(ks,js) = val;
    if (j < (o->ce[1]+1-NG)) {
      if (k < (o->ce[2]+1-NG)) { EFC(ks+1,js+1) = val; EFC(ks,js+1) = val; EFC(ks+1,js) = val; }
      else EFC(ks,js+1) = val;
    } else if (k < (o->ce[2]+1-NG)) EFC(ks+1,js) = val;
#undef EFC
  }
}
