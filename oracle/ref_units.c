/* ref_units.c -- thin C harness around the UNMODIFIED reference's pencil-level routines.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.py into oracle/_ref/<cfg>/libref_units.so by
 * compiling the reference's own convert_var.c, lr_states_plm.c, esystem_prim.c, roe.c (+hlle.c) and
 * esystem_roe.c where they lie under /root/reference/src (nothing is copied), plus this file, which
 * only supplies the globals those units expect and flat-array entry points for ctypes.
 * It is used to generate the function-level golden vectors in tests/golden/units_*.npz and to check
 * oracle/athena_oracle.c against the reference function by function.
 *
 * All state vectors are passed as 6 doubles (d,Mx|Vx,My|Vy,Mz|Vz,E|P,s|r); with NSCALARS=0 the sixth
 * entry is ignored on input and returned as 0.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "defs.h"
#include "athena.h"
#include "prototypes.h"

/* globals.h symbols the units reference */
Real Gamma, Gamma_1, Gamma_2;
#ifdef ROE_FLUX
extern Real etah;                       /* defined in rsolvers/roe.c:33 */
#else
static Real etah;                       /* hllc.c takes no H-correction */
#endif

void ath_error(char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
  abort();
}

int ath_perr(const int level, const char *fmt, ...) { (void)level; (void)fmt; return 0; }   /* hllc.c:151 diagnostic */

#define NV ((int)(sizeof(Cons1DS)/sizeof(Real)))

static void in_c(Cons1DS *u, const double *x) { memcpy(u, x, sizeof(Cons1DS)); }
static void in_p(Prim1DS *w, const double *x) { memcpy(w, x, sizeof(Prim1DS)); }
static void out6(double *x, const void *s) { x[5] = 0.0; memcpy(x, s, NV*sizeof(Real)); }

int ref_nvar(void) { return NV; }

void ref_set_gamma(double g) { Gamma = g; Gamma_1 = g - 1.0; Gamma_2 = g - 2.0; }

void ref_cons_to_prim(const double U[6], double W[6])
{
  Cons1DS u; Prim1DS w; Real bx = 0.0;
  in_c(&u, U); w = Cons1D_to_Prim1D(&u, &bx); out6(W, &w);
}

void ref_prim_to_cons(const double W[6], double U[6])
{
  Cons1DS u; Prim1DS w; Real bx = 0.0;
  in_p(&w, W); u = Prim1D_to_Cons1D(&w, &bx); out6(U, &u);
}

double ref_cfast(const double U[6])
{
  Cons1DS u; Real bx = 0.0;
  in_c(&u, U); return cfast(&u, &bx);
}

void ref_fluxes(const double Ul[6], const double Ur[6], const double Wl[6], const double Wr[6],
                double etah_in, double F[6])
{
  Cons1DS ul, ur, f; Prim1DS wl, wr;
  in_c(&ul, Ul); in_c(&ur, Ur); in_p(&wl, Wl); in_p(&wr, Wr);
  memset(&f, 0, sizeof f);
  etah = etah_in;
  fluxes(ul, ur, wl, wr, 0.0, &f);
  out6(F, &f);
}

/* W, Wl, Wr: n x 6 doubles.  Same call as integrate_3d_ctu.c:248. */
void ref_lr_states(const double *W, int n, double dt, double dx, int il, int iu, double *Wl, double *Wr)
{
  static int inited_n = 0;
  Prim1DS *w = malloc(n*sizeof(Prim1DS)), *wl = calloc(n,sizeof(Prim1DS)), *wr = calloc(n,sizeof(Prim1DS));
  Real *bxc = calloc(n,sizeof(Real));
  int i;
  if (inited_n < n) {
    MeshS M; DomainS D; DomainS *pD = &D; GridS G; int dpl = 1;
    memset(&M,0,sizeof M); memset(&D,0,sizeof D); memset(&G,0,sizeof G);
    G.Nx[0] = n; G.Nx[1] = 1; G.Nx[2] = 1;
    D.Grid = &G;
    M.NLevels = 1; M.DomainsPerLevel = &dpl; M.Domain = &pD;
    if (inited_n) lr_states_destruct();
    lr_states_init(&M);
    inited_n = n;
  }
  for (i=0; i<n; i++) in_p(&w[i], W + 6*i);
  lr_states(NULL, w, bxc, dt, dx, il, iu, wl, wr, 1);
  for (i=0; i<n; i++) { out6(Wl + 6*i, &wl[i]); out6(Wr + 6*i, &wr[i]); }
  free(w); free(wl); free(wr); free(bxc);
}
