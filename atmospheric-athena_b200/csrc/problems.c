/* problems.c -- host-side C problem generators and user hooks for the three target problems.
 *
 * In the reference these live in problem files that stay C (src/prob/blast.c, ifront.c, ioniz_sphere.c) and
 * reach the hot path only through problem(), Userwork_in_loop() and the StaticGravPot function pointer
 * (globals.h:24).  This file gives the Python host driver / bench the same three things as plain C
 * functions over SoA arrays ([n3][n2][n1] with 4 ghost zones, i fastest), so synthetic grids of any size
 * can be initialised without the reference binary.  It is host code: no CUDA, not part of the hot path.
 */
#include <math.h>
#include <stddef.h>

#define NG 4
#define TINY_NUMBER 1.0e-20
#define SQR(x) ((x)*(x))
#define MAXV(a,b) (((a) > (b)) ? (a) : (b))
#define MINV(a,b) (((a) < (b)) ? (a) : (b))

typedef struct {
  int nx1, nx2, nx3;
  double x1min, x2min, x3min, dx1, dx2, dx3;
  double *d, *M1, *M2, *M3, *E, *s;   /* s may be NULL (NSCALARS=0) */
} ProbGrid;

static size_t idx(const ProbGrid *g, int k, int j, int i)
{
  return ((size_t)k*(g->nx2+2*NG) + (size_t)j)*(g->nx1+2*NG) + (size_t)i;
}
static void ccpos(const ProbGrid *g, int i, int j, int k, double *x1, double *x2, double *x3)
{ /* cc_pos.c */
  *x1 = g->x1min + ((double)(i - NG) + 0.5)*g->dx1;
  *x2 = g->x2min + ((double)(j - NG) + 0.5)*g->dx2;
  *x3 = g->x3min + ((double)(k - NG) + 0.5)*g->dx3;
}

/* prob/blast.c:20-90 (hydro) */
void ab2_problem_blast(ProbGrid *g, double gamma, double pamb, double prat, double damb, double drat, double radius)
{
  int i,j,k; double x1,x2,x3;
  for (k=NG; k<g->nx3+NG; k++) for (j=NG; j<g->nx2+NG; j++) for (i=NG; i<g->nx1+NG; i++) {
    size_t c = idx(g,k,j,i);
    double rad, P = pamb, d = damb;
    ccpos(g,i,j,k,&x1,&x2,&x3);
    rad = sqrt(x1*x1 + x2*x2 + x3*x3);
    if (rad < radius) { P = prat*pamb; d = drat*damb; }
    g->d[c] = d; g->M1[c] = d*0.0; g->M2[c] = d*0.0; g->M3[c] = d*0.0;
    g->E[c] = P/(gamma-1.0) + 0.5*d*(SQR(0.0)+SQR(0.0)+SQR(0.0));
    if (g->s) g->s[c] = 0.0;
  }
}

/* prob/ifront.c:44-66 (loops to ie+1 etc. there; the extra layer is ghost and overwritten by bvals) */
void ab2_problem_ifront(ProbGrid *g, double gamma, double n_H, double m_H, double cs)
{
  int i,j,k;
  for (k=NG; k<=g->nx3+NG; k++) for (j=NG; j<=g->nx2+NG; j++) for (i=NG; i<=g->nx1+NG; i++) {
    size_t c = idx(g,k,j,i);
    double rho = n_H*m_H, pressure = rho*cs*cs;
    g->d[c] = rho; g->M1[c] = 0.0; g->M2[c] = 0.0; g->M3[c] = 0.0;
    g->E[c] = pressure/(gamma-1.0);
    g->s[c] = rho;
  }
}

/* prob/ioniz_sphere.c */
typedef struct { double GM, K, Cp, rp, rreset2, rho0, Rsoft, rin, rout, rhoedge, rhoout, Gamma, Gamma_1; } SphereState;
static SphereState S;

void ab2_ioniz_sphere_setup(double gamma, double mu, double cs, double rp, double mp, double np)
{ /* ioniz_sphere.c:52-100 */
  double Ggrav = 6.67e-8, rhop, powindex;
  S.Gamma = gamma; S.Gamma_1 = gamma - 1.0;
  S.rp = rp;
  S.GM = Ggrav * mp;
  rhop = np * mu;
  S.Rsoft = 0.01*rp;
  S.rin = 0.5*rp;
  S.rreset2 = 0.5625*rp*rp;
  powindex = 1.0/S.Gamma_1;
  S.K = pow(rhop,-S.Gamma_1)*cs*cs;
  S.rho0 = pow( pow(rhop,S.Gamma_1) - S.Gamma_1/S.Gamma*S.GM/S.K*(1.0/rp - 1.0/S.rin),powindex);
  S.Cp = pow(S.rho0,S.Gamma_1) - (S.Gamma_1/S.Gamma)*S.GM/S.K/S.rin;
  S.rhoedge = rhop/10;
  S.rout = 1./(S.Gamma/S.Gamma_1/S.GM*S.K*(pow(S.rhoedge, S.Gamma_1) - pow(S.rho0, S.Gamma_1)) + 1./S.rin);
  S.rhoout = S.rhoedge/10000.;
}

void ab2_problem_ioniz_sphere(ProbGrid *g)
{ /* ioniz_sphere.c:102-125 */
  int i,j,k; double x1,x2,x3;
  const double powindex = 1.0/S.Gamma_1;
  for (k=NG; k<=g->nx3+NG; k++) for (j=NG; j<=g->nx2+NG; j++) for (i=NG; i<=g->nx1+NG; i++) {
    size_t c = idx(g,k,j,i);
    double rad;
    ccpos(g,i,j,k,&x1,&x2,&x3);
    rad = sqrt(x1*x1 + x2*x2 + x3*x3);
    g->M1[c] = 0.0; g->M2[c] = 0.0; g->M3[c] = 0.0;
    if (rad <= S.rin) {
      g->d[c] = S.rho0;
      g->E[c] = S.K*pow(g->d[c],S.Gamma)/S.Gamma_1;
      g->s[c] = g->d[c];
    } else if (rad > S.rout) {
      g->d[c] = S.rhoout;
      g->E[c] = S.K*pow(S.rhoedge,S.Gamma)/S.Gamma_1;
      g->s[c] = g->d[c] * 1.0e-4;
    } else {
      g->d[c] = pow(S.Gamma_1/S.Gamma*S.GM/S.K/MAXV(rad,TINY_NUMBER) + S.Cp,powindex);
      g->E[c] = S.K*pow(g->d[c],S.Gamma)/S.Gamma_1;
      g->s[c] = g->d[c];
    }
  }
}

/* ioniz_sphere.c:316-330 PlanetPot (non-SHEARING_BOX branch): the StaticGravPot the problem enrols */
double ab2_planet_pot(double x1, double x2, double x3)
{
  double rad = sqrt(SQR(x1)+SQR(x2)+SQR(x3));
  double adist = 7.48e11;
  double GMstar = 6.67e-8 * 1.99e33;
  double omega = sqrt(GMstar / (pow(adist,3)));
  double radstar = sqrt(SQR(x1+adist) + SQR(x2) + SQR(x3));
  double rcentrif = sqrt(SQR(x1+adist) + SQR(x2));
  return -S.GM/(rad+S.Rsoft)-GMstar/radstar -.5*SQR(omega*rcentrif);
}

/* ioniz_sphere.c:255-306 Userwork_in_loop: list the active cells with r^2 <= rreset2 and the fixed state
 * they are reset to every step.  Returns the count; with cells==NULL only counts.  vals = 6 x n (d,M1,M2,M3,E,s). */
long ab2_ioniz_sphere_reset_table(const ProbGrid *g, long *cells, double *vals, long cap)
{
  int i,j,k; double x1,x2,x3; long n = 0;
  const double powindex = 1.0/S.Gamma_1;
  for (k=NG; k<g->nx3+NG; k++) for (j=NG; j<g->nx2+NG; j++) for (i=NG; i<g->nx1+NG; i++) {
    double rad2;
    ccpos(g,i,j,k,&x1,&x2,&x3);
    rad2 = x1*x1 + x2*x2 + x3*x3;
    if (rad2 <= S.rreset2) {
      if (cells != NULL && n < cap) {
        double myrho = pow(S.Gamma_1/S.Gamma*S.GM/S.K/MAXV(sqrt(rad2),TINY_NUMBER) + S.Cp,powindex);
        myrho = MINV(myrho, S.rho0);
        cells[n] = (long)idx(g,k,j,i);
        vals[0*cap+n] = myrho; vals[1*cap+n] = 0.0; vals[2*cap+n] = 0.0; vals[3*cap+n] = 0.0;
        vals[4*cap+n] = S.K*pow(myrho,S.Gamma)/S.Gamma_1;
        vals[5*cap+n] = myrho;
      }
      n++;
    }
  }
  return n;
}
