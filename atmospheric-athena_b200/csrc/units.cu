// units.cu -- function-level entry points: the pencil routines of physics.cuh on flat host arrays, so parity
// tests can compare lr_states / fluxes / Cons1D_to_Prim1D one call at a time against the reference's own
// objects (tests/golden/units_*.npz).  Rows of 6 doubles: (d,Mx,My,Mz,E,s) or (d,Vx,Vy,Vz,P,r).
#include <vector>
#include "../../include/athena_b200.h"
#include "../../include/athena_b200_units.h"
#include "engine.cuh"
#include "physics.cuh"
#include "glibc_math.cuh"

static __device__ __forceinline__ S6 ld6(const double* p) { S6 r; r.d=p[0]; r.x=p[1]; r.y=p[2]; r.z=p[3]; r.e=p[4]; r.s=p[5]; return r; }
static __device__ __forceinline__ void st6(double* p, const S6& r) { p[0]=r.d; p[1]=r.x; p[2]=r.y; p[3]=r.z; p[4]=r.e; p[5]=r.s; }

template <int NS> __global__ void k_unit_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                                              const double* etah, double Gamma, double* F)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  S6 f;
  const S6 ul = ld6(Ul+6*t), ur = ld6(Ur+6*t), wl = ld6(Wl+6*t), wr = ld6(Wr+6*t);
  { Arith<true> ar; roe_flux<NS>(ul, ur, wl, wr, etah[t], Gamma, Gamma-1.0, f, ar);
    if (ar.bad) { Arith<false> ax; roe_flux<NS>(ul, ur, wl, wr, etah[t], Gamma, Gamma-1.0, f, ax); } }
  st6(F+6*t, f);
}

template <int NS> __global__ void k_unit_hllc(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr, double Gamma, double* F)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  S6 f;
  const S6 ul = ld6(Ul+6*t), ur = ld6(Ur+6*t), wl = ld6(Wl+6*t), wr = ld6(Wr+6*t);
  AB2_GUARDED(hllc_flux<NS>(ul, ur, wl, wr, Gamma, Gamma-1.0, f, ar));
  st6(F+6*t, f);
}

template <int NS> __global__ void k_unit_convert(int n, const double* in, double Gamma, int to_prim, double* out, double* cf)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  const S6 a = ld6(in+6*t);
  S6 o; double c;
  { Arith<true> ar; o = to_prim ? cons_to_prim<NS>(a, Gamma-1.0, ar) : prim_to_cons<NS>(a, Gamma-1.0, ar); c = cfast(a, Gamma, Gamma-1.0, ar);
    if (ar.bad) { Arith<false> ax; o = to_prim ? cons_to_prim<NS>(a, Gamma-1.0, ax) : prim_to_cons<NS>(a, Gamma-1.0, ax); c = cfast(a, Gamma, Gamma-1.0, ax); } }
  st6(out+6*t, o);
  if (cf) cf[t] = c;
}

// one thread per (pencil, cell): cells il-1..iu+1 write Wl[i+1], Wr[i]  (lr_states_plm.c:96)
template <int NS> __global__ void k_unit_plm(int npen, int n, const double* W, double Gamma, double dtodx, int il, int iu, double* Wl, double* Wr, bool hll)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  const int ncell = iu - il + 3;
  if (t >= npen*ncell) return;
  const int pen = t / ncell, i = il - 1 + t % ncell;
  const double* w = W + (size_t)pen*n*6;
  S6 wl, wr;
  const S6 wm = ld6(w+6*(i-1)), w0 = ld6(w+6*i), wp = ld6(w+6*(i+1));
  AB2_GUARDED(plm_cell<NS>(wm, w0, wp, dtodx, Gamma, wl, wr, ar, hll));
  st6(Wl + ((size_t)pen*n + i+1)*6, wl);
  st6(Wr + ((size_t)pen*n + i)*6, wr);
}

// PPM: one thread per (pencil, cell i in [il-1, iu+1]); slopes of i-1, i, i+1 are recomputed per thread
template <int NS> __global__ void k_unit_ppm(int npen, int n, const double* W, double Gamma, double dtodx, int il, int iu, double* Wl, double* Wr, bool hll)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  const int ncell = iu - il + 3;
  if (t >= npen*ncell) return;
  const int pen = t / ncell, i = il - 1 + t % ncell;
  const double* w = W + (size_t)pen*n*6;
  S6 wl, wr;
  const S6 wmm = ld6(w+6*(i-2)), wm = ld6(w+6*(i-1)), w0 = ld6(w+6*i), wp = ld6(w+6*(i+1)), wpp = ld6(w+6*(i+2));
  AB2_GUARDED(
    double dm[6]; double d0[6]; double dp[6];
    const PrimEig em = prim_eig(wm, Gamma, ar); const PrimEig e0 = prim_eig(w0, Gamma, ar); const PrimEig ep = prim_eig(wp, Gamma, ar);
    ppm_slope<NS>(wmm, wm, w0, em, dm, ar); ppm_slope<NS>(wm, w0, wp, e0, d0, ar); ppm_slope<NS>(w0, wp, wpp, ep, dp, ar);
    double Fl[6]; double Fr[6];
    Fl[0] = ppm_face(w0.d, wm.d, d0[0], dm[0], ar); Fr[0] = ppm_face(wp.d, w0.d, dp[0], d0[0], ar);
    Fl[1] = ppm_face(w0.x, wm.x, d0[1], dm[1], ar); Fr[1] = ppm_face(wp.x, w0.x, dp[1], d0[1], ar);
    Fl[2] = ppm_face(w0.y, wm.y, d0[2], dm[2], ar); Fr[2] = ppm_face(wp.y, w0.y, dp[2], d0[2], ar);
    Fl[3] = ppm_face(w0.z, wm.z, d0[3], dm[3], ar); Fr[3] = ppm_face(wp.z, w0.z, dp[3], d0[3], ar);
    Fl[4] = ppm_face(w0.e, wm.e, d0[4], dm[4], ar); Fr[4] = ppm_face(wp.e, w0.e, dp[4], d0[4], ar);
    Fl[5] = NS ? Fr[0] : 0.0;
    Fr[5] = NS ? ppm_face(wp.s, w0.s, dp[5], dp[0], ar) : 0.0;
    ppm_cell<NS>(wm, w0, wp, Fl, Fr, e0, dtodx, wl, wr, hll));
  st6(Wl + ((size_t)pen*n + i+1)*6, wl);
  st6(Wr + ((size_t)pen*n + i)*6, wr);
}

// raw check of Arith<true>: out[0..n) = fast a/b (or NaN-tagged when the range test rejects), out2 = a/b; same for sqrt
__global__ void k_unit_ieee(int n, const double* a, const double* b, double* qf, double* qe, double* sf, double* se, int* nbad)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  Arith<true> ar;
  qf[t] = ar.div(a[t], b[t]);
  const unsigned bad_div = ar.bad; ar.bad = 0;
  sf[t] = ar.sqrt_(a[t]);
  const unsigned bad_sqrt = ar.bad;
  qe[t] = a[t]/b[t];
  se[t] = sqrt(a[t]);
  if (bad_div) { qf[t] = qe[t]; atomicAdd(&nbad[0], 1); }      // rejected operands take the plain-operator path
  if (bad_sqrt) { sf[t] = se[t]; atomicAdd(&nbad[1], 1); }
}

__global__ void k_unit_libm(int n, const double* x, const double* y, double* e, double* p)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  e[t] = gl_exp(x[t]);
  p[t] = gl_pow(y[t], x[t]);       // base y[t], exponent x[t]
}

struct DevBuf {
  double* p = nullptr;
  DevBuf(size_t n, const double* host = nullptr) { if (cudaMalloc(&p, n*sizeof(double)) != cudaSuccess) p = nullptr; else if (host) cudaMemcpy(p, host, n*sizeof(double), cudaMemcpyHostToDevice); else cudaMemset(p, 0, n*sizeof(double)); }
  ~DevBuf() { if (p) cudaFree(p); }
};

extern "C" {

int ab2_unit_roe_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                      int nscal, double gamma, const double* etah, double* F)
{
  DevBuf a(6*n,Ul), b(6*n,Ur), c(6*n,Wl), d(6*n,Wr), e(n,etah), f(6*n);
  if (!a.p || !b.p || !c.p || !d.p || !e.p || !f.p) { ab2_set_error("[ab2_unit_roe_flux]: cudaMalloc failed"); return -1; }
  if (nscal) k_unit_flux<1><<<(n+127)/128,128>>>(n,a.p,b.p,c.p,d.p,e.p,gamma,f.p);
  else k_unit_flux<0><<<(n+127)/128,128>>>(n,a.p,b.p,c.p,d.p,e.p,gamma,f.p);
  if (!ab2_check(cudaMemcpy(F, f.p, 6*n*sizeof(double), cudaMemcpyDeviceToHost), "unit flux")) return -1;
  return 0;
}

int ab2_unit_convert(int n, const double* in, int nscal, double gamma, int to_prim, double* out, double* cfast_out)
{
  DevBuf a(6*n,in), o(6*n), cf(n);
  if (!a.p || !o.p || !cf.p) { ab2_set_error("[ab2_unit_convert]: cudaMalloc failed"); return -1; }
  if (nscal) k_unit_convert<1><<<(n+127)/128,128>>>(n,a.p,gamma,to_prim,o.p,cfast_out ? cf.p : nullptr);
  else k_unit_convert<0><<<(n+127)/128,128>>>(n,a.p,gamma,to_prim,o.p,cfast_out ? cf.p : nullptr);
  if (!ab2_check(cudaMemcpy(out, o.p, 6*n*sizeof(double), cudaMemcpyDeviceToHost), "unit convert")) return -1;
  if (cfast_out && !ab2_check(cudaMemcpy(cfast_out, cf.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit cfast")) return -1;
  return 0;
}

int ab2_unit_ieee(int n, const double* a, const double* b, double* q_fast, double* q_exact, double* s_fast, double* s_exact, int* nrejected)
{
  DevBuf A(n,a), B(n,b), qf(n), qe(n), sf(n), se(n);
  int* nb = nullptr;
  if (!A.p || !B.p || !qf.p || !qe.p || !sf.p || !se.p || cudaMalloc(&nb, 2*sizeof(int)) != cudaSuccess) { ab2_set_error("[ab2_unit_ieee]: cudaMalloc failed"); return -1; }
  cudaMemset(nb, 0, 2*sizeof(int));
  k_unit_ieee<<<(n+127)/128,128>>>(n, A.p, B.p, qf.p, qe.p, sf.p, se.p, nb);
  bool ok = ab2_check(cudaMemcpy(q_fast, qf.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit ieee") &&
            ab2_check(cudaMemcpy(q_exact, qe.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit ieee") &&
            ab2_check(cudaMemcpy(s_fast, sf.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit ieee") &&
            ab2_check(cudaMemcpy(s_exact, se.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit ieee") &&
            ab2_check(cudaMemcpy(nrejected, nb, 2*sizeof(int), cudaMemcpyDeviceToHost), "unit ieee");
  cudaFree(nb);
  return ok ? 0 : -1;
}

int ab2_unit_libm(int n, const double* x, const double* y, double* exp_x, double* pow_y_x)
{
  DevBuf X(n,x), Y(n,y), E(n), Pw(n);
  if (!X.p || !Y.p || !E.p || !Pw.p) { ab2_set_error("[ab2_unit_libm]: cudaMalloc failed"); return -1; }
  k_unit_libm<<<(n+127)/128,128>>>(n, X.p, Y.p, E.p, Pw.p);
  if (!ab2_check(cudaMemcpy(exp_x, E.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit libm")) return -1;
  if (!ab2_check(cudaMemcpy(pow_y_x, Pw.p, n*sizeof(double), cudaMemcpyDeviceToHost), "unit libm")) return -1;
  return 0;
}

static int unit_lr_states(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                          int il, int iu, double* Wl, double* Wr, bool hll);
int ab2_unit_lr_states(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                       int il, int iu, double* Wl, double* Wr)
{ return unit_lr_states(npencil, n, W, nscal, gamma, dt, dx, il, iu, Wl, Wr, false); }
int ab2_unit_lr_states_hll(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                           int il, int iu, double* Wl, double* Wr)
{ return unit_lr_states(npencil, n, W, nscal, gamma, dt, dx, il, iu, Wl, Wr, true); }

int ab2_unit_lr_states_ppm(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                           int il, int iu, int hll, double* Wl, double* Wr)
{
  const size_t tot = (size_t)npencil*n*6;
  DevBuf w(tot,W), l(tot), r(tot);
  if (!w.p || !l.p || !r.p) { ab2_set_error("[ab2_unit_lr_states_ppm]: cudaMalloc failed"); return -1; }
  const int nt = npencil*(iu-il+3);
  if (nscal) k_unit_ppm<1><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p,hll != 0);
  else k_unit_ppm<0><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p,hll != 0);
  if (!ab2_check(cudaMemcpy(Wl, l.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit ppm")) return -1;
  if (!ab2_check(cudaMemcpy(Wr, r.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit ppm")) return -1;
  return 0;
}

int ab2_unit_hllc_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                       int nscal, double gamma, double* F)
{
  DevBuf a(6*n,Ul), b(6*n,Ur), c(6*n,Wl), d(6*n,Wr), f(6*n);
  if (!a.p || !b.p || !c.p || !d.p || !f.p) { ab2_set_error("[ab2_unit_hllc_flux]: cudaMalloc failed"); return -1; }
  if (nscal) k_unit_hllc<1><<<(n+127)/128,128>>>(n,a.p,b.p,c.p,d.p,gamma,f.p);
  else k_unit_hllc<0><<<(n+127)/128,128>>>(n,a.p,b.p,c.p,d.p,gamma,f.p);
  if (!ab2_check(cudaMemcpy(F, f.p, 6*n*sizeof(double), cudaMemcpyDeviceToHost), "unit hllc flux")) return -1;
  return 0;
}

} // extern "C"

static int unit_lr_states(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                          int il, int iu, double* Wl, double* Wr, bool hll)
{
  const size_t tot = (size_t)npencil*n*6;
  DevBuf w(tot,W), l(tot), r(tot);
  if (!w.p || !l.p || !r.p) { ab2_set_error("[ab2_unit_lr_states]: cudaMalloc failed"); return -1; }
  const int nt = npencil*(iu-il+3);
  if (nscal) k_unit_plm<1><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p,hll);
  else k_unit_plm<0><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p,hll);
  if (!ab2_check(cudaMemcpy(Wl, l.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit plm")) return -1;
  if (!ab2_check(cudaMemcpy(Wr, r.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit plm")) return -1;
  return 0;
}
