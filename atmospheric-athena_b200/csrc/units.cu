// units.cu -- function-level entry points: the pencil routines of physics.cuh on flat host arrays, so parity
// tests can compare lr_states / fluxes / Cons1D_to_Prim1D one call at a time against the reference's own
// objects (tests/golden/units_*.npz).  Rows of 6 doubles: (d,Mx,My,Mz,E,s) or (d,Vx,Vy,Vz,P,r).
#include <vector>
#include "../../include/athena_b200.h"
#include "engine.cuh"
#include "physics.cuh"

static __device__ __forceinline__ S6 ld6(const double* p) { S6 r; r.d=p[0]; r.x=p[1]; r.y=p[2]; r.z=p[3]; r.e=p[4]; r.s=p[5]; return r; }
static __device__ __forceinline__ void st6(double* p, const S6& r) { p[0]=r.d; p[1]=r.x; p[2]=r.y; p[3]=r.z; p[4]=r.e; p[5]=r.s; }

template <int NS> __global__ void k_unit_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                                              const double* etah, double Gamma, double* F)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  S6 f;
  roe_flux<NS>(ld6(Ul+6*t), ld6(Ur+6*t), ld6(Wl+6*t), ld6(Wr+6*t), etah[t], Gamma, Gamma-1.0, f);
  st6(F+6*t, f);
}

template <int NS> __global__ void k_unit_convert(int n, const double* in, double Gamma, int to_prim, double* out, double* cf)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  const S6 a = ld6(in+6*t);
  st6(out+6*t, to_prim ? cons_to_prim<NS>(a, Gamma-1.0) : prim_to_cons<NS>(a, Gamma-1.0));
  if (cf) cf[t] = cfast(a, Gamma, Gamma-1.0);
}

// one thread per (pencil, cell): cells il-1..iu+1 write Wl[i+1], Wr[i]  (lr_states_plm.c:96)
template <int NS> __global__ void k_unit_plm(int npen, int n, const double* W, double Gamma, double dtodx, int il, int iu, double* Wl, double* Wr)
{
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  const int ncell = iu - il + 3;
  if (t >= npen*ncell) return;
  const int pen = t / ncell, i = il - 1 + t % ncell;
  const double* w = W + (size_t)pen*n*6;
  S6 wl, wr;
  plm_cell<NS>(ld6(w+6*(i-1)), ld6(w+6*i), ld6(w+6*(i+1)), dtodx, Gamma, wl, wr);
  st6(Wl + ((size_t)pen*n + i+1)*6, wl);
  st6(Wr + ((size_t)pen*n + i)*6, wr);
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

int ab2_unit_lr_states(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                       int il, int iu, double* Wl, double* Wr)
{
  const size_t tot = (size_t)npencil*n*6;
  DevBuf w(tot,W), l(tot), r(tot);
  if (!w.p || !l.p || !r.p) { ab2_set_error("[ab2_unit_lr_states]: cudaMalloc failed"); return -1; }
  const int nt = npencil*(iu-il+3);
  if (nscal) k_unit_plm<1><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p);
  else k_unit_plm<0><<<(nt+127)/128,128>>>(npencil,n,w.p,gamma,dt/dx,il,iu,l.p,r.p);
  if (!ab2_check(cudaMemcpy(Wl, l.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit plm")) return -1;
  if (!ab2_check(cudaMemcpy(Wr, r.p, tot*sizeof(double), cudaMemcpyDeviceToHost), "unit plm")) return -1;
  return 0;
}

} // extern "C"
