// hydro.cu -- FP64 CUDA kernels (sm_100a) for the hydro half of the hot path:
//   bvals_mhd   (bvals_mhd.c:174-508, physical BCs :959-1949)
//   new_dt      (new_dt.c:69-139)
//   integrate_3d_ctu (integrators/integrate_3d_ctu.c:110-3368; lr_states: PLM lr_states_plm.c or PPM lr_states_ppm.c;
//                     fluxes: Roe roe.c/hlle.c with or without H-correction, or HLLC hllc.c -- the reference's configure-time
//                     choices, selected per Grid by ab2_set_reconstruction / ab2_set_riemann_solver)
//
// Layout: SoA, one double[n3][n2][n1] per variable, x1 fastest.  Face states and fluxes are kept in
// PHYSICAL component order (d,M1,M2,M3,E,s) so the reference's (x,y,z)->(z,x,y) permutations between
// sweeps reduce to component-wise operations; the sweep order (Mx,My,Mz) is only
// formed in registers when a pencil routine (PLM / Roe / cfast) needs it.
//
// Pass structure of one step (HBM round trips):
//   A  k_first_x1 / k_first_march<1|2> : U -> first-pass face states Ul,Ur + fluxes F     (steps 1-3)
//      (PPM: k_ppm_slope / k_ppm_states / k_ppm_flux per direction instead)
//   B  (d^{n+1/2} from first-pass fluxes, static potential only, step 8a: evaluated inside pass E)
//   CD k_cfinal                        : ONE kernel for steps 5-7 and 9: a (32 x 8) tile of cells marches along x3; per plane
//                                        the transverse flux-gradient + gravity corrections of the L/R states of all three
//                                        face directions and their eta go to a shared-memory tile (1-cell halo recomputed),
//                                        etah = max of 9 etas comes from that tile, then the Riemann flux.  Corrected states
//                                        and eta never touch HBM.  (AB2_HYDRO_FUSE=0: the unfused pair k_correct / k_final,
//                                        which writes corrected states and eta to HBM and reads them back.)
//   E  k_update                        : gravity source (11a) + three ordered flux-difference updates (12a-c)
// The scratch (Ul,Ur,F,Ff) may be chunked along x3 (ab2_alloc_scratch): passes A-E then run chunk by chunk on scratch that
// holds only the chunk's planes, reading U^n from a staged copy so that a chunk's update cannot disturb its neighbour's input.
#include <stdlib.h>
#include "engine.cuh"
#include "physics.cuh"

#define NG AB2_NG
#ifdef AB2_FORCE_ROE
#define RIEMANN_OF(g) 0
#else
#define RIEMANN_OF(g) ((g).riemann)
#endif

// ---------------------------------------------------------------------------------------------------
// sweep-order <-> physical-order register shuffles
// ---------------------------------------------------------------------------------------------------
template <int DIR, int NS> __device__ __forceinline__ S6 load_sweep(const double* const* A, long c)
{
  S6 r;
  r.d = A[0][c];
  r.x = A[1+DIR][c];
  r.y = A[1+(DIR+1)%3][c];
  r.z = A[1+(DIR+2)%3][c];
  r.e = A[4][c];
  r.s = NS ? A[5][c] : 0.0;
  return r;
}
template <int DIR, int NS> __device__ __forceinline__ void store_sweep(double* const* A, long c, const S6& r)
{
  A[0][c] = r.d;
  A[1+DIR][c] = r.x;
  A[1+(DIR+1)%3][c] = r.y;
  A[1+(DIR+2)%3][c] = r.z;
  A[4][c] = r.e;
  if (NS) A[5][c] = r.s;
}
template <int DIR, int NS> __device__ __forceinline__ S6 sweep_from_phys(const double* P)
{
  S6 r; r.d = P[0]; r.x = P[1+DIR]; r.y = P[1+(DIR+1)%3]; r.z = P[1+(DIR+2)%3]; r.e = P[4]; r.s = NS ? P[5] : 0.0;
  return r;
}

// ---------------------------------------------------------------------------------------------------
// Pass A, x1: one block = 128 consecutive cells of one (j,k) row -> 125 faces.  Primitive states and the
// left states travel between neighbouring threads through shared memory.
// ---------------------------------------------------------------------------------------------------
// tile width (threads = cells of one row) is chosen per grid so the last tile of a row is not mostly idle:
// a tile of T cells yields T-3 faces, and Nx1+4 faces per row must be covered (256 zones: 3 tiles of 96, 90 % busy,
// instead of 3 tiles of 128, 68 % busy)
// minimum resident blocks per SM (register caps), tuned on B200 at 256^3 (profiles/r1_tuning.md)
#ifndef AB2_LB_FIRST
#define AB2_LB_FIRST 2
#endif
#ifndef AB2_LB_FIRST_X1
#define AB2_LB_FIRST_X1 4
#endif
#ifndef AB2_LB_FINAL
#define AB2_LB_FINAL 5
#endif
#ifndef AB2_LB_CORRECT
#define AB2_LB_CORRECT 4
#endif
#ifndef AB2_LB_UPDATE
#define AB2_LB_UPDATE 5
#endif
#ifndef AB2_X1_THREADS_PER_SM
#define AB2_X1_THREADS_PER_SM (AB2_LB_FIRST_X1*128)
#endif
template <int NS, bool GRAV, int A1_T>
__global__ void __launch_bounds__(A1_T, (AB2_X1_THREADS_PER_SM/A1_T) > 0 ? (AB2_X1_THREADS_PER_SM/A1_T) : 1) k_first_x1(GridDev g, double dt, int klo)
{
  constexpr int A1_F = A1_T-3;
  __shared__ double sW[6][A1_T];
  __shared__ double sL[6][A1_T];
  const int t = threadIdx.x;
  const int il = NG-2, iu = g.nx[0]+NG+1;
  const int f0 = il+1 + blockIdx.x*A1_F;
  const int c = f0 - 2 + t;
  const int j = NG-2 + blockIdx.y, k = klo + blockIdx.z;
  const long row = (long)k*g.st[2] + (long)j*g.st[1];
  const double dtodx = dt/g.dx[0];

  S6 W0;
  if (c <= iu+1) {
    S6 U = load_sweep<0,NS>(g.Uin, row+c);
    AB2_GUARDED(W0 = cons_to_prim<NS>(U, g.Gamma_1, ar));
    sW[0][t] = W0.d; sW[1][t] = W0.x; sW[2][t] = W0.y; sW[3][t] = W0.z; sW[4][t] = W0.e; sW[5][t] = W0.s;
  }
  __syncthreads();
  S6 Wl_next, Wr_this;
  const bool do_plm = (t >= 1) && (t <= A1_T-2) && (c <= iu);
  if (do_plm) {
    S6 Wm, Wp;
    Wm.d = sW[0][t-1]; Wm.x = sW[1][t-1]; Wm.y = sW[2][t-1]; Wm.z = sW[3][t-1]; Wm.e = sW[4][t-1]; Wm.s = sW[5][t-1];
    Wp.d = sW[0][t+1]; Wp.x = sW[1][t+1]; Wp.y = sW[2][t+1]; Wp.z = sW[3][t+1]; Wp.e = sW[4][t+1]; Wp.s = sW[5][t+1];
    AB2_GUARDED(plm_cell<NS>(Wm, W0, Wp, dtodx, g.Gamma, Wl_next, Wr_this, ar, RIEMANN_OF(g) == AB2_RIEMANN_HLLC));
    if (GRAV) {   // integrate_3d_ctu.c:318-342
      const long cc = row + c;
      Wl_next.x -= dtodx*(g.phi[PHI_MH0][cc+1] - g.phi[PHI_M0][cc+1]);
      Wr_this.x -= dtodx*(g.phi[PHI_C][cc] - g.phi[PHI_MH0][cc]);
    }
    sL[0][t] = Wl_next.d; sL[1][t] = Wl_next.x; sL[2][t] = Wl_next.y; sL[3][t] = Wl_next.z; sL[4][t] = Wl_next.e; sL[5][t] = Wl_next.s;
  }
  __syncthreads();
  if ((t >= 2) && (t <= A1_T-2) && (c <= iu)) {
    S6 Wl;
    Wl.d = sL[0][t-1]; Wl.x = sL[1][t-1]; Wl.y = sL[2][t-1]; Wl.z = sL[3][t-1]; Wl.e = sL[4][t-1]; Wl.s = sL[5][t-1];
    S6 Ul, Ur, F;
    AB2_GUARDED(Ul = prim_to_cons<NS>(Wl, g.Gamma_1, ar); Ur = prim_to_cons<NS>(Wr_this, g.Gamma_1, ar);
                riemann_flux<NS>(RIEMANN_OF(g), Ul, Ur, Wl, Wr_this, 0.0, g.Gamma, g.Gamma_1, F, ar));
    store_sweep<0,NS>(g.Ul[0], row+c, Ul);
    store_sweep<0,NS>(g.Ur[0], row+c, Ur);
    store_sweep<0,NS>(g.F[0], row+c, F);
  }
}

// ---------------------------------------------------------------------------------------------------
// Pass A, x2/x3: threads lie along x1 (coalesced rows), each marches `seg` faces along the sweep axis with a
// 3-cell sliding window of primitives in registers; the left state of the next face is carried over.
// ---------------------------------------------------------------------------------------------------
template <int DIR, int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_FIRST) k_first_march(GridDev g, double dt, int seg, int olo, int ocount, int fl, int fu)
{
  // threads enumerate the (x1, other transverse axis) pencils in one flat index, so every warp is full;
  // olo/ocount: range [l, u] of the other transverse axis, fl/fu: faces [l+1, u] along the sweep (of the Grid or of an x3 chunk)
  const int il = NG-2, ni = g.nx[0]+4;
  constexpr int OTH = (DIR == 1) ? 2 : 1;                 // the transverse axis that is not x1
  const long tp = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (tp >= (long)ni*ocount) return;
  const int o = olo + (int)(tp / ni);
  const int i = il + (int)(tp - (long)(o-olo)*ni);
  const int fs = fl + blockIdx.y*seg;
  int fe = fs + seg; if (fe > fu+1) fe = fu+1;
  const long sd = g.st[DIR];
  const long base = (long)i + (long)o*g.st[OTH];
  const double dtodx = dt/g.dx[DIR];

  S6 Wm, W0;
  {
    const S6 Ua = load_sweep<DIR,NS>(g.Uin, base + (long)(fs-2)*sd), Ub = load_sweep<DIR,NS>(g.Uin, base + (long)(fs-1)*sd);
    AB2_GUARDED(Wm = cons_to_prim<NS>(Ua, g.Gamma_1, ar); W0 = cons_to_prim<NS>(Ub, g.Gamma_1, ar));
  }
  S6 Wl_prev;
  // Register prefetch: the conserved state of cell c+2 and the potential values of this iteration are requested at the
  // top of iteration c and consumed ~one PLM+Roe evaluation later, so the HBM/L2 latency (the kernel's dominant stall
  // in the first ncu profile: long-scoreboard waits on the first use of each freshly loaded value) overlaps the FP64 work.
  S6 Unext = load_sweep<DIR,NS>(g.Uin, base + (long)fs*sd);
  for (int c = fs-1; c < fe; c++) {
    const long cc = base + (long)c*sd;
    const S6 Ucur = Unext;                                                   // cell c+1
    if (c+1 < fe) Unext = load_sweep<DIR,NS>(g.Uin, cc + 2*sd);                // cell c+2 <= fu+1: always inside the array
    double phi_mh_p = 0.0, phi_m_p = 0.0, phi_c = 0.0, phi_mh = 0.0;
    if (GRAV) {
      phi_mh_p = g.phi[PHI_MH0+DIR][cc+sd]; phi_m_p = g.phi[PHI_M0+DIR][cc+sd];
      phi_c = g.phi[PHI_C][cc]; phi_mh = g.phi[PHI_MH0+DIR][cc];
    }
    S6 Wp, Wl_next, Wr_this;
    AB2_GUARDED(Wp = cons_to_prim<NS>(Ucur, g.Gamma_1, ar); plm_cell<NS>(Wm, W0, Wp, dtodx, g.Gamma, Wl_next, Wr_this, ar, RIEMANN_OF(g) == AB2_RIEMANN_HLLC));
    if (GRAV) {   // integrate_3d_ctu.c:635-645, 819-829
      Wl_next.x -= dtodx*(phi_mh_p - phi_m_p);
      Wr_this.x -= dtodx*(phi_c - phi_mh);
    }
    if (c >= fs) {
      S6 Ul, Ur, F;
      AB2_GUARDED(Ul = prim_to_cons<NS>(Wl_prev, g.Gamma_1, ar); Ur = prim_to_cons<NS>(Wr_this, g.Gamma_1, ar);
                  riemann_flux<NS>(RIEMANN_OF(g), Ul, Ur, Wl_prev, Wr_this, 0.0, g.Gamma, g.Gamma_1, F, ar));
      store_sweep<DIR,NS>(g.Ul[DIR], cc, Ul);
      store_sweep<DIR,NS>(g.Ur[DIR], cc, Ur);
      store_sweep<DIR,NS>(g.F[DIR], cc, F);
    }
    Wl_prev = Wl_next; Wm = W0; W0 = Wp;
  }
}

// ---------------------------------------------------------------------------------------------------
// thread -> (i,j,k) over a box [lo,hi] (inclusive); x-threads flatten (i,j), blockIdx.y is k
// ---------------------------------------------------------------------------------------------------
struct Box { int lo[3], hi[3]; };
__device__ __forceinline__ bool box_cell(const Box& b, int& i, int& j, int& k)
{
  const int ni = b.hi[0]-b.lo[0]+1, nj = b.hi[1]-b.lo[1]+1;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)ni*nj) return false;
  const int jj = (int)(t / ni);
  i = b.lo[0] + (int)(t - (long)jj*ni);
  j = b.lo[1] + jj;
  k = b.lo[2] + blockIdx.y;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// Pass A with PPM (reference built --with-order=3): three plain one-thread-per-cell kernels per sweep direction instead
// of the fused PLM kernels.  The slopes of a direction live in that direction's flux arrays until the flux kernel
// overwrites them; the face-state arrays carry PRIMITIVE states between the second and third kernel because the
// reference hands the primitive L/R states themselves to fluxes() in this pass (integrate_3d_ctu.c:248-300).
// ---------------------------------------------------------------------------------------------------
#ifndef AB2_LB_PPM_SLOPE
#define AB2_LB_PPM_SLOPE 6
#endif
#ifndef AB2_LB_PPM_STATES
#define AB2_LB_PPM_STATES 6
#endif
template <int DIR, int NS>
__global__ void __launch_bounds__(128, AB2_LB_PPM_SLOPE) k_ppm_slope(GridDev g, Box b)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  const long sd = g.st[DIR];
  const S6 Um = load_sweep<DIR,NS>(g.Uin, c-sd), U0 = load_sweep<DIR,NS>(g.Uin, c), Up = load_sweep<DIR,NS>(g.Uin, c+sd);
  double dWm[6];
  AB2_GUARDED(const S6 Wm = cons_to_prim<NS>(Um, g.Gamma_1, ar); const S6 W0 = cons_to_prim<NS>(U0, g.Gamma_1, ar);
              const S6 Wp = cons_to_prim<NS>(Up, g.Gamma_1, ar);
              const PrimEig e = prim_eig(W0, g.Gamma, ar);
              ppm_slope<NS>(Wm, W0, Wp, e, dWm, ar));
  S6 d; d.d = dWm[0]; d.x = dWm[1]; d.y = dWm[2]; d.z = dWm[3]; d.e = dWm[4]; d.s = dWm[5];
  store_sweep<DIR,NS>(g.F[DIR], c, d);
}

template <int DIR, int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_PPM_STATES) k_ppm_states(GridDev g, Box b, double dt)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  const long sd = g.st[DIR];
  const double dtodx = dt/g.dx[DIR];
  const S6 Um = load_sweep<DIR,NS>(g.Uin, c-sd), U0 = load_sweep<DIR,NS>(g.Uin, c), Up = load_sweep<DIR,NS>(g.Uin, c+sd);
  const S6 dm = load_sweep<DIR,NS>(g.F[DIR], c-sd), d0 = load_sweep<DIR,NS>(g.F[DIR], c), dp = load_sweep<DIR,NS>(g.F[DIR], c+sd);
  S6 Wl_next, Wr_this;
  AB2_GUARDED(
    const S6 Wm = cons_to_prim<NS>(Um, g.Gamma_1, ar); const S6 W0 = cons_to_prim<NS>(U0, g.Gamma_1, ar);
    const S6 Wp = cons_to_prim<NS>(Up, g.Gamma_1, ar);
    const PrimEig e = prim_eig(W0, g.Gamma, ar);
    double Fl[6]; double Fr[6];
    Fl[0] = ppm_face(W0.d, Wm.d, d0.d, dm.d, ar); Fr[0] = ppm_face(Wp.d, W0.d, dp.d, d0.d, ar);
    Fl[1] = ppm_face(W0.x, Wm.x, d0.x, dm.x, ar); Fr[1] = ppm_face(Wp.x, W0.x, dp.x, d0.x, ar);
    Fl[2] = ppm_face(W0.y, Wm.y, d0.y, dm.y, ar); Fr[2] = ppm_face(Wp.y, W0.y, dp.y, d0.y, ar);
    Fl[3] = ppm_face(W0.z, Wm.z, d0.z, dm.z, ar); Fr[3] = ppm_face(Wp.z, W0.z, dp.z, d0.z, ar);
    Fl[4] = ppm_face(W0.e, Wm.e, d0.e, dm.e, ar); Fr[4] = ppm_face(Wp.e, W0.e, dp.e, d0.e, ar);
    // the scalar entries as the reference's aliased work arrays hold them (physics.cuh, "PPM")
    Fl[5] = NS ? Fr[0] : 0.0;
    Fr[5] = NS ? ppm_face(Wp.s, W0.s, dp.s, dp.d, ar) : 0.0;
    ppm_cell<NS>(Wm, W0, Wp, Fl, Fr, e, dtodx, Wl_next, Wr_this, RIEMANN_OF(g) == AB2_RIEMANN_HLLC));
  if (GRAV) {   // integrate_3d_ctu.c:318-342, 635-645, 819-829
    Wl_next.x -= dtodx*(g.phi[PHI_MH0+DIR][c+sd] - g.phi[PHI_M0+DIR][c+sd]);
    Wr_this.x -= dtodx*(g.phi[PHI_C][c] - g.phi[PHI_MH0+DIR][c]);
  }
  store_sweep<DIR,NS>(g.Ul[DIR], c+sd, Wl_next);      // primitive for now; k_ppm_flux converts in place
  store_sweep<DIR,NS>(g.Ur[DIR], c, Wr_this);
}

// Pass A with PLM as TWO kernels per direction (AB2_FIRST_SPLIT, bit d = direction d): the same per-cell PLM and per-face
// Riemann evaluations as k_first_x1 / k_first_march, but the L/R primitive states take a round trip through HBM (the Ul/Ur
// arrays, converted in place by k_ppm_flux) so that neither kernel carries the other's live registers: more resident
// warps for +192 B per zone and direction.
#ifndef AB2_LB_PLM
#define AB2_LB_PLM 4
#endif
template <int DIR, int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_PLM) k_plm_states(GridDev g, Box b, double dt)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  const long sd = g.st[DIR];
  const double dtodx = dt/g.dx[DIR];
  const S6 Um = load_sweep<DIR,NS>(g.Uin, c-sd), U0 = load_sweep<DIR,NS>(g.Uin, c), Up = load_sweep<DIR,NS>(g.Uin, c+sd);
  S6 Wl_next, Wr_this;
  AB2_GUARDED(
    const S6 Wm = cons_to_prim<NS>(Um, g.Gamma_1, ar); const S6 W0 = cons_to_prim<NS>(U0, g.Gamma_1, ar);
    const S6 Wp = cons_to_prim<NS>(Up, g.Gamma_1, ar);
    plm_cell<NS>(Wm, W0, Wp, dtodx, g.Gamma, Wl_next, Wr_this, ar, RIEMANN_OF(g) == AB2_RIEMANN_HLLC));
  if (GRAV) {   // integrate_3d_ctu.c:318-342, 635-645, 819-829
    Wl_next.x -= dtodx*(g.phi[PHI_MH0+DIR][c+sd] - g.phi[PHI_M0+DIR][c+sd]);
    Wr_this.x -= dtodx*(g.phi[PHI_C][c] - g.phi[PHI_MH0+DIR][c]);
  }
  store_sweep<DIR,NS>(g.Ul[DIR], c+sd, Wl_next);      // primitive for now; k_ppm_flux converts in place
  store_sweep<DIR,NS>(g.Ur[DIR], c, Wr_this);
}

template <int DIR, int NS>
__global__ void __launch_bounds__(128, AB2_LB_FINAL) k_ppm_flux(GridDev g, Box b)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  const S6 Wl = load_sweep<DIR,NS>(g.Ul[DIR], c), Wr = load_sweep<DIR,NS>(g.Ur[DIR], c);
  S6 Ul, Ur, F;
  AB2_GUARDED(Ul = prim_to_cons<NS>(Wl, g.Gamma_1, ar); Ur = prim_to_cons<NS>(Wr, g.Gamma_1, ar);
              riemann_flux<NS>(RIEMANN_OF(g), Ul, Ur, Wl, Wr, 0.0, g.Gamma, g.Gamma_1, F, ar));
  store_sweep<DIR,NS>(g.Ul[DIR], c, Ul);
  store_sweep<DIR,NS>(g.Ur[DIR], c, Ur);
  store_sweep<DIR,NS>(g.F[DIR], c, F);
}

// ---------------------------------------------------------------------------------------------------
// Steps 5a-b/6a-b/7a-b + 5d/6d/7d for ONE face: the L/R states of the DIR-face of cell cR corrected with the transverse
// flux gradients (in ascending direction order, as the reference applies them) and the static-potential source terms.
// ---------------------------------------------------------------------------------------------------
template <int DIR, int NS, bool GRAV>
__device__ __forceinline__ void corrected_states(const GridDev& g, long cR, double dt, double UL[6], double UR[6])
{
  const long cL = cR - g.st[DIR];
  constexpr int T1 = (DIR == 0) ? 1 : 0;           // transverse directions in ascending order:
  constexpr int T2 = (DIR == 2) ? 1 : 2;           // the order the reference applies them in
  const double qT1 = 0.5*(dt/g.dx[T1]), qT2 = 0.5*(dt/g.dx[T2]);
  const long s1 = g.st[T1], s2 = g.st[T2];
#pragma unroll
  for (int v=0; v<5+NS; v++) { UL[v] = g.Ul[DIR][v][cR]; UR[v] = g.Ur[DIR][v][cR]; }
#pragma unroll
  for (int v=0; v<5+NS; v++) {
    UL[v] -= qT1*(g.F[T1][v][cL+s1] - g.F[T1][v][cL]);
    UR[v] -= qT1*(g.F[T1][v][cR+s1] - g.F[T1][v][cR]);
  }
#pragma unroll
  for (int v=0; v<5+NS; v++) {
    UL[v] -= qT2*(g.F[T2][v][cL+s2] - g.F[T2][v][cL]);
    UR[v] -= qT2*(g.F[T2][v][cR+s2] - g.F[T2][v][cR]);
  }
  if (GRAV) {
    // right state: cell cR (integrate_3d_ctu.c:1171-1192 and cyclic)
    double phic = g.phi[PHI_C][cR];
    double phir = g.phi[PHI_PH0+T1][cR], phil = g.phi[PHI_MH0+T1][cR];
    UR[1+T1] -= qT1*(phir-phil)*g.Uin[0][cR];
    UR[4] -= qT1*(g.F[T1][0][cR]*(phic - phil) + g.F[T1][0][cR+s1]*(phir - phic));
    phir = g.phi[PHI_PH0+T2][cR]; phil = g.phi[PHI_MH0+T2][cR];
    UR[1+T2] -= qT2*(phir-phil)*g.Uin[0][cR];
    UR[4] -= qT2*(g.F[T2][0][cR]*(phic - phil) + g.F[T2][0][cR+s2]*(phir - phic));
    // left state: cell cL, potentials as seen from cell cR shifted by -dx_DIR (:1196-1215)
    phic = g.phi[PHI_M0+DIR][cR];
    phir = g.phi[phi_mix(DIR,T1,1)][cR]; phil = g.phi[phi_mix(DIR,T1,0)][cR];
    UL[1+T1] -= qT1*(phir-phil)*g.Uin[0][cL];
    UL[4] -= qT1*(g.F[T1][0][cL]*(phic - phil) + g.F[T1][0][cL+s1]*(phir - phic));
    phir = g.phi[phi_mix(DIR,T2,1)][cR]; phil = g.phi[phi_mix(DIR,T2,0)][cR];
    UL[1+T2] -= qT2*(phir-phil)*g.Uin[0][cL];
    UL[4] -= qT2*(g.F[T2][0][cL]*(phic - phil) + g.F[T2][0][cL+s2]*(phir - phic));
  }
}

template <int DIR, int NS> __device__ __forceinline__ double eta_of(const GridDev& g, const double UL[6], const double UR[6])
{
  const S6 sl = sweep_from_phys<DIR,NS>(UL), sr = sweep_from_phys<DIR,NS>(UR);
  double eta;
  AB2_GUARDED(eta = eta_face(sl, sr, g.Gamma, g.Gamma_1, ar));
  return eta;
}

// ---------------------------------------------------------------------------------------------------
// Pass C (unfused pair): corrected L/R states back to HBM, and eta (9a).  One thread per face.
// ---------------------------------------------------------------------------------------------------
template <int DIR, int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_CORRECT) k_correct(GridDev g, Box b, double dt)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long cR = (long)k*g.st[2] + (long)j*g.st[1] + i;
  double UL[6], UR[6];
  corrected_states<DIR,NS,GRAV>(g, cR, dt, UL, UR);
#pragma unroll
  for (int v=0; v<5+NS; v++) { g.Ul[DIR][v][cR] = UL[v]; g.Ur[DIR][v][cR] = UR[v]; }
  {
    const S6 sl = sweep_from_phys<DIR,NS>(UL), sr = sweep_from_phys<DIR,NS>(UR);
    double eta;
    AB2_GUARDED(eta = eta_face(sl, sr, g.Gamma, g.Gamma_1, ar));
    g.eta[DIR][cR] = eta;
  }
}

// Pass C for the three face directions of a cell in ONE launch: every first-pass flux array is needed by the corrections of
// the two other directions and the potential tables by all three, so one thread doing its cell's three faces back to back
// finds them in L1 instead of in HBM a second and third time (784 instead of 3 x 328 compulsory bytes per zone).
template <int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_CORRECT) k_correct3(GridDev g, Box b, double dt)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;          // b = [l+1, u] in every direction; face d needs the transverse indices <= u-1
  const long cR = (long)k*g.st[2] + (long)j*g.st[1] + i;
  double UL[6], UR[6];
  if (j < b.hi[1] && k < b.hi[2]) {
    corrected_states<0,NS,GRAV>(g, cR, dt, UL, UR);
#pragma unroll
    for (int v=0; v<5+NS; v++) { g.Ul[0][v][cR] = UL[v]; g.Ur[0][v][cR] = UR[v]; }
    g.eta[0][cR] = eta_of<0,NS>(g, UL, UR);
  }
  if (i < b.hi[0] && k < b.hi[2]) {
    corrected_states<1,NS,GRAV>(g, cR, dt, UL, UR);
#pragma unroll
    for (int v=0; v<5+NS; v++) { g.Ul[1][v][cR] = UL[v]; g.Ur[1][v][cR] = UR[v]; }
    g.eta[1][cR] = eta_of<1,NS>(g, UL, UR);
  }
  if (i < b.hi[0] && j < b.hi[1]) {
    corrected_states<2,NS,GRAV>(g, cR, dt, UL, UR);
#pragma unroll
    for (int v=0; v<5+NS; v++) { g.Ul[2][v][cR] = UL[v]; g.Ur[2][v][cR] = UR[v]; }
    g.eta[2][cR] = eta_of<2,NS>(g, UL, UR);
  }
}

// ---------------------------------------------------------------------------------------------------
// Pass D (unfused pair): final fluxes with the H-correction (integrate_3d_ctu.c:2350-2437)
// ---------------------------------------------------------------------------------------------------
template <int DIR, int NS>
__global__ void __launch_bounds__(128, AB2_LB_FINAL) k_final(GridDev g, Box b)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  constexpr int A = (DIR == 0) ? 1 : 0;
  constexpr int B = (DIR == 2) ? 1 : 2;
  const long sd = g.st[DIR], sa = g.st[A], sb = g.st[B];
  // the nine loads stay unconditional so they issue together with the state loads below (behind a branch they
  // serialise: +20 % kernel time measured); a build without H-correction just discards the result
  double etah = AB2_MAX(g.eta[A][c-sd], g.eta[A][c]);
  etah = AB2_MAX(etah, g.eta[A][c-sd+sa]);
  etah = AB2_MAX(etah, g.eta[A][c+sa]);
  etah = AB2_MAX(etah, g.eta[B][c-sd]);
  etah = AB2_MAX(etah, g.eta[B][c]);
  etah = AB2_MAX(etah, g.eta[B][c-sd+sb]);
  etah = AB2_MAX(etah, g.eta[B][c+sb]);
  etah = AB2_MAX(etah, g.eta[DIR][c]);
  if (RIEMANN_OF(g) != AB2_RIEMANN_ROE) etah = 0.0;    // H-correction only where the reference compiles it in
  const S6 Ul = load_sweep<DIR,NS>(g.Ul[DIR], c);
  const S6 Ur = load_sweep<DIR,NS>(g.Ur[DIR], c);
  S6 F;
  AB2_GUARDED(const S6 Wl = cons_to_prim<NS>(Ul, g.Gamma_1, ar); const S6 Wr = cons_to_prim<NS>(Ur, g.Gamma_1, ar);
              riemann_flux<NS>(RIEMANN_OF(g), Ul, Ur, Wl, Wr, etah, g.Gamma, g.Gamma_1, F, ar));
  store_sweep<DIR,NS>(g.Ff[DIR], c, F);
}

// ---------------------------------------------------------------------------------------------------
// Passes C+D fused (steps 5-7 and 9, integrate_3d_ctu.c:978-1065, 1282-1366, 1691-1776, 2300-2437).
//
// A block owns a (CF_TX x CF_TY) tile of (x1,x2) columns, its outer ring being a halo, and marches along x3.  What makes
// the unfused pair expensive is that etah of a face is the maximum over eta of NINE faces of all three directions, so
// eta has to exist on the neighbours before any final flux can be solved; the pair therefore parks the corrected states
// (2 x nvar doubles per face and direction) and eta in HBM.  Here eta lives in a two-plane shared-memory tile instead:
//   per plane k:  eta1(k), eta2(k) of every tile cell (halo included: recomputed, not exchanged)        -> smem
//                 __syncthreads
//                 x3 flux of plane k from the x3 states kept from the previous iteration (etah needs eta1,2 of k-1 and k)
//                 corrected x3 states of plane k+1 (kept, thread-private in smem) and eta3(k+1)         -> smem
//                 __syncthreads
//                 x1 and x2 fluxes of plane k (etah needs eta3 of k and k+1); their corrected states are recomputed from
//                 the very same operands (same bits) rather than held in registers across two Riemann solves.
// HBM traffic per zone: first-pass Ul,Ur,F read once (plus halo re-reads served by L2), final fluxes written once;
// the corrected states and eta of the unfused pair (3 x (2 nvar + 1) doubles written and read back) are gone.
// Every per-face operation is the unfused pair's, in the same order, so results are bit-identical to it.
// ---------------------------------------------------------------------------------------------------
#define CF_TX 32
#define CF_TY 8
#ifndef AB2_LB_CFINAL
#define AB2_LB_CFINAL 2
#endif

template <int DIR> __device__ __forceinline__ double etah_tile(const double (*se)[2][CF_TY][CF_TX], int tx, int ty, int k)
{ // the nine faces of integrate_3d_ctu.c:2354-2363 / 2385-2394 / 2416-2425, in k_final's order (ternary MAX: order matters for NaN)
  constexpr int A = (DIR == 0) ? 1 : 0;
  constexpr int B = (DIR == 2) ? 1 : 2;
  // offset of "-1 along DIR" and "+1 along A / B" in (tx, ty, plane)
  constexpr int dxd = (DIR == 0) ? -1 : 0, dyd = (DIR == 1) ? -1 : 0, dkd = (DIR == 2) ? -1 : 0;
  constexpr int dxa = (A == 0) ? 1 : 0, dya = (A == 1) ? 1 : 0;                 // A is never 2
  constexpr int dyb = (B == 1) ? 1 : 0, dkb = (B == 2) ? 1 : 0;                 // B is never 0
#define ETA(dir, ddx, ddy, ddk) se[dir][(k+(ddk)) & 1][ty+(ddy)][tx+(ddx)]
  double etah = AB2_MAX(ETA(A, dxd, dyd, dkd), ETA(A, 0, 0, 0));
  etah = AB2_MAX(etah, ETA(A, dxd+dxa, dyd+dya, dkd));
  etah = AB2_MAX(etah, ETA(A, dxa, dya, 0));
  etah = AB2_MAX(etah, ETA(B, dxd, dyd, dkd));
  etah = AB2_MAX(etah, ETA(B, 0, 0, 0));
  etah = AB2_MAX(etah, ETA(B, dxd, dyd+dyb, dkd+dkb));
  etah = AB2_MAX(etah, ETA(B, 0, dyb, dkb));
  etah = AB2_MAX(etah, ETA(DIR, 0, 0, 0));
#undef ETA
  return etah;
}

template <int DIR, int NS> __device__ __forceinline__ void solve_store(const GridDev& g, long c, const double UL[6], const double UR[6], double etah)
{
  const S6 Ul = sweep_from_phys<DIR,NS>(UL), Ur = sweep_from_phys<DIR,NS>(UR);
  S6 F;
  AB2_GUARDED(const S6 Wl = cons_to_prim<NS>(Ul, g.Gamma_1, ar); const S6 Wr = cons_to_prim<NS>(Ur, g.Gamma_1, ar);
              riemann_flux<NS>(RIEMANN_OF(g), Ul, Ur, Wl, Wr, etah, g.Gamma, g.Gamma_1, F, ar));
  store_sweep<DIR,NS>(g.Ff[DIR], c, F);
}

// ks..ke: the x3 planes whose final fluxes this launch produces (x1/x2 faces for planes <= e3, x3 faces for all of them);
// blockIdx.z picks a segment of `kseg` planes.
template <int NS, bool GRAV>
__global__ void __launch_bounds__(CF_TX*CF_TY, AB2_LB_CFINAL) k_cfinal(GridDev g, double dt, int ks, int ke, int e3, int kseg)
{
  constexpr int NV = 5+NS;
  __shared__ double s_eta[3][2][CF_TY][CF_TX];
  __shared__ double s_st3[2*NV][CF_TX*CF_TY];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty*CF_TX + tx;
  const int e1 = g.nx[0]+NG-1, e2 = g.nx[1]+NG-1;
  const int i = NG-1 + (int)blockIdx.x*(CF_TX-2) + tx;
  const int j = NG-1 + (int)blockIdx.y*(CF_TY-2) + ty;
  const int kA = ks + (int)blockIdx.z*kseg;
  int kB = kA + kseg; if (kB > ke+1) kB = ke+1;
  const bool hcorr = (RIEMANN_OF(g) == AB2_RIEMANN_ROE);
  const bool inC = (i <= e1+1) && (j <= e2+1);                   // the box [s-1, e+1] on which eta is needed
  const bool inner = inC && tx >= 1 && tx <= CF_TX-2 && ty >= 1 && ty <= CF_TY-2;
  const bool d1 = inner && (j <= e2);                            // x1 faces [s, e1+1], active transversally
  const bool d2 = inner && (i <= e1);
  const bool d3 = inner && (i <= e1) && (j <= e2);
  const long col = (long)j*g.st[1] + i;
  double UL[6], UR[6];

  // prologue: eta1, eta2 of plane kA-1; corrected x3 states and eta3 of plane kA
  if (inC) {
    const long c = col + (long)(kA-1)*g.st[2];
    if (hcorr) {
      corrected_states<0,NS,GRAV>(g, c, dt, UL, UR); s_eta[0][(kA-1)&1][ty][tx] = eta_of<0,NS>(g, UL, UR);
      corrected_states<1,NS,GRAV>(g, c, dt, UL, UR); s_eta[1][(kA-1)&1][ty][tx] = eta_of<1,NS>(g, UL, UR);
    }
    corrected_states<2,NS,GRAV>(g, c + g.st[2], dt, UL, UR);
    if (hcorr) s_eta[2][kA&1][ty][tx] = eta_of<2,NS>(g, UL, UR);
#pragma unroll
    for (int v=0; v<NV; v++) { s_st3[v][tid] = UL[v]; s_st3[NV+v][tid] = UR[v]; }
  }
  for (int k = kA; k < kB; k++) {
    const long c = col + (long)k*g.st[2];
    if (inC && hcorr && k <= e3+1) {
      corrected_states<0,NS,GRAV>(g, c, dt, UL, UR); s_eta[0][k&1][ty][tx] = eta_of<0,NS>(g, UL, UR);
      corrected_states<1,NS,GRAV>(g, c, dt, UL, UR); s_eta[1][k&1][ty][tx] = eta_of<1,NS>(g, UL, UR);
    }
    __syncthreads();
    if (d3) {                                                     // x3 face of plane k
#pragma unroll
      for (int v=0; v<NV; v++) { UL[v] = s_st3[v][tid]; UR[v] = s_st3[NV+v][tid]; }
      const double etah = hcorr ? etah_tile<2>(s_eta, tx, ty, k) : 0.0;
      solve_store<2,NS>(g, c, UL, UR, etah);
    }
    if (inC && k+1 <= e3+1) {                                     // states and eta3 of plane k+1
      corrected_states<2,NS,GRAV>(g, c + g.st[2], dt, UL, UR);
      if (hcorr) s_eta[2][(k+1)&1][ty][tx] = eta_of<2,NS>(g, UL, UR);
#pragma unroll
      for (int v=0; v<NV; v++) { s_st3[v][tid] = UL[v]; s_st3[NV+v][tid] = UR[v]; }
    }
    __syncthreads();
    if (k <= e3) {                                                // x1 and x2 faces of plane k
      if (d1) {
        corrected_states<0,NS,GRAV>(g, c, dt, UL, UR);
        solve_store<0,NS>(g, c, UL, UR, hcorr ? etah_tile<0>(s_eta, tx, ty, k) : 0.0);
      }
      if (d2) {
        corrected_states<1,NS,GRAV>(g, c, dt, UL, UR);
        solve_store<1,NS>(g, c, UL, UR, hcorr ? etah_tile<1>(s_eta, tx, ty, k) : 0.0);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Pass E: static-potential source for the full step (11a, :2741-2782) then U -= dt/dx dF in the reference's
// three ordered passes (12a-c, :2981-3050), per active cell.
// ---------------------------------------------------------------------------------------------------
template <int NS, bool GRAV>
__global__ void __launch_bounds__(128, AB2_LB_UPDATE) k_update(GridDev g, Box b, double dt)
{
  int i,j,k; if (!box_cell(b,i,j,k)) return;
  const long c = (long)k*g.st[2] + (long)j*g.st[1] + i;
  const double dtodx[3] = { dt/g.dx[0], dt/g.dx[1], dt/g.dx[2] };
  double U[6];
#pragma unroll
  for (int v=0; v<5+NS; v++) U[v] = g.Uin[v][c];
  if (GRAV) {
    const double phic = g.phi[PHI_C][c];
    // d^{n+1/2} (step 8a, integrate_3d_ctu.c:2104-2125) from the FIRST-pass mass fluxes, which are still in F
    const double q1 = 0.5*(dt/g.dx[0]), q2 = 0.5*(dt/g.dx[1]), q3 = 0.5*(dt/g.dx[2]);
    const double dh = U[0]
      - q1*(g.F[0][0][c+1]       - g.F[0][0][c])
      - q2*(g.F[1][0][c+g.st[1]] - g.F[1][0][c])
      - q3*(g.F[2][0][c+g.st[2]] - g.F[2][0][c]);
#pragma unroll
    for (int d=0; d<3; d++) {
      const double phir = g.phi[PHI_PH0+d][c], phil = g.phi[PHI_MH0+d][c];
      U[1+d] -= dtodx[d]*(phir-phil)*dh;
      U[4] -= dtodx[d]*(g.Ff[d][0][c]*(phic - phil) + g.Ff[d][0][c+g.st[d]]*(phir - phic));
    }
  }
#pragma unroll
  for (int d=0; d<3; d++) {
#pragma unroll
    for (int v=0; v<5+NS; v++) U[v] -= dtodx[d]*(g.Ff[d][v][c+g.st[d]] - g.Ff[d][v][c]);
  }
#pragma unroll
  for (int v=0; v<5+NS; v++) g.U[v][c] = U[v];
}

// ---------------------------------------------------------------------------------------------------
// bvals_mhd: one kernel per face; threads over (ghost layer, transverse plane)
// ---------------------------------------------------------------------------------------------------
__global__ void k_bvals(GridDev g, int nvar, int dir, int side, int flag, Box tr /* ranges of the two other axes in slots of their own axis */)
{
  // thread -> (a, b, layer) ; a = fastest transverse axis
  const int A = (dir == 0) ? 1 : 0, B = (dir == 2) ? 1 : 2;
  const int na = tr.hi[A]-tr.lo[A]+1, nb = tr.hi[B]-tr.lo[B]+1;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)na*nb*NG) return;
  const int ia = (int)(t % na);
  const long r = t / na;
  const int layer = 1 + (int)(r % NG);
  const int ib = (int)(r / NG);
  const int s = NG, e = g.nx[dir]+NG-1;
  int dst, src;
  if (side == 0) {
    dst = s - layer;
    src = (flag == 2) ? s : (flag == 4) ? e-(layer-1) : s+(layer-1);
  } else {
    dst = e + layer;
    src = (flag == 2) ? e : (flag == 4) ? s+(layer-1) : e-(layer-1);
  }
  const long off = (long)(tr.lo[A]+ia)*g.st[A] + (long)(tr.lo[B]+ib)*g.st[B];
  const long cd = off + (long)dst*g.st[dir], cs = off + (long)src*g.st[dir];
  const bool refl = (flag == 1 || flag == 5);
  for (int v=0; v<nvar; v++) {
    double x = g.U[v][cs];
    if (refl && v == 1+dir) x = -x;
    g.U[v][cd] = x;
  }
}

// a problem file's boundary function in table form (api.cu ab2_set_bc_table): ghost cell <- constant, or +-(a cell of the same variable)
__global__ void k_bvals_table(GridDev g, int nvar, long n, const long* dst, const long* src, const double* coef)
{
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long d = dst[t];
  for (int v=0; v<nvar; v++) {
    const long sidx = src[(long)v*n + t];
    const double c = coef[(long)v*n + t];
    g.U[v][d] = (sidx >= 0) ? c*g.U[v][sidx] : c;
  }
}

int hydro_bvals(ab2_grid* h, const int bc[6])
{
  GridDev& g = h->g;
  const int nvar = 5 + h->nscal;
  for (int dir=0; dir<3; dir++) {
    if (g.nx[dir] <= 1) continue;
    Box tr;
    for (int a=0; a<3; a++) { tr.lo[a] = 0; tr.hi[a] = 0; }
    // x1: active j,k; x2: all i, active k; x3: all i,j   (bvals_mhd.c:969, 1087-1089, 1206-1208)
    for (int a=0; a<3; a++) {
      if (a == dir) continue;
      if (a < dir) { tr.lo[a] = 0; tr.hi[a] = g.n[a]-1; }
      else { tr.lo[a] = NG; tr.hi[a] = g.nx[a]+NG-1; }
    }
    const int A = (dir == 0) ? 1 : 0, B = (dir == 2) ? 1 : 2;
    const long nt = (long)(tr.hi[A]-tr.lo[A]+1)*(tr.hi[B]-tr.lo[B]+1)*NG;
    for (int side=0; side<2; side++) {
      const int flag = bc[2*dir+side];
      if (flag == -2) {                       // the problem file's own function for this face, in table form
        const int f = 2*dir+side;
        if (h->bc_n[f] > 0)
          AB2_LAUNCH(h, KID_BVALS, (k_bvals_table<<<(unsigned)((h->bc_n[f]+255)/256), 256, 0, h->stream>>>(g, nvar, h->bc_n[f], h->bc_dst[f], h->bc_src[f], h->bc_coef[f])));
        continue;
      }
      if (flag != 1 && flag != 2 && flag != 4 && flag != 5) continue;   // <=0: neighbouring Grid / prolongated face
      AB2_LAUNCH(h, KID_BVALS, (k_bvals<<<(unsigned)((nt+255)/256), 256, 0, h->stream>>>(g, nvar, dir, side, flag, tr)));
    }
  }
  return ab2_check(cudaGetLastError(), "bvals_mhd") ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------
// new_dt: max over active zones of |v_d| + sqrt(gamma P/d)  (new_dt.c:69-139)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_pos(unsigned long long* p, double x)
{ // MAX(cur, x) with the reference's ternary semantics: NaN never replaces; values are >= 0 so the bit pattern orders
  if (x > 0.0) atomicMax(p, (unsigned long long)__double_as_longlong(x));
}

__global__ void __launch_bounds__(256) k_max_speeds(GridDev g, unsigned long long* scal)
{
  const long nact = (long)g.nx[0]*g.nx[1]*g.nx[2];
  double m1 = 0.0, m2 = 0.0, m3 = 0.0;
  for (long t = (long)blockIdx.x*blockDim.x + threadIdx.x; t < nact; t += (long)gridDim.x*blockDim.x) {
    const int i = (int)(t % g.nx[0]);
    const long r = t / g.nx[0];
    const int j = (int)(r % g.nx[1]), k = (int)(r / g.nx[1]);
    const long c = (long)(k+NG)*g.st[2] + (long)(j+NG)*g.st[1] + (i+NG);
    const double d = g.U[0][c];
    const double di = 1.0/d;
    const double v1 = g.U[1][c]*di, v2 = g.U[2][c]*di, v3 = g.U[3][c]*di;
    const double qsq = v1*v1 + v2*v2 + v3*v3;
    const double p = AB2_MAX(g.Gamma_1*(g.U[4][c] - 0.5*d*qsq), AB2_TINY);
    const double asq = g.Gamma*p*di;
    const double a = sqrt(asq);
    m1 = AB2_MAX(m1, fabs(v1)+a);
    m2 = AB2_MAX(m2, fabs(v2)+a);
    m3 = AB2_MAX(m3, fabs(v3)+a);
  }
  for (int o=16; o>0; o>>=1) {
    double x;
    x = __shfl_xor_sync(0xffffffffu, m1, o); m1 = AB2_MAX(m1, x);
    x = __shfl_xor_sync(0xffffffffu, m2, o); m2 = AB2_MAX(m2, x);
    x = __shfl_xor_sync(0xffffffffu, m3, o); m3 = AB2_MAX(m3, x);
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_pos(&scal[SC_MAXV1], m1);
    atomic_max_pos(&scal[SC_MAXV2], m2);
    atomic_max_pos(&scal[SC_MAXV3], m3);
  }
}

int hydro_max_speeds(ab2_grid* h, double out[3])
{
  GridDev& g = h->g;
  if (!ab2_check(cudaMemsetAsync(h->d_scal + SC_MAXV1, 0, 3*sizeof(unsigned long long), h->stream), "new_dt memset")) return -1;
  const long nact = (long)g.nx[0]*g.nx[1]*g.nx[2];
  int blocks = (int)((nact + 255)/256); if (blocks > 148*8) blocks = 148*8;
  AB2_LAUNCH(h, KID_MAX_SPEEDS, (k_max_speeds<<<blocks, 256, 0, h->stream>>>(g, h->d_scal)));
  if (h->comm) {                               // new_dt.c:175-179: the maximum over the Grids of the Domain
    if (comm_read_scalars(h) != 0) return -1;
  } else {
    if (!ab2_check(cudaMemcpyAsync(h->h_scal + SC_MAXV1, h->d_scal + SC_MAXV1, 3*sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, h->stream), "new_dt D2H")) return -1;
    if (!ab2_check(cudaStreamSynchronize(h->stream), "new_dt sync")) return -1;
  }
  for (int d=0; d<3; d++) {
    long long bits = (long long)h->h_scal[SC_MAXV1+d];
    double v; memcpy(&v, &bits, sizeof v);
    out[d] = (g.nx[d] > 1) ? v : 0.0;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Userwork_in_loop replay: overwrite a static list of cells with fixed values
// ---------------------------------------------------------------------------------------------------
__global__ void k_apply_fixed(GridDev g, int nvar, long n, const long* cells, const double* vals)
{
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long c = cells[t];
  for (int v=0; v<nvar; v++) g.U[v][c] = vals[(long)v*n + t];
}

int hydro_apply_fixed(ab2_grid* h)
{
  if (h->n_fixed <= 0) return 0;
  AB2_LAUNCH(h, KID_APPLY_FIXED, (k_apply_fixed<<<(unsigned)((h->n_fixed+255)/256), 256, 0, h->stream>>>(h->g, 5+h->nscal, h->n_fixed, h->d_fixed_cells, h->d_fixed_vals)));
  return ab2_check(cudaGetLastError(), "apply_fixed") ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------
// integrate_3d_ctu driver
// ---------------------------------------------------------------------------------------------------
static inline dim3 box_grid(const Box& b, int threads)
{
  const long nij = (long)(b.hi[0]-b.lo[0]+1)*(b.hi[1]-b.lo[1]+1);
  return dim3((unsigned)((nij + threads-1)/threads), (unsigned)(b.hi[2]-b.lo[2]+1), 1);
}

// One pass of steps 1-12 over the active x3 planes [ka, kb] of the Grid, on scratch arrays addressed through `g`
// (the whole Grid: g = h->g; an x3 chunk: base pointers shifted so U's cell index addresses the chunk's scratch).
template <int NS, bool GRAV>
static int integrate_range(ab2_grid* h, const GridDev& g, double dt, int ka, int kb)
{
  cudaStream_t st = h->stream;
  const int s[3] = {NG,NG,ka}, e[3] = {g.nx[0]+NG-1, g.nx[1]+NG-1, kb};
  int l[3], u[3];
  for (int a=0; a<3; a++) { l[a] = s[a]-2; u[a] = e[a]+2; }
  // faces per marching segment: ~32, balanced so the last segment of a pencil is not a stub
  auto seg_for = [](int nf) { const int ns = (nf + 31)/32; return (nf + ns-1)/ns; };

  // Pass A
  if (g.order == 3) {
    for (int d=0; d<3; d++) {
      Box bs, bc, bf;      // slopes on [l-1,u+1], states from cells [l,u], faces [l+1,u] along d; [l,u] transversally
      for (int a=0; a<3; a++) {
        bs.lo[a] = (a == d) ? l[a]-1 : l[a]; bs.hi[a] = (a == d) ? u[a]+1 : u[a];
        bc.lo[a] = l[a];                     bc.hi[a] = u[a];
        bf.lo[a] = (a == d) ? l[a]+1 : l[a]; bf.hi[a] = u[a];
      }
      if (d == 0) {
        AB2_LAUNCH(h, KID_FIRST_X1, (k_ppm_slope<0,NS><<<box_grid(bs,128), 128, 0, st>>>(g, bs)));
        AB2_LAUNCH(h, KID_FIRST_X1, (k_ppm_states<0,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
        AB2_LAUNCH(h, KID_FIRST_X1, (k_ppm_flux<0,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf)));
      } else if (d == 1) {
        AB2_LAUNCH(h, KID_FIRST_X2, (k_ppm_slope<1,NS><<<box_grid(bs,128), 128, 0, st>>>(g, bs)));
        AB2_LAUNCH(h, KID_FIRST_X2, (k_ppm_states<1,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
        AB2_LAUNCH(h, KID_FIRST_X2, (k_ppm_flux<1,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf)));
      } else {
        AB2_LAUNCH(h, KID_FIRST_X3, (k_ppm_slope<2,NS><<<box_grid(bs,128), 128, 0, st>>>(g, bs)));
        AB2_LAUNCH(h, KID_FIRST_X3, (k_ppm_states<2,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
        AB2_LAUNCH(h, KID_FIRST_X3, (k_ppm_flux<2,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf)));
      }
    }
  } else {
    const int nf = g.nx[0]+4;
    int bestT = 128; double bestw = 1e30;
    // least padded tile width.  (136, two tiles for the 260 faces of a 256-zone row at 98 % busy lanes, was measured SLOWER
    // than three tiles of 96 at 90 %: 1.87 vs 1.62 ms -- 5 resident blocks of 96 hide latency better than 3 of 136; it stays
    // selectable through AB2_X1_TILE for experiments.)
    for (int T : {64, 96, 128, 160, 192, 256}) {
      const int nt = (nf + T-4)/(T-3);
      const double w = (double)nt*T/nf + 0.02*(256.0/T);  // padding, with a small bias against very narrow tiles
      if (w < bestw) { bestw = w; bestT = T; }
    }
    if (const char* env = getenv("AB2_X1_TILE")) { const int T = atoi(env); if (T == 64 || T == 96 || T == 128 || T == 136 || T == 160 || T == 192 || T == 256) bestT = T; }
    // per direction: fused PLM+Riemann kernel (default) or the two-kernel split (h->first_split bit d)
    auto split_dir = [&](int d) {
      Box bc, bf;      // states from cells [l,u] (all directions), faces [l+1,u] along d
      for (int a=0; a<3; a++) { bc.lo[a] = l[a]; bc.hi[a] = u[a]; bf.lo[a] = (a == d) ? l[a]+1 : l[a]; bf.hi[a] = u[a]; }
      if (d == 0) { AB2_LAUNCH(h, KID_FIRST_X1, (k_plm_states<0,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
                    AB2_LAUNCH(h, KID_FIRST_X1, (k_ppm_flux<0,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf))); }
      if (d == 1) { AB2_LAUNCH(h, KID_FIRST_X2, (k_plm_states<1,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
                    AB2_LAUNCH(h, KID_FIRST_X2, (k_ppm_flux<1,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf))); }
      if (d == 2) { AB2_LAUNCH(h, KID_FIRST_X3, (k_plm_states<2,NS,GRAV><<<box_grid(bc,128), 128, 0, st>>>(g, bc, dt)));
                    AB2_LAUNCH(h, KID_FIRST_X3, (k_ppm_flux<2,NS><<<box_grid(bf,128), 128, 0, st>>>(g, bf))); }
    };
    const dim3 grid((nf + bestT-4)/(bestT-3), u[1]-l[1]+1, u[2]-l[2]+1);
    if (h->first_split & 1) split_dir(0);
    else
    switch (bestT) {
      case 64:  AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,64><<<grid, 64, 0, st>>>(g, dt, l[2]))); break;
      case 96:  AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,96><<<grid, 96, 0, st>>>(g, dt, l[2]))); break;
      case 136: AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,136><<<grid, 136, 0, st>>>(g, dt, l[2]))); break;
      case 160: AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,160><<<grid, 160, 0, st>>>(g, dt, l[2]))); break;
      case 192: AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,192><<<grid, 192, 0, st>>>(g, dt, l[2]))); break;
      case 256: AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,256><<<grid, 256, 0, st>>>(g, dt, l[2]))); break;
      default:  AB2_LAUNCH(h, KID_FIRST_X1, (k_first_x1<NS,GRAV,128><<<grid, 128, 0, st>>>(g, dt, l[2]))); break;
    }
    const long ni = u[0]-l[0]+1;
    const int nf2 = u[1]-l[1], nf3 = u[2]-l[2];             // faces [l+1, u] along the sweep
    const int seg2 = seg_for(nf2), seg3 = seg_for(nf3);
    dim3 g2((unsigned)((ni*(u[2]-l[2]+1)+127)/128), (nf2 + seg2-1)/seg2, 1);
    if (h->first_split & 2) split_dir(1);
    else AB2_LAUNCH(h, KID_FIRST_X2, (k_first_march<1,NS,GRAV><<<g2, 128, 0, st>>>(g, dt, seg2, l[2], u[2]-l[2]+1, l[1]+1, u[1])));
    dim3 g3((unsigned)((ni*(u[1]-l[1]+1)+127)/128), (nf3 + seg3-1)/seg3, 1);
    if (h->first_split & 4) split_dir(2);
    else AB2_LAUNCH(h, KID_FIRST_X3, (k_first_march<2,NS,GRAV><<<g3, 128, 0, st>>>(g, dt, seg3, l[1], u[1]-l[1]+1, l[2]+1, u[2])));
  }
  Box act; for (int a=0; a<3; a++) { act.lo[a] = s[a]; act.hi[a] = e[a]; }
  // Pass B (d^{n+1/2}) is folded into pass E: the first-pass fluxes outlive pass D now that the final ones have their own arrays
  if (h->fuse) {
    // Passes C+D in one kernel: x3 planes [s, e+1]
    const int npl = e[2]+1 - s[2] + 1;
    const int nseg = (npl + 47)/48, kseg = (npl + nseg-1)/nseg;
    const dim3 grid((g.nx[0]+1 + CF_TX-3)/(CF_TX-2), (g.nx[1]+1 + CF_TY-3)/(CF_TY-2), (npl + kseg-1)/kseg);
    AB2_LAUNCH(h, KID_CFINAL, (k_cfinal<NS,GRAV><<<grid, dim3(CF_TX,CF_TY,1), 0, st>>>(g, dt, s[2], e[2]+1, e[2], kseg)));
  } else {
    // Pass C: faces of direction d over [l+1, u-1] transversally and [l+1, u] along d
    if (h->correct3) {
      Box b; for (int a=0; a<3; a++) { b.lo[a] = l[a]+1; b.hi[a] = u[a]; }
      AB2_LAUNCH(h, KID_CORRECT3, (k_correct3<NS,GRAV><<<box_grid(b,128), 128, 0, st>>>(g, b, dt)));
    } else
    for (int d=0; d<3; d++) {
      Box b; for (int a=0; a<3; a++) { b.lo[a] = l[a]+1; b.hi[a] = (a == d) ? u[a] : u[a]-1; }
      if (d == 0) AB2_LAUNCH(h, KID_CORRECT_X1+0, (k_correct<0,NS,GRAV><<<box_grid(b,128), 128, 0, st>>>(g, b, dt)));
      if (d == 1) AB2_LAUNCH(h, KID_CORRECT_X1+1, (k_correct<1,NS,GRAV><<<box_grid(b,128), 128, 0, st>>>(g, b, dt)));
      if (d == 2) AB2_LAUNCH(h, KID_CORRECT_X1+2, (k_correct<2,NS,GRAV><<<box_grid(b,128), 128, 0, st>>>(g, b, dt)));
    }
    // Pass D: faces [s, e+1] along d, active transversally (the +-1 transverse ring of the reference's loops feeds MHD only)
    for (int d=0; d<3; d++) {
      Box b; for (int a=0; a<3; a++) { b.lo[a] = s[a]; b.hi[a] = (a == d) ? e[a]+1 : e[a]; }
      if (d == 0) AB2_LAUNCH(h, KID_FINAL_X1+0, (k_final<0,NS><<<box_grid(b,128), 128, 0, st>>>(g, b)));
      if (d == 1) AB2_LAUNCH(h, KID_FINAL_X1+1, (k_final<1,NS><<<box_grid(b,128), 128, 0, st>>>(g, b)));
      if (d == 2) AB2_LAUNCH(h, KID_FINAL_X1+2, (k_final<2,NS><<<box_grid(b,128), 128, 0, st>>>(g, b)));
    }
  }
  // Pass E
  AB2_LAUNCH(h, KID_UPDATE, (k_update<NS,GRAV><<<box_grid(act,128), 128, 0, st>>>(g, act, dt)));
  return ab2_check(cudaGetLastError(), "integrate_3d_ctu") ? 0 : -1;
}

// The reference allocates its scratch for the whole Grid (integrate_3d_ctu.c:3374-3472).  Here the scratch may hold only
// `chunk_planes` active x3 planes (+4 either side): the step then runs chunk by chunk.  A chunk reads U^n from a staged copy
// of its planes (its lowest 4 planes come from a carry buffer saved before the chunk below was updated) and writes U^{n+1}
// straight into U, so results are those of the unchunked step bit for bit.
template <int NS, bool GRAV>
static int integrate_impl(ab2_grid* h, double dt)
{
  const GridDev& g0 = h->g;
  const int ks = NG, ke = g0.nx[2]+NG-1;
  if (h->chunk_planes <= 0 || h->chunk_planes >= g0.nx[2]) return integrate_range<NS,GRAV>(h, g0, dt, ks, ke);
  const int nvar = 5 + NS;
  const size_t P = (size_t)g0.st[2];
  const int depth = h->chunk_planes + 2*NG;
  // scratch layout (ab2_alloc_scratch): 12 nvar + 3 arrays of depth*P, then Uin (nvar arrays of depth*P), then carry (nvar x NG planes)
  double* uin = h->scratch + (size_t)(12*nvar + 3)*depth*P;
  double* carry = uin + (size_t)nvar*depth*P;
  cudaStream_t st = h->stream;
  for (int ka = ks; ka <= ke; ka += h->chunk_planes) {
    const int kb = (ka + h->chunk_planes - 1 < ke) ? ka + h->chunk_planes - 1 : ke;
    const int base = ka - NG;                               // lowest plane the chunk touches
    const size_t npl = (size_t)(kb + NG - base + 1);
    GridDev g = g0;
    const long shift = (long)base*(long)P;
    for (int d=0; d<3; d++) for (int v=0; v<nvar; v++) { g.Ul[d][v] -= shift; g.Ur[d][v] -= shift; g.F[d][v] -= shift; g.Ff[d][v] -= shift; }
    for (int d=0; d<3; d++) g.eta[d] -= shift;
    for (int v=0; v<nvar; v++) {
      double* dst = uin + (size_t)v*depth*P;
      if (!ab2_check(cudaMemcpyAsync(dst, g0.U[v] + shift, npl*P*sizeof(double), cudaMemcpyDeviceToDevice, st), "chunk stage")) return -1;
      if (ka > ks && !ab2_check(cudaMemcpyAsync(dst, carry + (size_t)v*NG*P, NG*P*sizeof(double), cudaMemcpyDeviceToDevice, st), "chunk carry in")) return -1;
      if (kb < ke && !ab2_check(cudaMemcpyAsync(carry + (size_t)v*NG*P, g0.U[v] + (size_t)(kb-NG+1)*P, NG*P*sizeof(double), cudaMemcpyDeviceToDevice, st), "chunk carry out")) return -1;
      g.Uin[v] = dst - shift;
    }
    if (integrate_range<NS,GRAV>(h, g, dt, ka, kb) != 0) return -1;
  }
  return 0;
}

int hydro_integrate(ab2_grid* h, double dt)
{
  if (h->nscal) return h->has_grav ? integrate_impl<1,true>(h, dt) : integrate_impl<1,false>(h, dt);
  return h->has_grav ? integrate_impl<0,true>(h, dt) : integrate_impl<0,false>(h, dt);
}
