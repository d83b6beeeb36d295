// ion.cu -- FP64 CUDA kernels (sm_100a) for the ion-plane radiative transfer + photoionisation sub-cycle:
//   ion_radtransfer_3d   (ionradiation/ionrad_3d.c:862-1047)
//   get_ph_rate_plane    (ionradiation/ionradplane_3d.c:88-407, dir = -1 branch :253-322)
//   compute_chem_rates / compute_therm_rates / ionization_update / floors / check_range / compute_dt_hydro
//                        (ionrad_3d.c:288-409, 414-562, 565-590, 70-157, 206-283, 593-678)
//   bvals_ionrad         (ionradiation/bvals_ionrad.c:63-146, 308-332)
//
// The reference makes nine full-grid sweeps per sub-cycle iteration; here an iteration is two kernels,
// separated by the grid-wide min that chooses the sub-step:
//   k_ion_rates  : one WARP per ray; 32 cells at a time, exp(-tau) in parallel, attenuation as the reference's
//                  sequential product (every lane applies the factors of the lanes before it, in order) with its
//                  1e-3 ray cut, then photo/chem/thermal rates and the per-cell dt limits of those same cells
//                  while they are in registers.  The first pass of a call also does the call's prologue (floors,
//                  save_energy_and_x) on the fly.
//   k_ion_update : E,s update, temperature and neutral floors, check_range count and the new hydro dt.
// From the second iteration of a call on both run over the rays k_ion_walk lists, i.e. those with 32-cell chunks whose
// inputs changed; everything else re-uses the previous iteration's results bit for bit ("Sparse sub-cycle" below).
// Small Grids take k_ion_rates<UPDATE = true> instead: update(n) and rates(n+1) in one launch, no memo.
// Arithmetic: every per-cell formula keeps the reference's operations and order; exp() and pow() are the host glibc's
// own algorithms re-implemented operation for operation (glibc_math.cuh), and the attenuation along a ray is the
// reference's sequential product.  On hosts whose libm selects the same (FMA) variants this path is bit-identical to
// the reference, sub-cycle decisions included; tests hold it to <= 1e-10 per cell after a step in any case.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include "engine.cuh"
#include "physics.cuh"
#include "glibc_math.cuh"

#define NG AB2_NG
#define MINFLUXFRAC 1.0e-3     // ionrad.h:26
#define IONFRACFLOOR 1.0e-4    // ionrad.h:31
#define CION 8.0e5             // ionrad.h:36
#define MAXCELLCOUNT 20        // ionrad.h:38
#define LARGE DBL_MAX          // ionrad.h:16
#define KB 1.38e-16            // ionrad_chemistry.c:43
#define MAXSIGNCOUNT 4         // ionrad_3d.c:286
#define DAMPFACTOR 0.5         // ionrad_3d.c:287

#define ERR_DT_CHEM_NEG 1ull

// minimum resident blocks per SM (register caps) and whether the per-cell arithmetic uses the guarded fast divisions
#ifndef AB2_LB_ION_RATES
#define AB2_LB_ION_RATES 7
#endif
#ifndef AB2_LB_ION_UPDATE
#define AB2_LB_ION_UPDATE 12
#endif
#ifndef AB2_ION_FAST_DIV
#define AB2_ION_FAST_DIV 1
#endif
#if AB2_ION_FAST_DIV
#define ION_GUARDED(...) AB2_GUARDED(__VA_ARGS__)
#else
#define ION_GUARDED(...) { Arith<false> ar; __VA_ARGS__; }
#endif

__device__ __forceinline__ void atomic_min_double(unsigned long long* addr, double val)
{ // `if (val < cur) cur = val` (NaN never wins), any sign
  unsigned long long old = *addr, assumed;
  do {
    assumed = old;
    if (!(val < __longlong_as_double((long long)assumed))) break;
    old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(val));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(unsigned long long* addr, double val)
{ // MAX(cur,val) = cur > val ? cur : val  -> replace unless cur > val ... but keep cur when val is NaN
  unsigned long long old = *addr, assumed;
  do {
    assumed = old;
    if (!(val > __longlong_as_double((long long)assumed))) break;
    old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(val));
  } while (assumed != old);
}

struct CellChem { double n_H, n_Hplus, n_e, x, ke, e_thermal, T; };
__device__ __forceinline__ bool same_bits_h(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }

// The kinetic energy keeps the plain operator: its numerator M.M is below the fast path's range (2^-120) wherever the
// gas is nearly at rest in CGS units, and those cells would all take the guarded block's second pass.
__device__ __forceinline__ double kinetic(double d, double M1, double M2, double M3)
{ return 0.5 * (M1*M1 + M2*M2 + M3*M3) / d; }

// Divisions go through Arith (physics.cuh): the same correctly rounded quotients as `/`, straight-line, and the
// reciprocal refinement of a repeated divisor (m_H, 14 m_H, k_B, d, nHdot, edot ...) is a common subexpression the
// compiler evaluates once.  Callers wrap a cell's arithmetic in AB2_GUARDED.
template <bool FAST>
__device__ __forceinline__ CellChem chem_state(const IonParDev& p, double Gamma_1, double d, double M1, double M2, double M3, double E, double s, Arith<FAST>& ar)
{ // ionrad_3d.c:82-101 == :313-330 == :438-456
  CellChem c;
  c.n_H = ar.div(s, p.m_H);
  c.n_Hplus = ar.div(d - s, p.m_H);
  c.n_e = c.n_Hplus + ar.div(d * p.alpha_C, 14.0 * p.m_H);
  c.x = ar.div(c.n_e, c.n_H + c.n_Hplus);
  c.ke = kinetic(d,M1,M2,M3);
  c.e_thermal = E - c.ke;
  const double e_sp = ar.div(c.e_thermal, d);
  c.T = ar.div(Gamma_1 * e_sp * (c.x*0.5*p.m_H+(1.0-c.x)*p.mu), p.k_B);
  return c;
}

// apply_temp_floor (ionrad_3d.c:70-136) then apply_neutral_floor (:140-157) for one cell: (E,s) -> (Eo,so)
template <bool FAST>
__device__ __forceinline__ void floors(const IonParDev& p, double Gamma_1, double d, double M1, double M2, double M3, double E, double s,
                                       double& Eo, double& so, Arith<FAST>& ar)
{
  const CellChem c = chem_state(p, Gamma_1, d, M1, M2, M3, E, s, ar);
  Eo = E; so = s;
  if (c.T < p.tfloor) {
    const double e_sp = ar.div(p.tfloor * p.k_B, (c.x*0.5*p.m_H+(1.0-c.x)*p.mu) * Gamma_1);
    Eo = c.ke + e_sp * d;
  }
  if ((c.T > p.tceil) && (p.tceil > 0)) {
    const double e_sp = ar.div(p.tceil * p.k_B, (c.x*0.5*p.m_H+(1.0-c.x)*p.mu) * Gamma_1);
    Eo = c.ke + e_sp * d;
  }
  double d_nlim = d*IONFRACFLOOR;
  d_nlim = d_nlim < p.d_nlo ? d_nlim : p.d_nlo;
  if (s < d_nlim) so = d_nlim;
  else if (s > d) so = d;
}

// The same two floors for a caller that wants the chemistry state of the floored cell as well (the first rates pass of a
// call, which does ion_radtransfer_3d's prologue on the fly): cs is the state of (Eo,so) on return.
template <bool FAST>
__device__ __forceinline__ void floors_state(const IonParDev& p, double Gamma_1, double d, double M1, double M2, double M3, double E, double s,
                                             double& Eo, double& so, CellChem& cs, Arith<FAST>& ar)
{
  cs = chem_state(p, Gamma_1, d, M1, M2, M3, E, s, ar);
  Eo = E; so = s;
  if (cs.T < p.tfloor) {
    const double e_sp = ar.div(p.tfloor * p.k_B, (cs.x*0.5*p.m_H+(1.0-cs.x)*p.mu) * Gamma_1);
    Eo = cs.ke + e_sp * d;
  }
  if ((cs.T > p.tceil) && (p.tceil > 0)) {
    const double e_sp = ar.div(p.tceil * p.k_B, (cs.x*0.5*p.m_H+(1.0-cs.x)*p.mu) * Gamma_1);
    Eo = cs.ke + e_sp * d;
  }
  double d_nlim = d*IONFRACFLOOR;
  d_nlim = d_nlim < p.d_nlo ? d_nlim : p.d_nlo;
  if (s < d_nlim) so = d_nlim;
  else if (s > d) so = d;
  if (!same_bits_h(Eo, E) || !same_bits_h(so, s)) cs = chem_state(p, Gamma_1, d, M1, M2, M3, Eo, so, ar);
}

// thread -> active cell
__device__ __forceinline__ bool active_cell(const GridDev& g, long& c)
{
  const long nij = (long)g.nx[0]*g.nx[1];
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nij) return false;
  const int j = (int)(t / g.nx[0]);
  const int i = (int)(t - (long)j*g.nx[0]);
  c = (long)(blockIdx.y+NG)*g.st[2] + (long)(j+NG)*g.st[1] + (i+NG);
  return true;
}

// ion_radtransfer_3d prologue (:896-905): floors + save_energy_and_x (:162-201)
__global__ void __launch_bounds__(128) k_ion_prep(GridDev g, IonParDev p, IonDev ion)
{
  long c; if (!active_cell(g, c)) return;
  const double d = g.U[0][c], M1 = g.U[1][c], M2 = g.U[2][c], M3 = g.U[3][c];
  const double E0 = g.U[4][c], s0 = g.U[5][c];
  double E, s, x_init;
  {                               // plain operators: this pass runs once per step at HBM speed, the guarded form only costs registers
    Arith<false> ar;
    floors(p, g.Gamma_1, d, M1, M2, M3, E0, s0, E, s, ar);
    const double n_H = ar.div(s, p.m_H);
    const double n_Hplus = ar.div(d - s, p.m_H);
    const double n_e = n_Hplus + ar.div(d * p.alpha_C, 14.0 * p.m_H);
    x_init = ar.div(n_e, n_H + n_Hplus);
  }
  g.U[4][c] = E; g.U[5][c] = s;
  ion.e_init[c] = E;
  ion.e_th_init[c] = E - kinetic(d,M1,M2,M3);
  ion.x_init[c] = x_init;
  ion.sign_state[c] = 1;          // last_sign = 0, sign_count = 0
}

// ---------------------------------------------------------------------------------------------------
// ionization_update (:565-590) + floors (:970-971) + check_range (:206-283) + compute_dt_hydro (:593-678) for ONE cell:
// (E,s) -> (Eo,so), whether the cell counts for check_range, and |v_d| + a per direction.  Pure in its inputs.
// ---------------------------------------------------------------------------------------------------
struct CellUpd { double E, s, x1, x2, x3, nHdot, edot; unsigned count; bool applied; };

template <bool FAST>
__device__ __forceinline__ CellUpd cell_update(const GridDev& g, const IonParDev& p, const IonDev& ion, long c, double dt,
                                               double d, double M1, double M2, double M3, double E, double s, Arith<FAST>& ar)
{
    CellUpd u;
    const double nHdot = ion.nHdot[c];
    double d_nlim = d*IONFRACFLOOR;
    d_nlim = d_nlim < p.d_nlo ? d_nlim : p.d_nlo;
    u.applied = (nHdot > 0) || (s > 1.0001*d_nlim);
    u.nHdot = nHdot; u.edot = 0.0;
    if (u.applied) {
      u.edot = ion.edot[c];
      E += u.edot * dt;
      s += nHdot * dt * p.m_H;
    }
    // floors (apply_temp_floor :70-136, apply_neutral_floor :140-157), keeping the cell's chemistry state: when no
    // floor fires (the common case) check_range below re-evaluates the very same expressions on the same values,
    // so the seven divisions are shared; when one fires they are recomputed from the floored E,s as the reference does
    CellChem cs = chem_state(p, g.Gamma_1, d, M1, M2, M3, E, s, ar);
    bool changed = false;
    if (cs.T < p.tfloor) {
      const double e_sp = ar.div(p.tfloor * p.k_B, (cs.x*0.5*p.m_H+(1.0-cs.x)*p.mu) * g.Gamma_1);
      E = cs.ke + e_sp * d; changed = true;
    }
    if ((cs.T > p.tceil) && (p.tceil > 0)) {
      const double e_sp = ar.div(p.tceil * p.k_B, (cs.x*0.5*p.m_H+(1.0-cs.x)*p.mu) * g.Gamma_1);
      E = cs.ke + e_sp * d; changed = true;
    }
    if (s < d_nlim) { s = d_nlim; changed = true; }
    else if (s > d) { s = d; changed = true; }
    u.E = E; u.s = s;
    if (changed) cs = chem_state(p, g.Gamma_1, d, M1, M2, M3, E, s, ar);

    // check_range (ionrad_3d.c:222-264)
    u.count = 0;
    if (!(ar.div(ion.ph_rate[c], p.min_area * cs.n_H) > 2.0*CION)) {
      const double eth0 = ion.e_th_init[c], e0 = ion.e_init[c], x0 = ion.x_init[c];
      if ((ar.div(cs.e_thermal, eth0) >= 1 + p.max_de_therm_step) || (ar.div(eth0, cs.e_thermal) >= 1 + p.max_de_therm_step)) u.count = 1;
      else if ((p.max_de_step > 0) && ((ar.div(E, e0) >= 1 + p.max_de_step) || (ar.div(e0, E) >= 1 + p.max_de_step))) u.count = 1;
      else if ((p.max_dx_step > 0) && ((ar.div(cs.x, x0) >= 1 + p.max_dx_step) || (ar.div(x0, cs.x) >= 1 + p.max_dx_step))) u.count = 1;
    }
    // compute_dt_hydro (:609-663): max over cells of (|v_d|+a)/dx_d.  Division by the positive constant dx_d is
    // monotone under rounding, so max_c fl(x_c/dx) == fl((max_c x_c)/dx): keep three per-direction maxima here and
    // divide once on the host (new_dt.c:126-161 has this very form).
    const double di = ar.div(1.0, d);
    const double v1 = M1*di, v2 = M2*di, v3 = M3*di;
    const double qsq = v1*v1 + v2*v2 + v3*v3;
    const double pr = AB2_MAX(g.Gamma_1*(E - 0.5*d*qsq), AB2_TINY);
    const double asq = g.Gamma*pr*di;
    const double a = ar.sqrt_(asq);
    u.x1 = fabs(v1)+a; u.x2 = fabs(v2)+a; u.x3 = fabs(v3)+a;
    return u;
}

// ---------------------------------------------------------------------------------------------------
// compute_chem_rates (ionrad_3d.c:307-394) + compute_therm_rates (:433-550) for ONE cell with photoionisation rate ph.
// Pure in its inputs; dt_therm = LARGE when the cell does not constrain the thermal step.
// ---------------------------------------------------------------------------------------------------
struct CellRates { double nHdot, edot, dt_chem, dt_therm; int sign_state; };

template <bool FAST>
__device__ __forceinline__ CellRates cell_rates(const GridDev& g, const IonParDev& p, double d, double M1, double M2, double M3,
                                                double E, double s, double ph, int st, Arith<FAST>& ar, const CellChem* have = nullptr)
{
      CellRates r;
      // --- chemistry (ionrad_3d.c:313-391); `have`: the same state, already evaluated on these very E, s by the caller
      const CellChem cs = have ? *have : chem_state(p, g.Gamma_1, d, M1, M2, M3, E, s, ar);
      double T = cs.T;
      if (T < p.tfloor) T = p.tfloor;
      double nHdot = 2.59e-13*gl_pow(ar.div(T, 1.0e4), -0.7) * p.time_unit * cs.n_e * cs.n_Hplus - ph * cs.n_H;
      int last_sign = (st & 3) - 1, sign_count = st >> 2;
      if (nHdot < 0.0) {
        if (last_sign == 1) sign_count++;
        else if (sign_count > 0) sign_count--;
        last_sign = -1;
      } else if (nHdot > 0.0) {
        if (last_sign == -1) sign_count++;
        else if (sign_count > 0) sign_count--;
        last_sign = 1;
      } else {
        sign_count = last_sign = 0;
      }
      r.sign_state = (sign_count << 2) | (last_sign + 1);
      for (int n=MAXSIGNCOUNT; n<sign_count; n++) nHdot *= DAMPFACTOR;
      r.nHdot = nHdot;

      double d_nlim = d*IONFRACFLOOR;
      d_nlim = d_nlim < p.d_nlo ? d_nlim : p.d_nlo;
      double dt1, dt2;
      if (nHdot == 0.0) {
        dt1 = dt2 = LARGE;
      } else if (nHdot > 0.0) {
        dt1 = ar.div(ar.div(p.max_dx_iter, 1+p.max_dx_iter) * cs.n_e, nHdot);
        dt2 = ar.div(p.max_dx_iter * cs.n_H, nHdot);
      } else if (s > 1.0001*d_nlim) {
        dt1 = ar.div(-p.max_dx_iter * cs.n_e, nHdot);
        dt2 = ar.div(ar.div(-p.max_dx_iter, 1+p.max_dx_iter) * cs.n_H, nHdot);
      } else {
        dt1 = dt2 = LARGE;
      }
      r.dt_chem = (dt1 < dt2) ? dt1 : dt2;

      // --- thermal (ionrad_3d.c:438-547), unfloored T
      double edot;
      bool constrain = true;
      if (cs.T < p.tfloor) { edot = 0.0; constrain = false; }
      else if ((nHdot < 0) && (s < 1.0001*d_nlim)) { edot = 0.0; constrain = false; }
      else {
        const double rc = (cs.T < 100.0) ? 0.0 : 6.11e-10*gl_pow(cs.T,-0.89)*KB*cs.T;
        const double lya = -7.5e-19*cs.n_e*cs.n_H*gl_exp(ar.div(-118348.0, cs.T));
        edot = ph * p.e_gamma * cs.n_H
          - rc * p.time_unit * cs.n_Hplus * cs.n_e
          + lya * p.time_unit;
        if (edot == 0.0) {
          dt1 = dt2 = LARGE;
        } else if (edot > 0.0) {
          dt1 = ar.div(p.max_de_iter * E, edot);
          dt2 = ar.div(p.max_de_therm_iter * cs.e_thermal, edot);
        } else {
          const double e_sp_min = ar.div(p.tfloor * p.k_B, (cs.x*0.5*p.m_H+(1.0-cs.x)*p.mu) * g.Gamma_1);
          const double e_th_min = e_sp_min * d;
          const double e_min = cs.ke + e_th_min;
          if ((ar.div(cs.e_thermal, 1.0+p.max_de_therm_iter) < e_th_min) && (ar.div(E, 1.0+p.max_de_iter) < e_min)) constrain = false;
          else {
            dt1 = ar.div(ar.div(-p.max_de_iter, 1+p.max_de_iter) * E, edot);
            dt2 = ar.div(ar.div(-p.max_de_therm_iter, 1+p.max_de_therm_iter) * cs.e_thermal, edot);
          }
        }
      }
      r.edot = edot;
      r.dt_therm = constrain ? ((dt1 < dt2) ? dt1 : dt2) : LARGE;
      return r;
}

// ---------------------------------------------------------------------------------------------------
// Sparse sub-cycle (engine.cuh IonChunks).  ion_radtransfer_3d re-evaluates every cell in every iteration; most cells of
// an irradiated-planet box are ambient gas that is already ionised down to the neutral floor, where ionization_update's
// own condition (ionrad_3d.c:575 `nHdot > 0 || s > 1.0001 d_nlim`) leaves E and s alone.  A 32-cell chunk of a ray
//   - whose sign states sit at their fixed point (count 0, last_sign = sign(nHdot))              [CH_RFIX, k_ion_rates]
//   - whose cells the update left bitwise untouched and would for any sub-step still to come in this call: the
//     condition false, or the increments rounding away even at the longest possible sub-step     [CH_UINERT, k_ion_update]
//   - which the ray enters as it did when the chunk was last evaluated (same cut state, same flux bits)
// gets bit-identical rates, dt limits, outgoing flux, check_range count and wave-speed maxima from a re-evaluation.
// So from the second iteration of a call on
//   k_ion_walk   reads the chunk records ray by ray, finds the first chunk of each ray that has to be evaluated, deposits
//                the recorded reductions of the chunks before it, and lists the rays that have such a chunk;
//   k_ion_rates / k_ion_update then run over that list only, each ray from its first such chunk on, applying the same
//                test to the chunks further down (a chunk the rates pass evaluates carries the pass number as a stamp,
//                which is how the update behind it tells evaluated from re-used).
// Nothing is approximated: a chunk is re-used only while every input of the reference's arithmetic is bitwise unchanged.
// The pair k_ion_rates<false> / k_ion_update carries the memo; the fused k_ion_rates<true> (small Grids, where an
// iteration is launch-bound) evaluates everything.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ordered_bits(double x)
{ // non-negative and not NaN: the bit pattern orders like the value
  return (unsigned long long)__double_as_longlong(x) <= 0x7ff0000000000000ull;
}
__device__ __forceinline__ double warp_min_ordered(double x)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
__device__ __forceinline__ double warp_max_ordered(double x)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
__device__ __forceinline__ bool same_bits(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
__device__ __forceinline__ double scal_as_double(const unsigned long long* scal, int i) { return __longlong_as_double((long long)scal[i]); }

// ---------------------------------------------------------------------------------------------------
// k_ion_rates: ray trace (ionradplane_3d.c:254-310) + chem rates (ionrad_3d.c:307-394) + therm rates (:433-550)
// ---------------------------------------------------------------------------------------------------
// UPDATE = true fuses the previous iteration's k_ion_update (same arithmetic, same order) in front of the rates of
// the next one: the cells are updated, floored and range-checked while they stream through for the ray trace, which
// saves one pass over U per sub-cycle iteration and lets the host read all six loop scalars in ONE synchronisation.
// The new EdgeFlux goes to ef_out (the alternate buffer) because the iteration may turn out to be the last one,
// in which case the reference's EdgeFlux is the one written before this update.
// tasks != NULL (sparse pass of the pair, UPDATE = false): warp w evaluates ray tasks[w].ray from chunk tasks[w].ch0 on, entered with
// the flux and cut state k_ion_walk found there; the chunks before it keep what is in memory.  The first warp also folds the
// dt limits k_ion_walk deposited for the re-used chunks into the pass scalars.
// MEMO: the pass writes the chunk records; TASKS: it runs over k_ion_walk's list (both compile-time, so that the full pass
// without the memo is the plain kernel).
// FIRST: the first rates pass of a call also does the call's prologue (ionrad_3d.c:896-905: temperature and neutral floors,
// save_energy_and_x, sign states reset) on the cells as they stream through, instead of a k_ion_prep pass before it.
template <bool UPDATE, bool MEMO, bool TASKS, bool FIRST = false>
__global__ void __launch_bounds__(128, AB2_LB_ION_RATES) k_ion_rates(GridDev g, IonParDev p, IonDev ion, double flux_in_root, int finegrid, double dt_upd,
                                                   const double* ef_in_base, double* ef_out_base, unsigned long long* scal,
                                                   const IonTask* tasks, int ntask, unsigned stamp)
{
  static_assert(!(UPDATE && (MEMO || TASKS)) && !(TASKS && !MEMO), "the memo belongs to the rates / update pair");
  static_assert(!(FIRST && (UPDATE || TASKS)), "the prologue rides on the first, full rates pass");
  __shared__ double s_etau[128];
  const int lane = threadIdx.x & 31;
  const long w = (long)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nrays = (long)g.nx[1]*g.nx[2];
  constexpr bool memo = MEMO;
  long ray = w;
  int c_first = 0;
  if (TASKS) {
    if (w == 0 && lane == 0) {
      atomic_min_double(&scal[SC_DT_CHEM], scal_as_double(scal, SC_W_DTC));
      atomic_min_double(&scal[SC_DT_THERM], scal_as_double(scal, SC_W_DTT));
    }
    if (w >= ntask) return;
    ray = tasks[w].ray; c_first = tasks[w].ch0*32;
  }
  if (ray >= nrays) return;
  const int jj = (int)(ray % g.nx[1]), kk = (int)(ray / g.nx[1]);
  const long row = (long)(kk+NG)*g.st[2] + (long)(jj+NG)*g.st[1] + NG;
  const long efo = ((long)kk*(g.nx[1]+1) + jj)*(g.nx[0]+1);
  double* ef = ef_out_base + efo;
  const double dx1 = g.dx[0];
  const int nx1 = g.nx[0];

  // root level: the time-ramped plane flux; finer levels: the value handed down into EdgeFlux[k][j][0] (ionradplane_3d.c:263-272)
  const double flux_in = finegrid ? ef_in_base[efo] : flux_in_root;
  unsigned long long count = 0;                       // check_range / compute_dt_hydro of the fused update
  double m1 = 0.0, m2 = 0.0, m3 = 0.0;
  double carry = flux_in;         // flux entering the current chunk
  bool alive = true;              // false once the ray has been cut
  if (TASKS) { carry = tasks[w].carry; alive = tasks[w].alive != 0; }
  double dt_chem_min = LARGE, dt_therm_min = LARGE;
  unsigned long long err = 0;

  long q = ray*ion.ch.nch + (c_first >> 5);           // the chunk's record
  for (int c0 = c_first; c0 < nx1; c0 += 32, q++) {
    const int ii = c0 + lane;
    const bool valid = ii < nx1;
    const long c = row + (valid ? ii : nx1-1);
    if (TASKS) {
      // a chunk further down a listed ray may still be re-usable (the neutral, shielded inside of the planet behind its
      // ionisation front): same test as k_ion_walk's, with the ray's actual state in hand
      const unsigned f = ion.ch.flags[q];
      if ((f & CH_RFIX) && (ion.ch.uflags[q] & CH_UINERT) && (((f & CH_ALIVE_IN) != 0) == alive)) {
        const double r_in = ion.ch.rec[q*8 + 0], r_out = ion.ch.rec[q*8 + 1], r_dtc = ion.ch.rec[q*8 + 2], r_dtt = ion.ch.rec[q*8 + 3];
        if (!alive || same_bits(carry, r_in)) {
          if (r_dtc < dt_chem_min) dt_chem_min = r_dtc;
          if (r_dtt < dt_therm_min) dt_therm_min = r_dtt;
          carry = r_out;
          alive = (f & CH_ALIVE_OUT) != 0;
          continue;                                   // flags keep their old stamp: k_ion_update re-uses the chunk too
        }
      }
    }
    unsigned fl = (stamp << CH_STAMP_SHIFT) | (alive ? CH_ALIVE_IN : 0u);     // the record's flags, rebuilt by every evaluation
    if (memo && lane == 0) ion.ch.rec[q*8 + 0] = carry;
    const double d = g.U[0][c], M1 = g.U[1][c], M2 = g.U[2][c], M3 = g.U[3][c];
    double E = g.U[4][c], s = g.U[5][c];
    CellChem cs1;                                      // FIRST: chemistry state of the floored cell, handed on to cell_rates
    if (FIRST && valid) {
      double Ef, sf;
      ION_GUARDED(floors_state(p, g.Gamma_1, d, M1, M2, M3, E, s, Ef, sf, cs1, ar));
      if (!same_bits_h(Ef, E)) g.U[4][c] = Ef;
      if (!same_bits_h(sf, s)) g.U[5][c] = sf;
      E = Ef; s = sf;
      ion.e_init[c] = E;                               // save_energy_and_x (ionrad_3d.c:162-201)
      ion.e_th_init[c] = cs1.e_thermal;
      ion.x_init[c] = cs1.x;
    }
    if (UPDATE && valid) {
      CellUpd u;
      ION_GUARDED(u = cell_update(g, p, ion, c, dt_upd, d, M1, M2, M3, E, s, ar));
      E = u.E; s = u.s;
      g.U[4][c] = E; g.U[5][c] = s;
      count += u.count;
      m1 = AB2_MAX(m1, u.x1); m2 = AB2_MAX(m2, u.x2); m3 = AB2_MAX(m3, u.x3);
    }

    // --- ray: tau, attenuation, photoionisation rate
    double n_H;
    ION_GUARDED(n_H = ar.div(s, p.m_H));
    const double tau = p.sigma_ph * n_H * dx1;
    const double etau = valid ? gl_exp(-tau) : 1.0;
    // flux entering each cell = the reference's sequential product ((F*e0)*e1)*... (ionradplane_3d.c:298 `flux *= etau`):
    // every lane applies the attenuation factors of the lanes before it IN ORDER (32 broadcast+multiply steps in
    // lockstep).  A log-step prefix scan would be 5 steps but reassociates the product, and a last-bit difference in a
    // flux is enough to flip a check_range / dt decision of the sub-cycle tens of steps later.
    double flux_before = carry;
    __syncwarp();                       // the previous chunk's reads of s_etau are done
    s_etau[threadIdx.x] = etau;         // broadcast through shared memory: one LDS per step instead of the two SHFL of a 64-bit shuffle
    __syncwarp();
    const double* se = s_etau + (threadIdx.x & ~31);
#pragma unroll
    for (int l=0; l<31; l++) {
      const double el = se[l];
      if (lane > l) flux_before *= el;
    }
    const double flux_after = flux_before*etau;
    double flux_frac, ph_lit;
    ION_GUARDED(flux_frac = ar.div(flux_after, flux_in + 1e-12); ph_lit = ar.div(flux_before * (1.0-etau), n_H*dx1));
    const unsigned cutmask = __ballot_sync(0xffffffffu, valid && (flux_frac < MINFLUXFRAC));
    const int first = cutmask ? (__ffs(cutmask) - 1) : 32;
    const bool lit = alive && (lane <= first);
    const double ph = lit ? ph_lit : 0.0;
    if (valid) {
      ef[ii] = lit ? flux_before : 0.0;
      ion.ph_rate[c] = ph;
    }
    carry = __shfl_sync(0xffffffffu, flux_after, 31);
    if (first < 32) alive = false;

    bool keep = true;                   // this cell lets the chunk's record stand in for a re-evaluation
    double cdtc = LARGE, cdtt = LARGE;
    if (valid) {
      CellRates r;
      const int st = FIRST ? 1 : ion.sign_state[c];     // 1 = last_sign 0, sign_count 0
      ION_GUARDED(r = cell_rates(g, p, d, M1, M2, M3, E, s, ph, st, ar, FIRST ? &cs1 : nullptr));
      ion.sign_state[c] = r.sign_state;
      ion.nHdot[c] = r.nHdot;
      ion.edot[c] = r.edot;
      if (r.dt_chem < dt_chem_min) dt_chem_min = r.dt_chem;
      if (r.dt_chem < 0) err |= ERR_DT_CHEM_NEG;
      if (r.dt_therm < dt_therm_min) dt_therm_min = r.dt_therm;
      if (memo) {
        // sign_count 0 after the transition = fixed point (last_sign is always the current sign); the dt limits are plain
        // non-negative numbers.  Whether the update leaves the cells alone is k_ion_update's finding (CH_UINERT).
        keep = ((r.sign_state >> 2) == 0) && ordered_bits(r.dt_chem) && ordered_bits(r.dt_therm);
        cdtc = r.dt_chem; cdtt = r.dt_therm;
      }
    }
    if (memo) {
      if (alive) fl |= CH_ALIVE_OUT;
      if (__all_sync(0xffffffffu, keep)) {
        cdtc = warp_min_ordered(cdtc); cdtt = warp_min_ordered(cdtt);
        if (lane >= 1 && lane < 4) ion.ch.rec[q*8 + lane] = (lane == 1) ? carry : (lane == 2) ? cdtc : cdtt;
        fl |= CH_RFIX;
      }
      if (lane == 0) ion.ch.flags[q] = fl;              // CH_UINERT lives in uflags: the update that follows decides
    }
  }
  if (lane == 0) ef[nx1] = alive ? carry : 0.0;      // ionradplane_3d.c:308

  for (int o=16; o>0; o>>=1) {
    double y = __shfl_xor_sync(0xffffffffu, dt_chem_min, o); if (y < dt_chem_min) dt_chem_min = y;
    y = __shfl_xor_sync(0xffffffffu, dt_therm_min, o); if (y < dt_therm_min) dt_therm_min = y;
    err |= __shfl_xor_sync(0xffffffffu, err, o);
  }
  if (UPDATE) {
    for (int o=16; o>0; o>>=1) {
      count += __shfl_xor_sync(0xffffffffu, count, o);
      double y;
      y = __shfl_xor_sync(0xffffffffu, m1, o); m1 = AB2_MAX(m1, y);
      y = __shfl_xor_sync(0xffffffffu, m2, o); m2 = AB2_MAX(m2, y);
      y = __shfl_xor_sync(0xffffffffu, m3, o); m3 = AB2_MAX(m3, y);
    }
  }
  if (lane == 0) {
    atomic_min_double(&scal[SC_DT_CHEM], dt_chem_min);
    atomic_min_double(&scal[SC_DT_THERM], dt_therm_min);
    if (err) atomicOr(&scal[SC_ERR], err);
    if (UPDATE) {
      if (count) atomicAdd(&scal[SC_COUNT], count);
      atomic_max_double(&scal[SC_MAXV1], m1);
      atomic_max_double(&scal[SC_MAXV2], m2);
      atomic_max_double(&scal[SC_MAXV3], m3);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_ion_update: ionization_update (:565-590) + floors (:970-971) + check_range (:206-283) + compute_dt_hydro (:593-678).
// One warp per 32-cell chunk of a ray, so that it addresses the chunk records of k_ion_rates.
// tasks != NULL (sparse pass): warp w takes ray tasks[w].ray and the chunks from tasks[w].ch0 on; the first warp folds in the
// count and maxima k_ion_walk deposited for the re-used chunks.
// ---------------------------------------------------------------------------------------------------
// dt_max: no later update of this call can use a longer sub-step (what is left of the hydro step / coarse time).
template <bool MEMO, bool TASKS>
__global__ void __launch_bounds__(128, AB2_LB_ION_UPDATE) k_ion_update(GridDev g, IonParDev p, IonDev ion, double dt, unsigned long long* scal,
                                                                      const IonTask* tasks, int ntask, unsigned stamp, double dt_max)
{
  static_assert(!(TASKS && !MEMO), "the task list belongs to the memo");
  const int lane = threadIdx.x & 31;
  const long w = (long)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
  // consecutive warps take consecutive chunks of a ray (of a listed ray: from its first chunk to evaluate on)
  long ray = w / ion.ch.nch;
  int ch = (int)(w - ray*ion.ch.nch);
  if (TASKS) {
    if (w == 0 && lane == 0) {
      const unsigned long long wc = scal[SC_W_CNT];
      if (wc) atomicAdd(&scal[SC_COUNT], wc);
      atomic_max_double(&scal[SC_MAXV1], scal_as_double(scal, SC_W_M1));
      atomic_max_double(&scal[SC_MAXV2], scal_as_double(scal, SC_W_M2));
      atomic_max_double(&scal[SC_MAXV3], scal_as_double(scal, SC_W_M3));
    }
    if (ray >= ntask) return;
    const IonTask t = tasks[ray];
    ray = t.ray; ch += t.ch0;
  }
  if (ray >= (long)g.nx[1]*g.nx[2] || ch >= ion.ch.nch) return;
  const int jj = (int)(ray % g.nx[1]), kk = (int)(ray / g.nx[1]);
  const int ii = ch*32 + lane;
  const bool valid = ii < g.nx[0];
  const long c = (long)(kk+NG)*g.st[2] + (long)(jj+NG)*g.st[1] + NG + (valid ? ii : g.nx[0]-1);
  if (TASKS) {
    // the rates pass of this iteration re-used the chunk (its flags still carry an older stamp, which it only does for a chunk
    // the previous update left alone): rates unchanged, cells unchanged, sub-step not longer -> this update changes nothing
    const long q = ray*ion.ch.nch + ch;
    if ((ion.ch.flags[q] >> CH_STAMP_SHIFT) != stamp) {
      if (lane == 0) {
        const unsigned long long cc = (unsigned long long)ion.ch.rec[q*8 + 7];
        if (cc) atomicAdd(&scal[SC_COUNT], cc);
        atomic_max_double(&scal[SC_MAXV1], ion.ch.rec[q*8 + 4]);
        atomic_max_double(&scal[SC_MAXV2], ion.ch.rec[q*8 + 5]);
        atomic_max_double(&scal[SC_MAXV3], ion.ch.rec[q*8 + 6]);
      }
      return;
    }
  }
  CellUpd u; u.x1 = u.x2 = u.x3 = 0.0; u.count = 0; u.applied = false;
  bool inert = true;
  if (valid) {
    const double d = g.U[0][c], M1 = g.U[1][c], M2 = g.U[2][c], M3 = g.U[3][c];
    const double E = g.U[4][c], s = g.U[5][c];
    ION_GUARDED(u = cell_update(g, p, ion, c, dt, d, M1, M2, M3, E, s, ar));
    g.U[4][c] = u.E; g.U[5][c] = u.s;
    // Inert: E and s came out with the bits they went in with (no floor fired), and would for ANY sub-step up to dt_max --
    // either the update's condition is false, or its increments round away even at dt_max (rounding is monotone, so they do
    // for every shorter sub-step: fl(E + fl(edot dt)) lies between E and fl(E + fl(edot dt_max)) = E; neutral, shielded gas).
    inert = same_bits(u.E, E) && same_bits(u.s, s) && ordered_bits(u.x1) && ordered_bits(u.x2) && ordered_bits(u.x3)
            && (!u.applied || (same_bits(E + u.edot * dt_max, E) && same_bits(s + u.nHdot * dt_max * p.m_H, s)));
  }
  unsigned long long count = u.count;
  double m1 = u.x1, m2 = u.x2, m3 = u.x3;           // >= 0 or NaN; a NaN never wins below (as in the running MAX from 0.0)
  if (MEMO) {
    // the chunk's record: with unchanged rates the next update of this chunk changes nothing either and contributes this
    // count and these maxima
    const long q = ray*ion.ch.nch + ch;
    inert = __all_sync(0xffffffffu, inert);
    if (inert) {
      const double c1 = warp_max_ordered(u.x1), c2 = warp_max_ordered(u.x2), c3 = warp_max_ordered(u.x3);
      const unsigned cc = __popc(__ballot_sync(0xffffffffu, u.count != 0));
      if (lane >= 4 && lane < 8) ion.ch.rec[q*8 + lane] = (lane == 4) ? c1 : (lane == 5) ? c2 : (lane == 6) ? c3 : (double)cc;
    }
    if (lane == 0) ion.ch.uflags[q] = inert ? CH_UINERT : 0u;          // its own array: a plain store, no read-modify-write of the rates' flags
  }
  m1 = AB2_MAX(0.0, m1); m2 = AB2_MAX(0.0, m2); m3 = AB2_MAX(0.0, m3);
  for (int o=16; o>0; o>>=1) {
    count += __shfl_xor_sync(0xffffffffu, count, o);
    double y;
    y = __shfl_xor_sync(0xffffffffu, m1, o); m1 = AB2_MAX(m1, y);
    y = __shfl_xor_sync(0xffffffffu, m2, o); m2 = AB2_MAX(m2, y);
    y = __shfl_xor_sync(0xffffffffu, m3, o); m3 = AB2_MAX(m3, y);
  }
  if (lane == 0) {
    if (count) atomicAdd(&scal[SC_COUNT], count);
    atomic_max_double(&scal[SC_MAXV1], m1);
    atomic_max_double(&scal[SC_MAXV2], m2);
    atomic_max_double(&scal[SC_MAXV3], m3);
  }
}

// ---------------------------------------------------------------------------------------------------
// k_ion_walk: one warp per ray, one lane per chunk record.  Chunk ch is re-usable when its record says so (CH_RFIX: rates
// evaluated at the sign states' fixed point; CH_UINERT: the update after them touched nothing) AND it would be entered as
// it was when the record was made: same cut state, same flux bits -- where "entered" is what the record of chunk ch-1
// says it hands on, which holds by induction while every chunk before is re-usable too.  The first chunk that fails
// starts the ray's task; the chunks before it deposit their recorded dt limits (rates pass) and count / wave-speed
// maxima (update pass) in the SC_W_* scalars.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ion_walk(GridDev g, IonDev ion, double flux_in_root, int finegrid, const double* ef_in_base,
                                                  unsigned long long* scal)
{ // persistent: every warp walks rays w, w + (warps in the grid), ...; one set of deposits per block at the end
  __shared__ double s_red[5][8];
  __shared__ unsigned s_cnt[2][8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long nrays = (long)g.nx[1]*g.nx[2];
  const long nwarps = (long)gridDim.x*(blockDim.x >> 5);
  const int nch = ion.ch.nch;
  double mdtc = LARGE, mdtt = LARGE, mm1 = 0.0, mm2 = 0.0, mm3 = 0.0;    // lane-level, over all rays of this warp
  unsigned mcnt = 0, nskip = 0;
  for (long ray = (long)blockIdx.x*(blockDim.x >> 5) + wib; ray < nrays; ray += nwarps) {
    const int jj = (int)(ray % g.nx[1]), kk = (int)(ray / g.nx[1]);
    const long efo = ((long)kk*(g.nx[1]+1) + jj)*(g.nx[0]+1);
    double prev_out = finegrid ? ef_in_base[efo] : flux_in_root;   // what the chunk before hands on
    int prev_alive = 1;
    for (int base = 0; base < nch; base += 32) {
      const int ch = base + lane;
      const bool valid = ch < nch;
      const long q = ray*nch + (valid ? ch : nch-1);
      const unsigned f = valid ? (ion.ch.flags[q] | ion.ch.uflags[q]) : 0u;
      const double r_in = ion.ch.rec[q*8 + 0], r_out = ion.ch.rec[q*8 + 1];
      double po = __shfl_up_sync(0xffffffffu, r_out, 1);
      int pa = __shfl_up_sync(0xffffffffu, (int)((f & CH_ALIVE_OUT) != 0), 1);
      if (lane == 0) { po = prev_out; pa = prev_alive; }
      const bool ok = (f & CH_RFIX) && (f & CH_UINERT) && (((f & CH_ALIVE_IN) != 0) == (pa != 0)) && (!pa || same_bits(r_in, po));
      const unsigned bad = __ballot_sync(0xffffffffu, valid && !ok);
      const int first = bad ? (__ffs(bad) - 1) : 32;
      if (valid && lane < first) {
        const double a = ion.ch.rec[q*8 + 2], b = ion.ch.rec[q*8 + 3];
        const double c1 = ion.ch.rec[q*8 + 4], c2 = ion.ch.rec[q*8 + 5], c3 = ion.ch.rec[q*8 + 6], cc = ion.ch.rec[q*8 + 7];
        if (a < mdtc) mdtc = a;
        if (b < mdtt) mdtt = b;
        mm1 = AB2_MAX(mm1, c1); mm2 = AB2_MAX(mm2, c2); mm3 = AB2_MAX(mm3, c3);
        mcnt += (unsigned)cc;
        nskip++;
      }
      if (bad) {
        const double tcarry = __shfl_sync(0xffffffffu, po, first);
        const int talive = __shfl_sync(0xffffffffu, pa, first);
        if (lane == 0) {
          IonTask t; t.ray = (int)ray; t.ch0 = base + first; t.carry = tcarry; t.alive = talive; t.pad = 0;
          ion.ch.tasks[atomicAdd(&scal[SC_NTASK], 1ull)] = t;
        }
        break;
      }
      prev_out = __shfl_sync(0xffffffffu, r_out, 31);
      prev_alive = __shfl_sync(0xffffffffu, (int)((f & CH_ALIVE_OUT) != 0), 31);
    }
  }
  // recorded values are non-negative, never NaN (checked when recorded): bit-pattern reductions
  mdtc = warp_min_ordered(mdtc); mdtt = warp_min_ordered(mdtt);
  mm1 = warp_max_ordered(mm1); mm2 = warp_max_ordered(mm2); mm3 = warp_max_ordered(mm3);
  mcnt = __reduce_add_sync(0xffffffffu, mcnt); nskip = __reduce_add_sync(0xffffffffu, nskip);
  if (lane == 0) { s_red[0][wib] = mdtc; s_red[1][wib] = mdtt; s_red[2][wib] = mm1; s_red[3][wib] = mm2; s_red[4][wib] = mm3; s_cnt[0][wib] = mcnt; s_cnt[1][wib] = nskip; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i=1; i<8; i++) {
      if (s_red[0][i] < mdtc) mdtc = s_red[0][i];
      if (s_red[1][i] < mdtt) mdtt = s_red[1][i];
      mm1 = AB2_MAX(mm1, s_red[2][i]); mm2 = AB2_MAX(mm2, s_red[3][i]); mm3 = AB2_MAX(mm3, s_red[4][i]);
      mcnt += s_cnt[0][i]; nskip += s_cnt[1][i];
    }
    if (nskip) {
      atomic_min_double(&scal[SC_W_DTC], mdtc);
      atomic_min_double(&scal[SC_W_DTT], mdtt);
      atomic_max_double(&scal[SC_W_M1], mm1);
      atomic_max_double(&scal[SC_W_M2], mm2);
      atomic_max_double(&scal[SC_W_M3], mm3);
      if (mcnt) atomicAdd(&scal[SC_W_CNT], (unsigned long long)mcnt);
      atomicAdd(&scal[SC_W_NSKIP], (unsigned long long)nskip);
    }
  }
}

__global__ void k_bvals_ionrad(IonDev ion, int nx1, int nx2, int nx3, double flux_i)
{ // bvals_ionrad.c:308-332 outflow_flux_ix1 (both EdgeFlux buffers: the edge rows j=Nx2, k=Nx3 are never rewritten)
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)(nx2+1)*(nx3+1)) return;
  ion.edgeflux[t*(nx1+1)] = flux_i;
  ion.edgeflux_alt[t*(nx1+1)] = flux_i;
}

int ion_memo_count(ab2_grid* h, long long out[3])
{ // what the latest k_ion_walk of this Grid found (it is queued behind every update but a call's last)
  out[0] = h->ion_last_nskip; out[1] = h->ion_last_ntask; out[2] = (long long)h->ion_nchunks;
  return 0;
}

int ion_bvals_ionrad(ab2_grid* h, double flux_i)
{
  const GridDev& g = h->g;
  const long n = (long)(g.nx[1]+1)*(g.nx[2]+1);
  AB2_LAUNCH(h, KID_BVALS_IONRAD, (k_bvals_ionrad<<<(unsigned)((n+255)/256), 256, 0, h->stream>>>(h->ion, g.nx[0], g.nx[1], g.nx[2], flux_i)));
  return ab2_check(cudaGetLastError(), "bvals_ionrad") ? 0 : -1;
}

static inline double bits_to_double(unsigned long long b) { double v; memcpy(&v, &b, sizeof v); return v; }
static inline unsigned long long double_to_bits(double v) { unsigned long long b; memcpy(&b, &v, sizeof b); return b; }

// ion_radtransfer_3d for a root-level Grid (ionrad_3d.c:862-1047).  flux_in = flux entering at the upstream
// face (the caller evaluates the reference's time ramp, ionradplane_3d.c:265).  reduce(vals,n,op,ctx) combines
// sub-cycle scalars across the Grids of a Domain (op 0 MIN, 1 MAX, 2 SUM): the MPI_Allreduce sites of
// ionrad_3d.c:275,399,554,672; NULL for a single Grid.
static int ion_radtransfer_impl(ab2_grid* h, int finegrid, double flux_in, double cour_no, double tcoarse, double* dt_io,
                                int* niter_out, int* nchem_out, int* ntherm_out, ab2_reduce_fn_t reduce, void* ctx);

int ion_radtransfer(ab2_grid* h, double flux_in, double cour_no, double* dt_io, int* niter_out, int* nchem_out, int* ntherm_out,
                    ab2_reduce_fn_t reduce, void* ctx)
{
  return ion_radtransfer_impl(h, 0, flux_in, cour_no, 0.0, dt_io, niter_out, nchem_out, ntherm_out, reduce, ctx);
}

// fine level (ionrad_3d.c:919-1014 with finegrid = 1): sub-cycle until the coarse time is reached; no check_range / dt_hydro stop
int ion_radtransfer_fine(ab2_grid* h, double tcoarse, double* dt_io, int* niter_out, int* nchem_out, int* ntherm_out,
                         ab2_reduce_fn_t reduce, void* ctx)
{
  return ion_radtransfer_impl(h, 1, 0.0, 0.0, tcoarse, dt_io, niter_out, nchem_out, ntherm_out, reduce, ctx);
}

static int ion_radtransfer_impl(ab2_grid* h, int finegrid, double flux_in, double cour_no, double tcoarse, double* dt_io,
                                int* niter_out, int* nchem_out, int* ntherm_out, ab2_reduce_fn_t reduce, void* ctx)
{
  GridDev& g = h->g;
  cudaStream_t st = h->stream;
  const IonParDev& p = h->ionpar;
  const long nij = (long)g.nx[0]*g.nx[1];
  const dim3 cgrid((unsigned)((nij+127)/128), (unsigned)g.nx[2], 1);
  const long nrays = (long)g.nx[1]*g.nx[2];
  const unsigned rgrid = (unsigned)((nrays + 3)/4);
  const long nch = h->ion.ch.nch;
  const unsigned ugrid = (unsigned)((nrays*nch + 3)/4);            // k_ion_update: warp per (ray, chunk)
  unsigned long long* hs = h->h_scal;

  // one host<->device round trip of the loop scalars: reset, run `launch`, read back
  auto reset_scalars = [&]() -> bool {
    hs[SC_MAXV1] = hs[SC_MAXV2] = hs[SC_MAXV3] = double_to_bits(0.0);
    hs[SC_DT_CHEM] = double_to_bits(LARGE); hs[SC_DT_THERM] = double_to_bits(LARGE);
    hs[SC_MAX_DTI] = double_to_bits(0.0); hs[SC_COUNT] = 0;
    return ab2_check(cudaMemcpyAsync(h->d_scal + SC_MAXV1, hs + SC_MAXV1, 7*sizeof(unsigned long long), cudaMemcpyHostToDevice, st), "ion scalars H2D");
  };
  // one copy + one synchronisation; with a communicator attached the scalars of all ranks are all-gathered first and
  // combined (the MPI_Allreduce sites of ionrad_3d.c:275,399,554,672 in ONE exchange, comm.cu)
  auto read_scalars = [&]() -> bool { return comm_read_scalars(h) == 0; };

  // the call's prologue (floors, save_energy_and_x): on the fly in the first rates pass, or (AB2_ION_PREP=1) as its own pass
  static const bool prep_pass = getenv("AB2_ION_PREP") && atoi(getenv("AB2_ION_PREP")) != 0;
  if (prep_pass) AB2_LAUNCH(h, KID_ION_PREP, (k_ion_prep<<<cgrid, 128, 0, st>>>(g, p, h->ion)));

  // Fusing update(n) with rates(n+1) halves the host synchronisations and cross-rank exchanges per iteration but
  // spends one discarded rates evaluation whenever check_range / dt_hydro end the loop unannounced.  Measured on B200:
  // a gain where an iteration is launch/latency-bound (64^3: 0.86 vs 1.03 ms per step), a loss from 128^3 on (2.45 vs 2.26 ms),
  // where the pair's sparse sub-cycle takes over (profiles/r2f_fuse_threshold.json).
  bool fuse = (reduce != nullptr) || ((long)g.nx[0]*g.nx[1]*g.nx[2] < (1L<<20));
  if (const char* env = getenv("AB2_ION_FUSE")) fuse = (atoi(env) != 0);
  // Sparse sub-cycle (the rates / update pair only): after the first update of a call k_ion_walk lists the rays that still
  // have chunks to evaluate; the passes of the later iterations run over that list.  The walk for iteration n+1 is queued
  // right behind update(n), so its task count arrives with the scalars of update(n) in the same synchronisation; the
  // list lives exactly as long as this call (the hydro step in between changes every cell).
  const bool memo = !fuse && h->ion.ch.flags != nullptr;
  unsigned long long* h_walk = hs + SC_W_DTC;                       // pinned: initial values of the SC_W_* block
  int ntask = -1;                                                   // -1: full pass; else rays listed for the coming passes
  auto queue_walk = [&]() -> bool {
    h_walk[0] = h_walk[1] = double_to_bits(LARGE);                                      // SC_W_DTC, SC_W_DTT
    h_walk[2] = h_walk[3] = h_walk[4] = double_to_bits(0.0);                            // SC_W_M1..3
    h_walk[5] = h_walk[6] = h_walk[7] = 0;                                              // SC_W_CNT, SC_NTASK, SC_W_NSKIP
    if (!ab2_check(cudaMemcpyAsync(h->d_scal + SC_W_DTC, h_walk, 8*sizeof(unsigned long long), cudaMemcpyHostToDevice, st), "ion walk scalars H2D")) return false;
    AB2_LAUNCH(h, KID_ION_WALK, (k_ion_walk<<<(unsigned)((nrays + 7)/8 < 148*8 ? (nrays + 7)/8 : 148*8), 256, 0, st>>>(g, h->ion, flux_in, finegrid, h->ion.edgeflux, h->d_scal)));
    return ab2_check(cudaMemcpyAsync(h_walk, h->d_scal + SC_W_DTC, 8*sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "ion walk scalars D2H");
  };
  const IonTask* tasks = h->ion.ch.tasks;

  // first iteration's rates (ph_rate_init + get_ph_rate_plane + compute_chem_rates + compute_therm_rates, :922-938)
  if (!reset_scalars()) return -1;
#define AB2_RATES1(M, F) AB2_LAUNCH(h, (F) ? KID_ION_RATES_FIRST : KID_ION_RATES, (k_ion_rates<false,M,false,F><<<rgrid, 128, 0, st>>>(g, p, h->ion, flux_in, finegrid, 0.0, h->ion.edgeflux, h->ion.edgeflux, h->d_scal, nullptr, 0, 1u)))
  if (memo) { if (prep_pass) AB2_RATES1(true, false); else AB2_RATES1(true, true); }
  else      { if (prep_pass) AB2_RATES1(false, false); else AB2_RATES1(false, true); }
#undef AB2_RATES1
  if (!read_scalars()) return -1;
  if (hs[SC_ERR] & ERR_DT_CHEM_NEG) { ab2_set_error("[compute_chem_rates]: dt_chem < 0"); return -2; }
  double v[4] = { 0.0, 0.0, bits_to_double(hs[SC_DT_CHEM]), bits_to_double(hs[SC_DT_THERM]) };
  if (reduce) reduce(v + 2, 2, 0, ctx);
  double dt_chem = v[2], dt_therm = v[3];

  double grid_dt = *dt_io, dt_done = 0.0, dt = 0.0, dt_hydro = 0.0;
  int niter = 0, hydro_done = 0, nchem = 0, ntherm = 0, coarsetime_done = 0;
  while (1) {
    if (dt_chem < dt_therm) nchem++; else ntherm++;          // ionrad_3d.c:943-945
    dt = AB2_MIN(dt_therm, dt_chem);
    if (!finegrid) { if (dt_done + dt > grid_dt) { dt = grid_dt - dt_done; hydro_done = 1; } }
    else if (dt_done + dt > tcoarse) { dt = tcoarse - dt_done; coarsetime_done = 1; }      // :958-961
    const bool last = hydro_done || coarsetime_done;          // the loop ends after this update whatever it finds

    if (!reset_scalars()) return -1;
    if (last || !fuse) {
      const unsigned stamp = (unsigned)(niter + 1);                          // of the rates pass this update follows
      const double dt_left = (finegrid ? tcoarse : grid_dt) - dt_done;       // >= every sub-step still to come in this call
      if (!memo)          AB2_LAUNCH(h, KID_ION_UPDATE, (k_ion_update<false,false><<<ugrid, 128, 0, st>>>(g, p, h->ion, dt, h->d_scal, nullptr, 0, stamp, dt_left)));
      else if (ntask < 0) AB2_LAUNCH(h, KID_ION_UPDATE, (k_ion_update<true,false><<<ugrid, 128, 0, st>>>(g, p, h->ion, dt, h->d_scal, nullptr, 0, stamp, dt_left)));
      else AB2_LAUNCH(h, KID_ION_UPDATE_MEMO, (k_ion_update<true,true><<<(unsigned)((ntask*nch + 3)/4 > 0 ? (ntask*nch + 3)/4 : 1), 128, 0, st>>>(g, p, h->ion, dt, h->d_scal, tasks, ntask, stamp, dt_left)));
      if (memo && !last && !queue_walk()) return -1;          // for the next iteration; arrives with this update's scalars
    }
    else      AB2_LAUNCH(h, KID_ION_FUSED, (k_ion_rates<true,false,false><<<rgrid, 128, 0, st>>>(g, p, h->ion, flux_in, finegrid, dt, h->ion.edgeflux, h->ion.edgeflux_alt, h->d_scal, nullptr, 0, 0u)));
    dt_done += dt;
    niter++;
    if (!read_scalars()) return -1;
    if (hs[SC_ERR] & ERR_DT_CHEM_NEG) { ab2_set_error("[compute_chem_rates]: dt_chem < 0"); return -2; }
    if (memo && !last) { ntask = (int)h_walk[6]; h->ion_last_ntask = ntask; h->ion_last_nskip = (long long)h_walk[7]; }
    double mx = 0.0;
    for (int dd=0; dd<3; dd++) if (g.nx[dd] > 1) { const double x = bits_to_double(hs[SC_MAXV1+dd])/g.dx[dd]; mx = AB2_MAX(mx, x); }
    v[0] = (double)hs[SC_COUNT]; v[1] = mx; v[2] = bits_to_double(hs[SC_DT_CHEM]); v[3] = bits_to_double(hs[SC_DT_THERM]);
    if (reduce) reduce(v, 4, 4, ctx);        // one exchange per iteration: {SUM count, MAX 1/dt_hydro, MIN dt_chem, MIN dt_therm}

    if (finegrid) { if (coarsetime_done) { grid_dt = dt_done; break; } }                   // :1008-1012
    else {
      if (v[0] > MAXCELLCOUNT) { grid_dt = dt_done; break; }   // check_range (:982)
      if (hydro_done) break;                                   // :989
      dt_hydro = cour_no/v[1];                                 // :997
      if (dt_hydro < dt_done) { grid_dt = dt_done; break; }
    }
    if (fuse) {            // continue: the rates computed behind this update become current
      std::swap(h->ion.edgeflux, h->ion.edgeflux_alt);
      dt_chem = v[2]; dt_therm = v[3];
    } else {               // separate rates pass of the next iteration
      if (!reset_scalars()) return -1;
      const unsigned stamp = (unsigned)(niter + 1);                          // rates pass of iteration niter+1
      if (!memo)          AB2_LAUNCH(h, KID_ION_RATES, (k_ion_rates<false,false,false><<<rgrid, 128, 0, st>>>(g, p, h->ion, flux_in, finegrid, 0.0, h->ion.edgeflux, h->ion.edgeflux, h->d_scal, nullptr, 0, stamp)));
      else if (ntask < 0) AB2_LAUNCH(h, KID_ION_RATES, (k_ion_rates<false,true,false><<<rgrid, 128, 0, st>>>(g, p, h->ion, flux_in, finegrid, 0.0, h->ion.edgeflux, h->ion.edgeflux, h->d_scal, nullptr, 0, stamp)));
      else AB2_LAUNCH(h, KID_ION_RATES_MEMO, (k_ion_rates<false,true,true><<<(unsigned)((ntask + 3)/4 > 0 ? (ntask + 3)/4 : 1), 128, 0, st>>>(g, p, h->ion, flux_in, finegrid, 0.0, h->ion.edgeflux, h->ion.edgeflux, h->d_scal, tasks, ntask, stamp)));
      if (!read_scalars()) return -1;
      if (hs[SC_ERR] & ERR_DT_CHEM_NEG) { ab2_set_error("[compute_chem_rates]: dt_chem < 0"); return -2; }
      v[2] = bits_to_double(hs[SC_DT_CHEM]); v[3] = bits_to_double(hs[SC_DT_THERM]);
      if (reduce) reduce(v + 2, 2, 0, ctx);
      dt_chem = v[2]; dt_therm = v[3];
    }
  }
  if (!finegrid) {
    if (niter == h->maxiter) grid_dt = dt_done;                // :1022
    h->tcoarse = dt_done;                                      // :1028
  }
  *dt_io = grid_dt;
  if (niter_out) *niter_out = niter;
  if (nchem_out) *nchem_out = nchem;
  if (ntherm_out) *ntherm_out = ntherm;
  if (!finegrid && grid_dt < 0) {                              // :1044-1046
    ab2_set_error("[ion_radtransfer_3d]: dt < 0");
    return -3;
  }
  return 0;
}
