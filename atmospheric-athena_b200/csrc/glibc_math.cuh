// glibc_math.cuh -- exp() and pow() with the bits of this image's glibc libm (x86-64, the FMA+AVX2 variants).
//
// The reference links libm dynamically; its photo-ionisation, recombination and Lyman-alpha rates are therefore
// whatever exp()/pow() of the host's glibc return (ionradplane_3d.c:295, ionrad_chemistry.c:121,147,352).  CUDA's
// math library is within 2 ulp of that but not bit-identical, and the sub-cycle is chaotic in the last bit (a
// differing ulp in one cell flips a check_range / dt decision ~70 steps later).  These two functions re-implement
// the library's algorithm operation for operation -- table look-up, the same polynomial, and a fused multiply-add
// exactly where the library's object code has one (read off the disassembly of the variants the dynamic linker
// selects on FMA+AVX2 hosts) -- so the ion path reproduces the reference bit for bit on such hosts.  Inputs outside
// the main path (non-positive or subnormal base, huge exponents, NaN) fall back to CUDA's own exp/pow: they only
// occur in states that are already invalid.  tests/test_gpu_parity.py::test_glibc_math_* compares 4 M device results
// with the host libm bit for bit.
#pragma once
#include <cuda_runtime.h>
#include "glibc_tables.h"

__device__ __forceinline__ double gl_exp_core(double x, double xtail, bool has_tail, bool& ok)
{
  // exp(x + xtail); |x| in [2^-54, 1024) handled here
  const unsigned abstop = (unsigned)((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ff;
  ok = true;
  bool special = false;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;          // tiny |x|: 1 + x (round-to-nearest branch of the library)
    if (abstop >= 0x409u) { ok = false; return 0.0; }            // |x| >= 1024: overflow / underflow handled by the caller
    special = true;                                              // 512 <= |x| < 1024: result may be subnormal
  }
  const double z = __fma_rn(x, GL_INVLN2N, GL_SHIFT);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(z);
  const double kd = z - GL_SHIFT;
  double r = __fma_rn(kd, GL_NEGLN2HIN, x);
  r = __fma_rn(kd, GL_NEGLN2LON, r);
  if (has_tail) r = xtail + r;
  const unsigned idx = 2u*(unsigned)(ki & 127u);
  const unsigned long long top = ki << 45;
  const double tail = __longlong_as_double((long long)gl_exp_tab[idx]);
  unsigned long long sbits = gl_exp_tab[idx+1] + top;
  const double p23 = __fma_rn(r, GL_C3, GL_C2);
  const double tr = r + tail;
  const double r2 = r*r;
  const double p45 = __fma_rn(r, GL_C5, GL_C4);
  double tmp = __fma_rn(p23, r2, tr);
  tmp = __fma_rn(r2*r2, p45, tmp);
  if (!special) {
    const double scale = __longlong_as_double((long long)sbits);
    return __fma_rn(scale, tmp, scale);
  }
  // specialcase(): 512 <= |x| < 1024
  if ((ki & 0x80000000ull) == 0) {                               // k > 0: scale down, then up (may overflow to inf)
    sbits -= 1009ull << 52;
    const double scale = __longlong_as_double((long long)sbits);
    return 0x1p1009 * __fma_rn(scale, tmp, scale);
  }
  sbits += 1022ull << 52;                                        // k < 0: careful rounding in the subnormal range
  const double scale = __longlong_as_double((long long)sbits);
  const double st = scale*tmp;
  double y = scale + st;
  if (fabs(y) < 1.0) {
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = scale - y + st;
    const double hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = __longlong_as_double((long long)(sbits & 0x8000000000000000ull));
  }
  return 0x1p-1022 * y;
}

// exp(x) as glibc's __exp_fma returns it
__device__ __forceinline__ double gl_exp(double x)
{
  bool ok;
  const double r = gl_exp_core(x, 0.0, false, ok);
  if (ok) return r;
  return exp(x);          // |x| >= 1024, inf, NaN: 0, inf or NaN in both libraries
}

// pow(x, y) as glibc's __pow_fma returns it, for positive normal x and 2^-65 <= |y| < 2^63
__device__ __forceinline__ double gl_pow(double x, double y)
{
  const unsigned long long ix = (unsigned long long)__double_as_longlong(x);
  const unsigned topx = (unsigned)(ix >> 52), topy = (unsigned)((unsigned long long)__double_as_longlong(y) >> 52);
  if (topx - 1u > 0x7fdu || (topy & 0x7ff) - 0x3beu > 0x7fu) return pow(x, y);      // outside the library's main path
  // log_inline
  const unsigned long long tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)((tmp >> 45) & 127u);
  const long long k = (long long)tmp >> 52;
  const unsigned long long iz = ix - (tmp & 0xfff0000000000000ull);
  const double z = __longlong_as_double((long long)iz);
  const double kd = (double)k;
  const double invc = gl_pow_log_tab[i][0], logc = gl_pow_log_tab[i][1], logctail = gl_pow_log_tab[i][2];
  const double t1 = __fma_rn(kd, GL_LN2HI, logc);
  const double lo1 = __fma_rn(kd, GL_LN2LO, logctail);
  const double r = __fma_rn(z, invc, -1.0);
  const double ar = r*GL_A0;
  const double p12 = __fma_rn(r, GL_A2, GL_A1);
  const double p34 = __fma_rn(r, GL_A4, GL_A3);
  const double t2 = r + t1;
  const double lo2 = (t1 - t2) + r;
  const double ar2 = r*ar;
  const double ar3 = r*ar2;
  const double lo3 = __fma_rn(ar, r, -ar2);
  const double hi = t2 + ar2;
  const double p56 = __fma_rn(r, GL_A6, GL_A5);
  const double lo4 = (t2 - hi) + ar2;
  const double q = __fma_rn(p56, ar2, p34);
  const double pin = __fma_rn(ar2, q, p12);
  double lo = lo1 + lo2;
  lo = lo + lo3;
  lo = lo + lo4;
  lo = __fma_rn(ar3, pin, lo);
  const double loghi = hi + lo;
  const double loglo = (hi - loghi) + lo;
  // exponent in two pieces
  const double ehi = y*loghi;
  const double e1 = __fma_rn(loghi, y, -ehi);
  const double elo = __fma_rn(y, loglo, e1);
  bool ok;
  const double res = gl_exp_core(ehi, elo, true, ok);
  if (ok) return res;
  return pow(x, y);
}
