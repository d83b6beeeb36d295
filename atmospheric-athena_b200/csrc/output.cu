// output.cu -- the device side of the driver's writers: what a writer needs of the state is reduced ON THE
// DEVICE to the bytes it puts in the file, so a dump moves those bytes over PCIe instead of the whole 48 B/cell AoS mirror.
//   dump_vtk.c:164-312    active zones as float32, big-endian, field by field (density | 3-vector | energy/pressure | scalar),
//                         conserved or primitive (Cons_to_Prim, convert_var.c:97-160 = Cons1D_to_Prim1D :389-421)
//   restart.c:548-760     active zones of one variable as FP64 in k-j-i order (the DENSITY / n-MOMENTUM / ENERGY / SCALAR blocks)
// The driver-side halves (file names, headers, fwrite) stay in C: shim/athena_shim.c __wrap_dump_vtk.
// dump_history.c keeps reading the host mirror: its volume sums are sequential over (k,j,i) and the columns that cancel
// to ~0 depend on that order in every printed digit.
#include <stdint.h>
#include <string.h>
#include "../../include/athena_b200.h"
#include "engine.cuh"
#include "physics.cuh"

#define NG AB2_NG

__device__ __forceinline__ float f32_big_endian(double x)
{ // (float)x as the reference's cast rounds (to nearest even), then ath_bswap (utils.c) on a little-endian host
  const unsigned u = __float_as_uint(__double2float_rn(x));
  return __uint_as_float(__byte_perm(u, 0u, 0x0123));
}

// field: 0 d, 1 (M1,M2,M3) or (V1,V2,V3) interleaved, 2 E or P, 3 s or r = s/d.  Planes [k0, k0+nk) of the ACTIVE box.
template <int NS>
__global__ void __launch_bounds__(256) k_vtk_field(GridDev g, int field, int prim, int k0, int nk, float* out)
{
  const long nij = (long)g.nx[0]*g.nx[1];
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nij*nk) return;
  const int kk = (int)(t / nij);
  const long r = t - (long)kk*nij;
  const int j = (int)(r / g.nx[0]), i = (int)(r - (long)j*g.nx[0]);
  const long c = (long)(k0+kk+NG)*g.st[2] + (long)(j+NG)*g.st[1] + (i+NG);
  S6 U;
  U.d = g.U[0][c]; U.x = g.U[1][c]; U.y = g.U[2][c]; U.z = g.U[3][c]; U.e = g.U[4][c]; U.s = NS ? g.U[5][c] : 0.0;
  if (prim) { Arith<false> ar; U = cons_to_prim<NS>(U, g.Gamma_1, ar); }
  if (field == 0) out[t] = f32_big_endian(U.d);
  else if (field == 1) { out[3*t] = f32_big_endian(U.x); out[3*t+1] = f32_big_endian(U.y); out[3*t+2] = f32_big_endian(U.z); }
  else if (field == 2) out[t] = f32_big_endian(U.e);
  else out[t] = f32_big_endian(U.s);
}

__global__ void __launch_bounds__(256) k_pack_active(GridDev g, int var, int k0, int nk, double* out)
{
  const long nij = (long)g.nx[0]*g.nx[1];
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nij*nk) return;
  const int kk = (int)(t / nij);
  const long r = t - (long)kk*nij;
  const int j = (int)(r / g.nx[0]), i = (int)(r - (long)j*g.nx[0]);
  out[t] = g.U[var][(long)(k0+kk+NG)*g.st[2] + (long)(j+NG)*g.st[1] + (i+NG)];
}

extern "C" {

// Page-lock a host array the driver owns (GridS.U, allocated by ath_array.c) so that the coherence copies run at PCIe
// speed straight into it instead of through the runtime's pageable staging.  Harmless to call twice; returns 0 even when
// the registration is refused (the copies then take the pageable path as before).
int ab2_pin_host(void* p, long long bytes)
{
  if (p == nullptr || bytes <= 0) return 0;
  if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) != cudaSuccess) cudaGetLastError();
  return 0;
}
int ab2_unpin_host(void* p)
{
  if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
  return 0;
}

// One field of dump_vtk's CELL_DATA section for the active zones, exactly the bytes the reference writes: float32,
// big-endian, k-j-i order (field 1: three components interleaved per zone).  dst: host, nx1*nx2*nx3*(field == 1 ? 3 : 1) floats.
int ab2_output_vtk_field(ab2_grid* h, int field, int prim, float* dst)
{
  const GridDev& g = h->g;
  if (field < 0 || field > 3 || (field == 3 && h->nscal == 0)) { ab2_set_error("[ab2_output_vtk_field]: field 0..3 (3 needs a passive scalar)"); return -1; }
  const int per = (field == 1) ? 3 : 1;
  const long nij = (long)g.nx[0]*g.nx[1];
  float* stage = (float*)h->scratch;
  const long cap = (long)(h->scratch_doubles*2/((size_t)nij*per));            // planes per slab
  if (cap < 1) { ab2_set_error("[ab2_output_vtk_field]: scratch too small"); return -1; }
  for (int k0 = 0; k0 < g.nx[2]; k0 += (int)cap) {
    const int nk = (k0 + cap <= g.nx[2]) ? (int)cap : g.nx[2] - k0;
    const long n = nij*nk;
    if (h->nscal) AB2_LAUNCH(h, KID_TRANSPOSE, (k_vtk_field<1><<<(unsigned)((n+255)/256), 256, 0, h->stream>>>(g, field, prim, k0, nk, stage)));
    else          AB2_LAUNCH(h, KID_TRANSPOSE, (k_vtk_field<0><<<(unsigned)((n+255)/256), 256, 0, h->stream>>>(g, field, prim, k0, nk, stage)));
    if (!ab2_check(cudaMemcpyAsync(dst + (size_t)k0*nij*per, stage, (size_t)n*per*sizeof(float), cudaMemcpyDeviceToHost, h->stream), "vtk field D2H")) return -1;
    if (!ab2_check(cudaStreamSynchronize(h->stream), "vtk field sync")) return -1;
  }
  return 0;
}

// Active zones of one conserved variable, FP64, k-j-i order: a DENSITY / n-MOMENTUM / ENERGY / SCALAR block of a restart dump.
int ab2_output_active_var(ab2_grid* h, int var, double* dst)
{
  const GridDev& g = h->g;
  if (var < 0 || var >= 5 + h->nscal) { ab2_set_error("[ab2_output_active_var]: bad variable"); return -1; }
  const long nij = (long)g.nx[0]*g.nx[1];
  const long cap = (long)(h->scratch_doubles/(size_t)nij);
  if (cap < 1) { ab2_set_error("[ab2_output_active_var]: scratch too small"); return -1; }
  for (int k0 = 0; k0 < g.nx[2]; k0 += (int)cap) {
    const int nk = (k0 + cap <= g.nx[2]) ? (int)cap : g.nx[2] - k0;
    const long n = nij*nk;
    AB2_LAUNCH(h, KID_TRANSPOSE, (k_pack_active<<<(unsigned)((n+255)/256), 256, 0, h->stream>>>(g, var, k0, nk, h->scratch)));
    if (!ab2_check(cudaMemcpyAsync(dst + (size_t)k0*nij, h->scratch, (size_t)n*sizeof(double), cudaMemcpyDeviceToHost, h->stream), "active var D2H")) return -1;
    if (!ab2_check(cudaStreamSynchronize(h->stream), "active var sync")) return -1;
  }
  return 0;
}

} // extern "C"
