// api.cu -- the C ABI of libathena_b200.so (include/athena_b200.h): handle lifetime, host<->device
// coherence, static-potential tabulation, and thin wrappers over hydro.cu / ion.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "../../include/athena_b200.h"
#include "engine.cuh"

static thread_local std::string g_err;
void ab2_set_error(const std::string& s) { g_err = s; }
bool ab2_check(cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return true;
  g_err = std::string("[athena_b200] ") + what + ": " + cudaGetErrorString(e);
  return false;
}
#define CK(call, what) do { if (!ab2_check((call), what)) return -1; } while (0)

static double* dev_alloc(ab2_grid* h, size_t ndouble, bool zero = true)
{
  void* p = nullptr;
  if (!ab2_check(cudaMalloc(&p, ndouble*sizeof(double)), "cudaMalloc")) return nullptr;
  if (zero && !ab2_check(cudaMemsetAsync(p, 0, ndouble*sizeof(double), h->stream), "cudaMemset")) return nullptr;
  h->allocs.push_back(p);
  h->bytes_allocated += ndouble*sizeof(double);
  return (double*)p;
}

// ---------------------------------------------------------------------------------------------------
// AoS (ConsS) <-> SoA transposition on the device
// ---------------------------------------------------------------------------------------------------
// `aos` holds cells [c0, c0+ncell) of the Grid.  Each block moves 256 cells through shared memory so that both the AoS side
// (256*nvar consecutive doubles) and the SoA side (256 consecutive doubles per variable) are written and read unit-stride;
// the plain per-thread form strides the AoS side by 48 B and reaches 46 % of the DRAM throughput.
__global__ void __launch_bounds__(256) k_aos_to_soa(GridDev g, int nvar, long c0, long ncell, const double* aos)
{
  __shared__ double tile[256*AB2_NVAR];
  const long b0 = (long)blockIdx.x*256;
  const long nc = (ncell - b0 < 256) ? ncell - b0 : 256;
  for (long t = threadIdx.x; t < nc*nvar; t += 256) tile[t] = aos[b0*nvar + t];
  __syncthreads();
  if (threadIdx.x < nc) for (int v=0; v<nvar; v++) g.U[v][c0 + b0 + threadIdx.x] = tile[threadIdx.x*nvar + v];
}
__global__ void __launch_bounds__(256) k_soa_to_aos(GridDev g, int nvar, long c0, long ncell, double* aos)
{
  __shared__ double tile[256*AB2_NVAR];
  const long b0 = (long)blockIdx.x*256;
  const long nc = (ncell - b0 < 256) ? ncell - b0 : 256;
  if (threadIdx.x < nc) for (int v=0; v<nvar; v++) tile[threadIdx.x*nvar + v] = g.U[v][c0 + b0 + threadIdx.x];
  __syncthreads();
  for (long t = threadIdx.x; t < nc*nvar; t += 256) aos[b0*nvar + t] = tile[t];
}

static const char* KERNEL_NAMES[KID_N] = {
  "k_bvals", "k_max_speeds", "k_first_x1", "k_first_march<1>", "k_first_march<2>", "k_dhalf",
  "k_correct<0>", "k_correct<1>", "k_correct<2>", "k_final<0>", "k_final<1>", "k_final<2>", "k_update",
  "k_apply_fixed", "k_ion_prep", "k_ion_rates", "k_ion_update", "k_bvals_ionrad", "k_transpose",
  "k_smr_restrict", "k_smr_flux_correct", "k_smr_prolongate", "k_ion_rates<fused update>", "k_cfinal", "k_correct3",
  "k_ion_rates[memo]", "k_ion_update[memo]", "k_ion_walk", "k_ion_rates[first]" };

cudaEvent_t ab2_prof_event(ab2_grid* h)
{
  cudaEvent_t e;
  if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}

static void prof_drain(ab2_grid* h)
{
  cudaStreamSynchronize(h->stream);
  for (auto& p : h->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { h->prof_ms[p.id] += ms; h->prof_cnt[p.id]++; }
    h->prof_pool.push_back(p.a); h->prof_pool.push_back(p.b);
  }
  h->prof_pending.clear();
}

int ab2_alloc_scratch(ab2_grid* h, int planes)
{
  GridDev& g = h->g;
  const int nvar = 5 + h->nscal;
  if (planes < 0 || (planes > 0 && planes < 4)) { ab2_set_error("[ab2_set_scratch_planes]: 0 (whole Grid) or at least 4 planes per chunk"); return -1; }
  if (planes >= g.nx[2]) planes = 0;
  if (planes > 0 && !h->own_U) { ab2_set_error("[ab2_set_scratch_planes]: chunked scratch needs library-owned U"); return -1; }
  const size_t P = (size_t)g.n[0]*g.n[1];
  const size_t depth = planes ? (size_t)(planes + 2*AB2_NG) : (size_t)g.n[2];
  size_t nd = (size_t)(12*nvar + 3)*depth*P;
  if (planes) nd += (size_t)nvar*depth*P + (size_t)nvar*AB2_NG*P;      // staged copy of the chunk's U^n, carry planes
  if (h->scratch) { cudaStreamSynchronize(h->stream); cudaFree(h->scratch); h->bytes_allocated -= h->scratch_doubles*sizeof(double); h->scratch = nullptr; }
  void* p = nullptr;
  if (!ab2_check(cudaMalloc(&p, nd*sizeof(double)), "cudaMalloc CTU scratch")) return -1;
  if (!ab2_check(cudaMemsetAsync(p, 0, nd*sizeof(double), h->stream), "cudaMemset CTU scratch")) return -1;
  h->scratch = (double*)p; h->scratch_doubles = nd; h->chunk_planes = planes;
  h->bytes_allocated += nd*sizeof(double);
  const size_t na = depth*P;
  size_t off = 0;
  for (int dir=0; dir<3; dir++) for (int v=0; v<nvar; v++) { g.Ul[dir][v] = h->scratch + off; off += na; }
  for (int dir=0; dir<3; dir++) for (int v=0; v<nvar; v++) { g.Ur[dir][v] = h->scratch + off; off += na; }
  for (int dir=0; dir<3; dir++) for (int v=0; v<nvar; v++) { g.F[dir][v] = h->scratch + off; off += na; }
  for (int dir=0; dir<3; dir++) for (int v=0; v<nvar; v++) { g.Ff[dir][v] = h->scratch + off; off += na; }
  for (int dir=0; dir<3; dir++) { g.eta[dir] = h->scratch + off; off += na; }
  for (int v=0; v<nvar; v++) g.Uin[v] = g.U[v];
  return 0;
}

extern "C" {

/* x3-chunked CTU scratch (the reference sizes its scratch to the whole Grid, integrate_3d_ctu.c:3374-3472): planes = active
 * x3 planes per chunk, 0 = whole Grid.  Results do not depend on it. */
int ab2_set_scratch_planes(ab2_grid* h, int planes) { return ab2_alloc_scratch(h, planes); }
int ab2_scratch_planes(const ab2_grid* h) { return h->chunk_planes; }

int ab2_profile_enable(ab2_grid* h, int on)
{
  prof_drain(h);
  h->prof_on = (on != 0);
  return 0;
}

int ab2_profile_kernels(void) { return KID_N; }
const char* ab2_profile_kernel_name(int id) { return (id >= 0 && id < KID_N) ? KERNEL_NAMES[id] : ""; }

int ab2_profile_read(ab2_grid* h, double* ms, long long* counts, int reset)
{
  prof_drain(h);
  for (int n=0; n<KID_N; n++) { ms[n] = h->prof_ms[n]; counts[n] = h->prof_cnt[n]; }
  if (reset) for (int n=0; n<KID_N; n++) { h->prof_ms[n] = 0.0; h->prof_cnt[n] = 0; }
  return 0;
}

const char* ab2_last_error(void) { return g_err.c_str(); }
const char* ab2_version(void) { return "athena_b200 0.1 (sm_100a, fp64, fmad=off)"; }

int ab2_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

ab2_grid* ab2_grid_create(const ab2_grid_desc* d)
{
  if (!d || d->nx1 < 1 || d->nx2 < 1 || d->nx3 < 1 || d->nscal < 0 || d->nscal > 1) {
    ab2_set_error("[ab2_grid_create]: bad descriptor"); return nullptr;
  }
  if (d->nx1 < 4 || d->nx2 < 4 || d->nx3 < 4) {
    ab2_set_error("[ab2_grid_create]: the 3D CTU path needs at least 4 zones per direction"); return nullptr;
  }
  if (ab2_device_count() <= d->device) {
    ab2_set_error("[ab2_grid_create]: no CUDA device (this engine has no CPU fallback)"); return nullptr;
  }
  if (!ab2_check(cudaSetDevice(d->device), "cudaSetDevice")) return nullptr;
  ab2_grid* h = new ab2_grid();
  memset(&h->g, 0, sizeof h->g);
  h->nscal = d->nscal; h->device = d->device;
  h->ion_ch_flags = nullptr; h->ion_nchunks = 0; h->ion_last_ntask = 0; h->ion_last_nskip = 0;
  h->has_grav = false; h->has_ion = false; h->n_fixed = 0; h->d_fixed_cells = nullptr; h->d_fixed_vals = nullptr;
  h->bytes_allocated = 0; h->launches = 0; h->maxiter = 0; h->tcoarse = 0.0;
  h->prof_on = false;
  for (int f=0; f<6; f++) { h->bc_n[f] = 0; h->bc_dst[f] = nullptr; h->bc_src[f] = nullptr; h->bc_coef[f] = nullptr; }
  h->comm = nullptr; h->lo_rank = h->hi_rank = -1; h->d_gather = nullptr; h->h_gather = nullptr;
  for (int n=0; n<KID_N; n++) { h->prof_ms[n] = 0.0; h->prof_cnt[n] = 0; }
  memset(&h->ion, 0, sizeof h->ion);
  if (d->use_external_stream) { h->stream = (cudaStream_t)d->stream; h->own_stream = false; }
  else {
    if (!ab2_check(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete h; return nullptr; }
    h->own_stream = true;
  }
  GridDev& g = h->g;
  g.nx[0] = d->nx1; g.nx[1] = d->nx2; g.nx[2] = d->nx3;
  for (int a=0; a<3; a++) g.n[a] = g.nx[a] + 2*AB2_NG;
  g.st[0] = 1; g.st[1] = g.n[0]; g.st[2] = (long)g.n[0]*g.n[1];
  g.dx[0] = d->dx1; g.dx[1] = d->dx2; g.dx[2] = d->dx3;
  g.Gamma = d->gamma; g.Gamma_1 = d->gamma - 1.0;
  g.riemann = AB2_RIEMANN_ROE;
  g.order = 2;
  h->xmin[0] = d->x1min; h->xmin[1] = d->x2min; h->xmin[2] = d->x3min;
  const size_t ncell = (size_t)g.n[0]*g.n[1]*g.n[2];
  const int nvar = 5 + d->nscal;

  h->own_U = (d->external_U[0] == nullptr);
  for (int v=0; v<nvar; v++) {
    g.U[v] = h->own_U ? dev_alloc(h, ncell) : d->external_U[v];
    if (!g.U[v]) { ab2_set_error(h->own_U ? g_err : std::string("[ab2_grid_create]: external_U incomplete")); ab2_grid_destroy(h); return nullptr; }
  }
  // CTU scratch (also the staging area of the AoS transposes).  AB2_CHUNK_PLANES=n forces x3 chunks of n active planes;
  // by default the scratch spans the whole Grid when that leaves room for the rest of the Grid's arrays, else it is chunked.
  h->scratch = nullptr; h->scratch_doubles = 0; h->chunk_planes = 0;
  // Passes C+D: the pair k_correct3 / k_final is the default.  The fused k_cfinal halves their HBM traffic (ncu: 25.9 -> 11.7 GB
  // per 256^3 step) but runs at 24 % occupancy with 30 % more instructions (halo + recomputed corrections) and measures
  // SLOWER on B200 (6.3 vs 5.0 ms; profiles/r2a_ncu_full_table.md): opt-in with AB2_HYDRO_FUSE=1.
  h->fuse = 0;
  if (const char* env = getenv("AB2_HYDRO_FUSE")) h->fuse = (atoi(env) != 0);
  h->first_split = 0;
  if (const char* env = getenv("AB2_FIRST_SPLIT")) h->first_split = atoi(env) & 7;
  h->correct3 = 1;
  if (const char* env = getenv("AB2_CORRECT3")) h->correct3 = (atoi(env) != 0);
  {
    int planes = 0;
    if (const char* env = getenv("AB2_CHUNK_PLANES")) planes = atoi(env);
    else {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const size_t P = (size_t)g.n[0]*g.n[1]*sizeof(double);
      const size_t rest = (size_t)24*ncell*sizeof(double);        // still to come: ion arrays (8.5), EdgeFlux (2), potential tables (4 typical, up to 22)
      const size_t full = (size_t)(12*nvar + 3)*ncell*sizeof(double);
      if (full + rest > (size_t)(0.92*(double)free_b)) {
        const double avail = 0.92*(double)free_b - (double)rest;
        long depth = (long)(avail/((double)(13*nvar + 3)*(double)P));       // 12 nvar + 3 scratch arrays + the staged copy of U
        planes = (int)(depth - 2*AB2_NG);
        planes -= planes % 8;
        if (planes < 8) { ab2_set_error("[ab2_grid_create]: not enough device memory for even an 8-plane chunk of CTU scratch"); ab2_grid_destroy(h); return nullptr; }
      }
    }
    if (ab2_alloc_scratch(h, planes) != 0) { ab2_grid_destroy(h); return nullptr; }
  }

  void* p = nullptr;
  if (!ab2_check(cudaMalloc(&p, SC_N*sizeof(unsigned long long)), "cudaMalloc scalars")) { ab2_grid_destroy(h); return nullptr; }
  h->d_scal = (unsigned long long*)p; h->allocs.push_back(p);
  cudaMemsetAsync(h->d_scal, 0, SC_N*sizeof(unsigned long long), h->stream);
  if (!ab2_check(cudaHostAlloc(&p, SC_N*sizeof(unsigned long long), cudaHostAllocDefault), "cudaHostAlloc")) { ab2_grid_destroy(h); return nullptr; }
  h->h_scal = (unsigned long long*)p;
  memset(h->h_scal, 0, SC_N*sizeof(unsigned long long));
  if (!ab2_check(cudaStreamSynchronize(h->stream), "create sync")) { ab2_grid_destroy(h); return nullptr; }
  return h;
}

void ab2_grid_destroy(ab2_grid* h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  prof_drain(h);
  if (h->scratch) cudaFree(h->scratch);
  if (h->d_fixed_cells) { cudaFree(h->d_fixed_cells); cudaFree(h->d_fixed_vals); }
  for (void* p : h->phi_allocs) cudaFree(p);
  for (int f=0; f<6; f++) if (h->bc_dst[f]) { cudaFree(h->bc_dst[f]); cudaFree(h->bc_src[f]); cudaFree(h->bc_coef[f]); }
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  for (void* p : h->allocs) cudaFree(p);
  if (h->h_scal) cudaFreeHost(h->h_scal);
  if (h->h_gather) cudaFreeHost(h->h_gather);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
}

long long ab2_bytes_allocated(const ab2_grid* h) { return (long long)h->bytes_allocated; }
long long ab2_kernel_launches(const ab2_grid* h) { return h->launches; }
int ab2_synchronize(ab2_grid* h) { CK(cudaStreamSynchronize(h->stream), "synchronize"); return 0; }
double* ab2_device_ptr(ab2_grid* h, int var) { return (var >= 0 && var < 5+h->nscal) ? h->g.U[var] : nullptr; }

// ---- coherence ----------------------------------------------------------------------------------------
// The transposes stage through the CTU scratch (dead between steps), in x3 slabs sized to it.
int ab2_upload_cons_aos(ab2_grid* h, const double* U_aos)
{
  GridDev& g = h->g;
  const int nvar = 5 + h->nscal;
  const long P = (long)g.n[0]*g.n[1];
  long slab = (long)(h->scratch_doubles/((size_t)P*nvar)); if (slab > g.n[2]) slab = g.n[2];
  for (long k0 = 0; k0 < g.n[2]; k0 += slab) {
    const long nk = (k0 + slab <= g.n[2]) ? slab : g.n[2] - k0;
    const long ncell = nk*P;
    CK(cudaMemcpyAsync(h->scratch, U_aos + (size_t)k0*P*nvar, (size_t)ncell*nvar*sizeof(double), cudaMemcpyHostToDevice, h->stream), "upload aos");
    AB2_LAUNCH(h, KID_TRANSPOSE, (k_aos_to_soa<<<(unsigned)((ncell+255)/256), 256, 0, h->stream>>>(g, nvar, k0*P, ncell, h->scratch)));
  }
  CK(cudaStreamSynchronize(h->stream), "upload aos sync");
  return 0;
}

int ab2_download_cons_aos(ab2_grid* h, double* U_aos)
{
  GridDev& g = h->g;
  const int nvar = 5 + h->nscal;
  const long P = (long)g.n[0]*g.n[1];
  long slab = (long)(h->scratch_doubles/((size_t)P*nvar)); if (slab > g.n[2]) slab = g.n[2];
  for (long k0 = 0; k0 < g.n[2]; k0 += slab) {
    const long nk = (k0 + slab <= g.n[2]) ? slab : g.n[2] - k0;
    const long ncell = nk*P;
    AB2_LAUNCH(h, KID_TRANSPOSE, (k_soa_to_aos<<<(unsigned)((ncell+255)/256), 256, 0, h->stream>>>(g, nvar, k0*P, ncell, h->scratch)));
    CK(cudaMemcpyAsync(U_aos + (size_t)k0*P*nvar, h->scratch, (size_t)ncell*nvar*sizeof(double), cudaMemcpyDeviceToHost, h->stream), "download aos");
  }
  CK(cudaStreamSynchronize(h->stream), "download aos sync");
  return 0;
}

int ab2_upload_cons_soa(ab2_grid* h, const double* const vars[6])
{
  const size_t nb = (size_t)h->g.n[0]*h->g.n[1]*h->g.n[2]*sizeof(double);
  for (int v=0; v<5+h->nscal; v++) CK(cudaMemcpyAsync(h->g.U[v], vars[v], nb, cudaMemcpyHostToDevice, h->stream), "upload soa");
  CK(cudaStreamSynchronize(h->stream), "upload soa sync");
  return 0;
}

int ab2_download_cons_soa(ab2_grid* h, double* const vars[6])
{
  const size_t nb = (size_t)h->g.n[0]*h->g.n[1]*h->g.n[2]*sizeof(double);
  for (int v=0; v<5+h->nscal; v++) CK(cudaMemcpyAsync(vars[v], h->g.U[v], nb, cudaMemcpyDeviceToHost, h->stream), "download soa");
  CK(cudaStreamSynchronize(h->stream), "download soa sync");
  return 0;
}

static size_t edge_count(const ab2_grid* h) { return (size_t)(h->g.nx[0]+1)*(h->g.nx[1]+1)*(h->g.nx[2]+1); }

int ab2_upload_edgeflux(ab2_grid* h, const double* ef)
{
  if (!h->has_ion) { ab2_set_error("[ab2_upload_edgeflux]: ab2_ion_init not called"); return -1; }
  CK(cudaMemcpyAsync(h->ion.edgeflux, ef, edge_count(h)*sizeof(double), cudaMemcpyHostToDevice, h->stream), "upload edgeflux");
  CK(cudaMemcpyAsync(h->ion.edgeflux_alt, ef, edge_count(h)*sizeof(double), cudaMemcpyHostToDevice, h->stream), "upload edgeflux");
  CK(cudaStreamSynchronize(h->stream), "upload edgeflux sync");
  return 0;
}

int ab2_download_edgeflux(ab2_grid* h, double* ef)
{
  if (!h->has_ion) { ab2_set_error("[ab2_download_edgeflux]: ab2_ion_init not called"); return -1; }
  CK(cudaMemcpyAsync(ef, h->ion.edgeflux, edge_count(h)*sizeof(double), cudaMemcpyDeviceToHost, h->stream), "download edgeflux");
  CK(cudaStreamSynchronize(h->stream), "download edgeflux sync");
  return 0;
}

// ---- hydro -----------------------------------------------------------------------------------------------
int ab2_bvals_mhd(ab2_grid* h, const int bc[6])
{ // physical faces by kernels, faces shared with a neighbouring Grid by the NCCL exchange (bvals_mhd.c:423-446)
  if (hydro_bvals(h, bc) != 0) return -1;
  return comm_exchange_x3(h);
}
int ab2_max_wave_speeds(ab2_grid* h, double out[3]) { return hydro_max_speeds(h, out); }

int ab2_new_dt(ab2_grid* h, double cour_no, double dt_old, int nstep, double time, double tlim,
               ab2_reduce_fn reduce, void* ctx, double* dt_out)
{ // new_dt.c:155-185
  double v[3];
  if (hydro_max_speeds(h, v) != 0) return -1;
  double max_dti = 0.0;
  for (int d=0; d<3; d++) if (h->g.nx[d] > 1) { const double x = v[d]/h->g.dx[d]; max_dti = (max_dti > x) ? max_dti : x; }
  double dt;
  if (nstep == 0) dt = cour_no/max_dti;
  else { const double a = 2.0*dt_old, b = cour_no/max_dti; dt = (a < b) ? a : b; }
  if (reduce) reduce(&dt, 1, 0, ctx);                 // new_dt.c:175-179 MPI_Allreduce(MIN)
  if ((time < tlim) && ((tlim - time) < dt)) dt = tlim - time;
  *dt_out = dt;
  return 0;
}

int ab2_integrate_3d_ctu(ab2_grid* h, double dt) { return hydro_integrate(h, dt); }

int ab2_set_fixed_cells(ab2_grid* h, long n, const long* cells, const double* vals)
{
  h->n_fixed = 0;
  CK(cudaStreamSynchronize(h->stream), "fixed cells sync");
  if (h->d_fixed_cells) { cudaFree(h->d_fixed_cells); cudaFree(h->d_fixed_vals); h->d_fixed_cells = nullptr; h->d_fixed_vals = nullptr; }   // a re-set replaces the table
  if (n <= 0) return 0;
  void* p = nullptr;
  CK(cudaMalloc(&p, n*sizeof(long)), "cudaMalloc fixed cells"); h->d_fixed_cells = (long*)p;
  CK(cudaMalloc(&p, (size_t)6*n*sizeof(double)), "cudaMalloc fixed vals"); h->d_fixed_vals = (double*)p;
  CK(cudaMemcpyAsync(h->d_fixed_cells, cells, n*sizeof(long), cudaMemcpyHostToDevice, h->stream), "fixed cells H2D");
  CK(cudaMemcpyAsync(h->d_fixed_vals, vals, (size_t)6*n*sizeof(double), cudaMemcpyHostToDevice, h->stream), "fixed vals H2D");
  CK(cudaStreamSynchronize(h->stream), "fixed sync");
  h->n_fixed = n;
  return 0;
}

int ab2_apply_fixed_cells(ab2_grid* h) { return hydro_apply_fixed(h); }

int ab2_set_bc_table(ab2_grid* h, int face, long n, const long* dst, const long* src, const double* coef)
{
  if (face < 0 || face > 5 || n < 0) { ab2_set_error("[ab2_set_bc_table]: face 0..5 (ix1, ox1, ix2, ox2, ix3, ox3)"); return -1; }
  const int nvar = 5 + h->nscal;
  const long ncell = (long)h->g.n[0]*h->g.n[1]*h->g.n[2];
  for (long t=0; t<n; t++) {
    if (dst[t] < 0 || dst[t] >= ncell) { ab2_set_error("[ab2_set_bc_table]: destination outside the Grid"); return -1; }
    for (int v=0; v<nvar; v++) if (src[(long)v*n+t] >= ncell) { ab2_set_error("[ab2_set_bc_table]: source outside the Grid"); return -1; }
  }
  CK(cudaStreamSynchronize(h->stream), "bc table sync");
  if (h->bc_dst[face]) { cudaFree(h->bc_dst[face]); cudaFree(h->bc_src[face]); cudaFree(h->bc_coef[face]); h->bc_dst[face] = nullptr; }
  h->bc_n[face] = 0;
  if (n == 0) return 0;
  void* p = nullptr;
  CK(cudaMalloc(&p, n*sizeof(long)), "cudaMalloc bc table"); h->bc_dst[face] = (long*)p;
  CK(cudaMalloc(&p, (size_t)nvar*n*sizeof(long)), "cudaMalloc bc table"); h->bc_src[face] = (long*)p;
  CK(cudaMalloc(&p, (size_t)nvar*n*sizeof(double)), "cudaMalloc bc table"); h->bc_coef[face] = (double*)p;
  CK(cudaMemcpyAsync(h->bc_dst[face], dst, n*sizeof(long), cudaMemcpyHostToDevice, h->stream), "bc table H2D");
  CK(cudaMemcpyAsync(h->bc_src[face], src, (size_t)nvar*n*sizeof(long), cudaMemcpyHostToDevice, h->stream), "bc table H2D");
  CK(cudaMemcpyAsync(h->bc_coef[face], coef, (size_t)nvar*n*sizeof(double), cudaMemcpyHostToDevice, h->stream), "bc table H2D");
  CK(cudaStreamSynchronize(h->stream), "bc table sync");
  h->bc_n[face] = n;
  return 0;
}

// ---- static potential -------------------------------------------------------------------------------------
// 1-D coordinate kinds per direction: the four argument forms the reference uses
enum { CK_C = 0, CK_M = 1, CK_MH = 2, CK_PH = 3 };

int ab2_set_riemann_solver(ab2_grid* h, int kind)
{
  if (!h || (kind != AB2_RIEMANN_ROE && kind != AB2_RIEMANN_HLLC && kind != AB2_RIEMANN_ROE_NOH)) {
    ab2_set_error("[ab2_set_riemann_solver]: kind must be 0 (Roe + H-correction), 1 (HLLC) or 2 (Roe without H-correction)");
    return -1;
  }
  h->g.riemann = kind;
  return 0;
}

int ab2_set_reconstruction(ab2_grid* h, int order)
{
  if (!h || (order != 2 && order != 3)) {
    ab2_set_error("[ab2_set_reconstruction]: order must be 2 (PLM) or 3 (PPM), both in characteristic variables");
    return -1;
  }
  h->g.order = order;
  return 0;
}

int ab2_set_static_grav_pot(ab2_grid* h, ab2_grav_pot_fn phi)
{
  GridDev& g = h->g;
  const size_t ncell = (size_t)g.n[0]*g.n[1]*g.n[2];
  CK(cudaStreamSynchronize(h->stream), "grav sync");
  for (void* p : h->phi_allocs) { cudaFree(p); h->bytes_allocated -= ncell*sizeof(double); }      // a re-set replaces the tables
  h->phi_allocs.clear();
  if (!phi) { h->has_grav = false; return 0; }
  // coordinates exactly as cc_pos() and the call sites form them
  std::vector<double> X[3][4];
  bool alias_m[3], alias_mh[3];
  for (int d=0; d<3; d++) {
    const int n = g.n[d];
    const double x0 = h->xmin[d], dx = g.dx[d];
    for (int kd=0; kd<4; kd++) X[d][kd].resize(n);
    for (int i=0; i<n; i++) {
      const double xc = x0 + ((double)(i - AB2_NG) + 0.5)*dx;      // cc_pos.c
      X[d][CK_C][i] = xc;
      X[d][CK_M][i] = (xc - dx);
      X[d][CK_MH][i] = (xc - 0.5*dx);
      X[d][CK_PH][i] = (xc + 0.5*dx);
    }
    alias_m[d] = true; alias_mh[d] = true;
    for (int i=1; i<n; i++) {
      if (X[d][CK_M][i] != X[d][CK_C][i-1]) alias_m[d] = false;
      if (X[d][CK_MH][i] != X[d][CK_PH][i-1]) alias_mh[d] = false;
    }
  }
  std::vector<double> host(ncell);
  auto build = [&](int k0, int k1, int k2) -> double* {
    const int kinds[3] = {k0,k1,k2};
    const int n1 = g.n[0], n2 = g.n[1], n3 = g.n[2];
    // A problem file's potential is plain C written for a serial code and may touch statics: it is called from ONE thread
    // unless the user vouches for re-entrancy with AB2_GRAV_THREADS=n (athena_b200.h).
    unsigned nthr = 1;
    const char* env = getenv("AB2_GRAV_THREADS"); if (env) { int v = atoi(env); if (v >= 1 && v <= 256) nthr = (unsigned)v; }
    std::vector<std::thread> pool;
    for (unsigned t=0; t<nthr; t++) pool.emplace_back([&, t]() {
      for (int k=(int)t; k<n3; k+=(int)nthr) for (int j=0; j<n2; j++) {
        double* rowp = host.data() + ((size_t)k*n2 + j)*n1;
        const double x2 = X[1][kinds[1]][j], x3 = X[2][kinds[2]][k];
        for (int i=0; i<n1; i++) rowp[i] = phi(X[0][kinds[0]][i], x2, x3);
      }
    });
    for (auto& th : pool) th.join();
    void* raw = nullptr;
    if (!ab2_check(cudaMalloc(&raw, ncell*sizeof(double)), "cudaMalloc phi table")) return nullptr;
    double* dptr = (double*)raw;
    h->phi_allocs.push_back(raw);
    h->bytes_allocated += ncell*sizeof(double);
    if (!ab2_check(cudaMemcpyAsync(dptr, host.data(), ncell*sizeof(double), cudaMemcpyHostToDevice, h->stream), "phi H2D")) return nullptr;
    if (!ab2_check(cudaStreamSynchronize(h->stream), "phi sync")) return nullptr;
    return dptr;
  };
  auto kinds_for = [](int d, int kind, int out[3]) { out[0] = out[1] = out[2] = CK_C; out[d] = kind; };

  const double* T[PHI_NTAB];
  int kk[3];
  if (!(T[PHI_C] = build(CK_C,CK_C,CK_C))) return -1;
  for (int d=0; d<3; d++) { kinds_for(d, CK_PH, kk); if (!(T[PHI_PH0+d] = build(kk[0],kk[1],kk[2]))) return -1; }
  for (int d=0; d<3; d++) {
    if (alias_mh[d]) T[PHI_MH0+d] = T[PHI_PH0+d] - g.st[d];
    else { kinds_for(d, CK_MH, kk); if (!(T[PHI_MH0+d] = build(kk[0],kk[1],kk[2]))) return -1; }
    if (alias_m[d]) T[PHI_M0+d] = T[PHI_C] - g.st[d];
    else { kinds_for(d, CK_M, kk); if (!(T[PHI_M0+d] = build(kk[0],kk[1],kk[2]))) return -1; }
  }
  for (int d=0; d<3; d++) for (int t=0; t<3; t++) {
    if (t == d) continue;
    for (int plus=0; plus<2; plus++) {
      const int id = phi_mix(d,t,plus);
      if (alias_m[d]) T[id] = T[(plus ? PHI_PH0 : PHI_MH0) + t] - g.st[d];
      else { kk[0] = kk[1] = kk[2] = CK_C; kk[d] = CK_M; kk[t] = plus ? CK_PH : CK_MH; if (!(T[id] = build(kk[0],kk[1],kk[2]))) return -1; }
    }
  }
  for (int n=0; n<PHI_NTAB; n++) g.phi[n] = T[n];
  CK(cudaStreamSynchronize(h->stream), "grav sync");
  h->has_grav = true;
  return 0;
}

// ---- ion ------------------------------------------------------------------------------------------------
void ab2_ion_derived(ab2_ion_par* p, double dx1, double dx2, double dx3)
{ // ionrad.c:115-131 (including the dx[1]-twice quirk of :129)
  const double area1 = dx1*dx2, area2 = dx1*dx3, area3 = dx2*dx3;
  if (area1 < area2) { if (area1 < area3) p->min_area = area1; else p->min_area = area3; }
  else { if (area2 < area3) p->min_area = area2; else p->min_area = area3; }
  double maxdx = dx1 > dx2 ? dx1 : dx2;
  maxdx = maxdx > dx3 ? maxdx : dx2;
  p->d_nlo = 1.0e-4 * p->m_H / (p->sigma_ph * maxdx);        // MINOPTDEPTH, ionrad.h:29
}

double ab2_ion_entering_flux(double flux_i, double time)
{ // ionradplane_3d.c:265
  return flux_i*(5.*(erf((time - 1.2e5)/8e4)+1)+0.1);
}

int ab2_ion_init(ab2_grid* h, const ab2_ion_par* par)
{
  if (h->nscal < 1) { ab2_set_error("[ab2_ion_init]: ion radiation needs NSCALARS >= 1"); return -1; }
  IonParDev& p = h->ionpar;
  p.sigma_ph = par->sigma_ph; p.m_H = par->m_H; p.mu = par->mu; p.e_gamma = par->e_gamma; p.alpha_C = par->alpha_C;
  p.k_B = par->k_B; p.time_unit = par->time_unit; p.max_de_iter = par->max_de_iter;
  p.max_de_therm_iter = par->max_de_therm_iter; p.max_dx_iter = par->max_dx_iter; p.max_de_step = par->max_de_step;
  p.max_de_therm_step = par->max_de_therm_step; p.max_dx_step = par->max_dx_step; p.tfloor = par->tfloor; p.tceil = par->tceil;
  p.min_area = par->min_area; p.d_nlo = par->d_nlo;
  h->maxiter = par->maxiter;
  if (!(par->max_de_therm_step > 0)) { ab2_set_error("[ab2_ion_init]: max_de_therm_step must be > 0 (check_range reads an unset value otherwise, ionrad_3d.c:227-241)"); return -1; }
  if (!h->has_ion) {
    const size_t ncell = (size_t)h->g.n[0]*h->g.n[1]*h->g.n[2];
    IonDev& I = h->ion;
    if (!(I.ph_rate = dev_alloc(h, ncell)) || !(I.edot = dev_alloc(h, ncell)) || !(I.nHdot = dev_alloc(h, ncell)) ||
        !(I.e_init = dev_alloc(h, ncell)) || !(I.e_th_init = dev_alloc(h, ncell)) || !(I.x_init = dev_alloc(h, ncell)) ||
        !(I.edgeflux = dev_alloc(h, edge_count(h))) || !(I.edgeflux_alt = dev_alloc(h, edge_count(h)))) return -1;
    double* s = dev_alloc(h, (ncell+1)/2);
    if (!s) return -1;
    I.sign_state = (int*)s;
    // sub-cycle memo (engine.cuh IonChunks): one record per ray and 32-cell chunk, one task slot per ray
    IonChunks& C = I.ch;
    C.nch = (h->g.nx[0] + 31)/32;
    const size_t nrays = (size_t)h->g.nx[1]*h->g.nx[2], nchunks = nrays*C.nch;
    double* fl = dev_alloc(h, nchunks + 1), *tk = dev_alloc(h, nrays*sizeof(IonTask)/sizeof(double));
    if (!fl || !tk || !(C.rec = dev_alloc(h, 8*nchunks))) return -1;
    C.tasks = (IonTask*)tk;
    h->ion_ch_flags = (unsigned*)fl; h->ion_nchunks = nchunks;
    C.uflags = (unsigned*)fl + nchunks;
    const char* env = getenv("AB2_ION_SKIP");
    C.flags = (env && atoi(env) == 0) ? nullptr : h->ion_ch_flags;
    CK(cudaStreamSynchronize(h->stream), "ion init sync");
    h->has_ion = true;
  }
  return 0;
}

int ab2_bvals_ionrad(ab2_grid* h, double flux_i)
{
  if (!h->has_ion) { ab2_set_error("[ab2_bvals_ionrad]: ab2_ion_init not called"); return -1; }
  return ion_bvals_ionrad(h, flux_i);
}

int ab2_ion_radtransfer_3d(ab2_grid* h, double flux_i, double time, double cour_no, double* dt,
                           int* niter, int* nchem, int* ntherm, ab2_reduce_fn reduce, void* ctx)
{
  if (!h->has_ion) { ab2_set_error("[ab2_ion_radtransfer_3d]: ab2_ion_init not called"); return -1; }
  return ion_radtransfer(h, ab2_ion_entering_flux(flux_i, time), cour_no, dt, niter, nchem, ntherm, reduce, ctx);
}

int ab2_ion_radtransfer_3d_fine(ab2_grid* h, double tcoarse, double* dt, int* niter, int* nchem, int* ntherm,
                                ab2_reduce_fn reduce, void* ctx)
{
  if (!h->has_ion) { ab2_set_error("[ab2_ion_radtransfer_3d_fine]: ab2_ion_init not called"); return -1; }
  return ion_radtransfer_fine(h, tcoarse, dt, niter, nchem, ntherm, reduce, ctx);
}

double ab2_ion_coarse_time(const ab2_grid* h) { return h->tcoarse; }

int ab2_set_ion_memo(ab2_grid* h, int on)
{
  if (!h->has_ion) { ab2_set_error("[ab2_set_ion_memo]: ab2_ion_init not called"); return -1; }
  h->ion.ch.flags = on ? h->ion_ch_flags : nullptr;
  return 0;
}

int ab2_ion_memo_stats(ab2_grid* h, long long out[3])
{
  if (!h->has_ion) { ab2_set_error("[ab2_ion_memo_stats]: ab2_ion_init not called"); return -1; }
  return ion_memo_count(h, out);
}

} // extern "C"
