// smr.cu -- static mesh refinement operators between a parent Grid and a child Grid on the same GPU:
//   RestrictCorrect        (smr.c:1207-2352)  restriction (8-cell average), injection, flux correction
//   ionradRestrictCorrect  (smr.c:85-1205)    E and s[0] only
//   Prolongate + ProCon    (smr.c:2359-2923, 3068-3325, mcd_slope :3478-3495)
//   ionrad_prolong_snd/rcv (ionradiation/ionrad_smr.c:345-459, 34-343)
// The reference stages everything through send/receive buffers and the myFlx planes captured at the end of
// integrate_3d_ctu (step 12e, :3076-3364).  Here both Grids' final fluxes are still resident in their F arrays
// (physical component order), so the kernels read parent and child memory directly: no buffers, no capture pass.
// Summation orders follow the reference, so results are bit-identical to it.
#include "../../include/athena_b200.h"
#include "engine.cuh"
#include "physics.cuh"

#define NG AB2_NG

struct OvDev { int cs[3], ce[3], face[6], cdisp[3]; };

static OvDev ov_dev(const ab2_overlap* o)
{
  OvDev d;
  for (int a=0; a<3; a++) { d.cs[a] = o->cs[a]; d.ce[a] = o->ce[a]; d.cdisp[a] = o->cdisp[a]; }
  for (int a=0; a<6; a++) d.face[a] = o->face[a];
  return d;
}

// make stream `to` wait for everything queued on stream `from`
static bool stream_wait(cudaStream_t to, cudaStream_t from)
{
  if (to == from) return true;
  cudaEvent_t e;
  if (!ab2_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "smr event")) return false;
  bool ok = ab2_check(cudaEventRecord(e, from), "smr event record") && ab2_check(cudaStreamWaitEvent(to, e, 0), "smr stream wait");
  cudaEventDestroy(e);
  return ok;
}

__device__ __forceinline__ long cidx(const GridDev& g, int k, int j, int i) { return (long)k*g.st[2] + (long)j*g.st[1] + i; }

// ---- restriction + injection: one thread per parent cell of the overlap ----------------------------------------
__global__ void k_smr_restrict(GridDev pa, GridDev ch, OvDev o, int v0, int v1)
{
  const int ni = o.ce[0]-o.cs[0]+1, nj = o.ce[1]-o.cs[1]+1, nk = o.ce[2]-o.cs[2]+1;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)ni*nj*nk) return;
  const int ii = (int)(t % ni), jj = (int)((t / ni) % nj), kk = (int)(t / ((long)ni*nj));
  const long p = cidx(pa, o.cs[2]+kk, o.cs[1]+jj, o.cs[0]+ii);
  const long c = cidx(ch, NG+2*kk, NG+2*jj, NG+2*ii);
  const long s1 = ch.st[1], s2 = ch.st[2];
  for (int v=v0; v<=v1; v++) {
    const double* a = ch.U[v];
    double sum = a[c] + a[c+1];                                            // smr.c:1877-1891
    sum += a[c+s1] + a[c+s1+1];                                            // :1904-1918
    sum += a[c+s2] + a[c+s2+1] + a[c+s2+s1] + a[c+s2+s1+1];                // :1932-1955
    pa.U[v][p] = sum*0.125;
  }
}

// ---- flux correction of the coarse cells just outside the child, face `dim` ------------------------------------
__global__ void k_smr_flux_correct(GridDev pa, GridDev ch, OvDev o, int nvar, int dim, double q)
{
  const int d = dim/2, side = dim%2;
  const int t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
  const int na = o.ce[t1]-o.cs[t1]+1, nb = o.ce[t2]-o.cs[t2]+1;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)na*nb) return;
  const int a = (int)(t % na), b = (int)(t / na);
  int pc[3], pf[3], cf[3];
  pc[t1] = o.cs[t1]+a; pc[t2] = o.cs[t2]+b; pc[d] = side ? o.ce[d]+1 : o.cs[d]-1;
  pf[t1] = pc[t1]; pf[t2] = pc[t2]; pf[d] = side ? o.ce[d]+1 : o.cs[d];
  cf[t1] = NG + 2*a; cf[t2] = NG + 2*b; cf[d] = side ? NG + ch.nx[d] : NG;
  const long xc = cidx(pa, pc[2], pc[1], pc[0]), xf = cidx(pa, pf[2], pf[1], pf[0]);
  const long c00 = cidx(ch, cf[2], cf[1], cf[0]), e1 = ch.st[t1], e2 = ch.st[t2];
  for (int v=0; v<nvar; v++) {
    const double* Fc = ch.Ff[d][v];
    double r = Fc[c00] + Fc[c00+e1];                                       // smr.c:2046-2064 (and x2/x3 analogues)
    r += Fc[c00+e2] + Fc[c00+e2+e1];                                       // :2072-2090
    r *= 0.25;
    pa.U[v][xc] -= q*(pa.Ff[d][v][xf] - r);                                 // :1503-1517
  }
}

// ---- prolongation -----------------------------------------------------------------------------------------------
__device__ __forceinline__ double mcd_slope(double vl, double vc, double vr)
{ // smr.c:3478-3495
  const double dvl = (vc - vl), dvr = (vr - vc);
  if (dvl > 0.0 && dvr > 0.0) {
    const double dv = 2.0*(dvl < dvr ? dvl : dvr);
    const double dvm = 0.5*(dvl + dvr);
    return (dvm < dv ? dvm : dv);
  } else if (dvl < 0.0 && dvr < 0.0) {
    const double dv = 2.0*(dvl > dvr ? dvl : dvr);
    const double dvm = 0.5*(dvl + dvr);
    return (dvm > dv ? dvm : dv);
  }
  return 0.0;
}
__device__ __forceinline__ double pres_like(const GridDev& g, long x)
{ // smr.c:3209
  return g.U[4][x] - 0.5*(AB2_SQR(g.U[1][x]) + AB2_SQR(g.U[2][x]) + AB2_SQR(g.U[3][x]))/g.U[0][x];
}

// one thread per coarse cell of the slab of face `dim`; writes its 2x2x2 child ghost cells (ProCon)
__global__ void k_smr_prolongate(GridDev pa, GridDev ch, OvDev o, int nvar, int dim)
{
  int cs[3], ps[3], pe[3], n[3];
  for (int a=0; a<3; a++) {
    cs[a] = o.cs[a] - (NG/2) - 1;                      // parent slab start (smr.c:2442-2463)
    ps[a] = 0; pe[a] = ch.nx[a] + 2*NG - 1;            // child ghost box (:2667-2688): [is-nghost, ie+nghost]
  }
  const int d = dim/2;
  if (dim%2 == 0) pe[d] = NG - 1;
  else { ps[d] = ch.nx[d] + NG; cs[d] = o.ce[d]; }
  for (int a=0; a<3; a++) n[a] = (pe[a]-ps[a]+1)/2;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)n[0]*n[1]*n[2]) return;
  const int ii = 1 + (int)(t % n[0]), jj = 1 + (int)((t / n[0]) % n[1]), kk = 1 + (int)(t / ((long)n[0]*n[1]));
  const long x0 = cidx(pa, cs[2]+kk, cs[1]+jj, cs[0]+ii);
  const long c0 = cidx(ch, ps[2]+2*(kk-1), ps[1]+2*(jj-1), ps[0]+2*(ii-1));
  const long p1 = 1, p2 = pa.st[1], p3 = pa.st[2];
  double PC[8][6];
  for (int v=0; v<nvar; v++) {
    double dq1, dq2, dq3, base;
    if (v == 4) {
      base = pres_like(pa, x0);
      dq1 = mcd_slope(pres_like(pa,x0-p1), base, pres_like(pa,x0+p1));
      dq2 = mcd_slope(pres_like(pa,x0-p2), base, pres_like(pa,x0+p2));
      dq3 = mcd_slope(pres_like(pa,x0-p3), base, pres_like(pa,x0+p3));
    } else {
      const double* a = pa.U[v];
      base = a[x0];
      dq1 = mcd_slope(a[x0-p1], base, a[x0+p1]);
      dq2 = mcd_slope(a[x0-p2], base, a[x0+p2]);
      dq3 = mcd_slope(a[x0-p3], base, a[x0+p3]);
    }
#pragma unroll
    for (int m=0; m<8; m++) {
      const int i = m&1, j = (m>>1)&1, k = m>>2;
      PC[m][v] = base + (0.5*i - 0.25)*dq1 + (0.5*j - 0.25)*dq2 + (0.5*k - 0.25)*dq3;   // smr.c:3116-3117
    }
  }
#pragma unroll
  for (int m=0; m<8; m++) {
    const int i = m&1, j = (m>>1)&1, k = m>>2;
    PC[m][4] += 0.5*(AB2_SQR(PC[m][1]) + AB2_SQR(PC[m][2]) + AB2_SQR(PC[m][3]))/PC[m][0];   // :3240
    const long c = c0 + i + (long)j*ch.st[1] + (long)k*ch.st[2];
    for (int v=0; v<nvar; v++) ch.U[v][c] = PC[m][v];
  }
}

// ---- ionising flux hand-down: parent EdgeFlux plane at the child's upstream x1 face -> child EdgeFlux[..][..][0] ----
__global__ void k_smr_ion_prolong(GridDev pa, GridDev ch, OvDev o, const double* efp, double* efc_cur, double* efc_alt)
{
  const int nj = o.ce[1]-o.cs[1]+2, nk = o.ce[2]-o.cs[2]+2;
  const long t = (long)blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= (long)nj*nk) return;
  const int j = o.cs[1]-NG + (int)(t % nj), k = o.cs[2]-NG + (int)(t / nj);
  const int pe1 = pa.nx[0]+1, pe2 = pa.nx[1]+1, ce1 = ch.nx[0]+1, ce2 = ch.nx[1]+1;
  const double val = efp[((long)k*pe2 + j)*pe1 + (o.cs[0]-NG)];             // ionrad_smr.c:375-387
  // :293-294 writes `k*2 - Disp` with k parent-local, which is right only below a root-level parent; the
  // same cell in parent-relative form, identical for level 1 and meaningful for deeper levels:
  const int ks = 2*(k - (o.cs[2]-NG)), js = 2*(j - (o.cs[1]-NG));
#define EFC(kk,jj) efc_cur[((long)(kk)*ce2 + (jj))*ce1] = efc_alt[((long)(kk)*ce2 + (jj))*ce1]
  EFC(ks,js) = val;
  if (j < (o.ce[1]+1-NG)) {
    if (k < (o.ce[2]+1-NG)) { EFC(ks+1,js+1) = val; EFC(ks,js+1) = val; EFC(ks+1,js) = val; }
    else EFC(ks,js+1) = val;
  } else if (k < (o.ce[2]+1-NG)) EFC(ks+1,js) = val;
#undef EFC
}

static int check_pair(ab2_grid* pa, ab2_grid* ch, const char* who)
{
  if (!pa || !ch || pa->device != ch->device || pa->nscal != ch->nscal) {
    ab2_set_error(std::string("[") + who + "]: parent and child must live on the same device with the same NSCALARS");
    return -1;
  }
  return 0;
}

extern "C" {

int ab2_smr_restrict_correct(ab2_grid* pa, ab2_grid* ch, const ab2_overlap* o, double dt_parent)
{
  if (check_pair(pa, ch, "RestrictCorrect")) return -1;
  if (pa->chunk_planes || ch->chunk_planes) {
    ab2_set_error("[RestrictCorrect]: the flux correction reads both Grids' final fluxes of the whole Grid: call ab2_set_scratch_planes(g, 0) on SMR Grids");
    return -1;
  }
  const OvDev od = ov_dev(o);
  const int nvar = 5 + pa->nscal;
  if (!stream_wait(pa->stream, ch->stream)) return -1;
  const long n = (long)(od.ce[0]-od.cs[0]+1)*(od.ce[1]-od.cs[1]+1)*(od.ce[2]-od.cs[2]+1);
  AB2_LAUNCH(pa, KID_SMR_RESTRICT, (k_smr_restrict<<<(unsigned)((n+127)/128), 128, 0, pa->stream>>>(pa->g, ch->g, od, 0, nvar-1)));
  for (int dim=0; dim<6; dim++) {
    if (!od.face[dim]) continue;
    const int d = dim/2, t1 = (d == 0) ? 1 : 0, t2 = (d == 2) ? 1 : 2;
    const long nf = (long)(od.ce[t1]-od.cs[t1]+1)*(od.ce[t2]-od.cs[t2]+1);
    const double q = (dim%2) ? (dt_parent/pa->g.dx[d]) : -(dt_parent/pa->g.dx[d]);       // smr.c:1499-1500
    AB2_LAUNCH(pa, KID_SMR_FLUXCORR, (k_smr_flux_correct<<<(unsigned)((nf+127)/128), 128, 0, pa->stream>>>(pa->g, ch->g, od, nvar, dim, q)));
  }
  if (!stream_wait(ch->stream, pa->stream)) return -1;
  return ab2_check(cudaGetLastError(), "RestrictCorrect") ? 0 : -1;
}

int ab2_smr_ion_restrict(ab2_grid* pa, ab2_grid* ch, const ab2_overlap* o)
{
  if (check_pair(pa, ch, "ionradRestrictCorrect")) return -1;
  const OvDev od = ov_dev(o);
  if (!stream_wait(pa->stream, ch->stream)) return -1;
  const long n = (long)(od.ce[0]-od.cs[0]+1)*(od.ce[1]-od.cs[1]+1)*(od.ce[2]-od.cs[2]+1);
  AB2_LAUNCH(pa, KID_SMR_RESTRICT, (k_smr_restrict<<<(unsigned)((n+127)/128), 128, 0, pa->stream>>>(pa->g, ch->g, od, 4, 5)));
  if (!stream_wait(ch->stream, pa->stream)) return -1;
  return ab2_check(cudaGetLastError(), "ionradRestrictCorrect") ? 0 : -1;
}

int ab2_smr_prolongate(ab2_grid* pa, ab2_grid* ch, const ab2_overlap* o)
{
  if (check_pair(pa, ch, "Prolongate")) return -1;
  const OvDev od = ov_dev(o);
  const int nvar = 5 + pa->nscal;
  if (!stream_wait(ch->stream, pa->stream)) return -1;
  for (int dim=0; dim<6; dim++) {
    if (!od.face[dim]) continue;
    long n = 1;
    for (int a=0; a<3; a++) n *= (a == dim/2) ? (NG/2) : (ch->g.nx[a] + 2*NG)/2;
    AB2_LAUNCH(ch, KID_SMR_PROLONG, (k_smr_prolongate<<<(unsigned)((n+63)/64), 64, 0, ch->stream>>>(pa->g, ch->g, od, nvar, dim)));
  }
  if (!stream_wait(pa->stream, ch->stream)) return -1;
  return ab2_check(cudaGetLastError(), "Prolongate") ? 0 : -1;
}

int ab2_smr_ion_prolong(ab2_grid* pa, ab2_grid* ch, const ab2_overlap* o)
{
  if (check_pair(pa, ch, "ionrad_prolong")) return -1;
  if (!pa->has_ion || !ch->has_ion) { ab2_set_error("[ionrad_prolong]: ab2_ion_init not called"); return -1; }
  const OvDev od = ov_dev(o);
  if (!od.face[0]) return 0;
  if (!stream_wait(ch->stream, pa->stream)) return -1;
  const long n = (long)(od.ce[1]-od.cs[1]+2)*(od.ce[2]-od.cs[2]+2);
  AB2_LAUNCH(ch, KID_SMR_PROLONG, (k_smr_ion_prolong<<<(unsigned)((n+127)/128), 128, 0, ch->stream>>>(pa->g, ch->g, od, pa->ion.edgeflux, ch->ion.edgeflux, ch->ion.edgeflux_alt)));
  if (!stream_wait(pa->stream, ch->stream)) return -1;
  return ab2_check(cudaGetLastError(), "ionrad_prolong") ? 0 : -1;
}

} // extern "C"
