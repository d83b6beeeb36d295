// comm.cu -- NCCL inside the library: the reference's MPI sites on the hot path
//   bvals_mhd.c:423-446 (+ pack/unpack :2317-3175)   ghost-zone exchange with the x3 neighbours
//   new_dt.c:175-179                                  MPI_Allreduce(MIN) of the timestep
//   ionradiation/ionrad_3d.c:275,399,554,672          MPI_Allreduce of the sub-cycle scalars
// One communicator per group of Grids that reduce together (a Domain's Grids), one rank per GPU / process.  The
// x3 ghost slabs of the SoA layout are contiguous per variable, so they travel in place: ncclSend/ncclRecv straight
// out of / into U, all transfers of one exchange in ONE NCCL group (one fused kernel over NVLink), no pack kernels.
// Reductions stay on the device: the scalars the kernels leave in d_scal are all-gathered and combined on the host
// after a single D2H copy (min / max / integer sum: order-independent, so every rank gets the same bits).
//
// libnccl is opened with dlopen at the first ab2_comm_* call (AB2_NCCL_LIB overrides the name): the library itself
// has no link-time dependency on it, loads on a box without NCCL, and inside a torch process binds to the libnccl
// torch already loaded.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../include/athena_b200.h"
#include "engine.cuh"

#define NG AB2_NG

struct NcclApi {
  void* so;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
};
static NcclApi N = {};

static bool nccl_load()
{
  if (N.so) return true;
  const char* names[] = { getenv("AB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
  void* so = nullptr;
  for (const char* n : names) { if (n && *n && (so = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
  if (!so) { ab2_set_error(std::string("[ab2_comm]: cannot load libnccl (set AB2_NCCL_LIB): ") + dlerror()); return false; }
#define SYM(field, name) do { *(void**)(&N.field) = dlsym(so, name); if (!N.field) { ab2_set_error("[ab2_comm]: libnccl lacks " name); return false; } } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(GetErrorString, "ncclGetErrorString"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather");
#undef SYM
  N.so = so;
  return true;
}

static bool nccl_ok(ncclResult_t r, const char* what)
{
  if (r == ncclSuccess) return true;
  ab2_set_error(std::string("[athena_b200] NCCL ") + what + ": " + N.GetErrorString(r));
  return false;
}

struct ab2_comm {
  ncclComm_t comm;
  int nranks, rank, device;
  cudaStream_t stream;               // for host-scalar reductions that belong to no Grid
  double* d_buf;                     // 64 doubles
  double* h_buf;                     // pinned
  long long n_exchanges, n_reductions;
};

extern "C" {

int ab2_comm_get_unique_id(char id[AB2_UNIQUE_ID_BYTES])
{
  if (!nccl_load()) return -1;
  ncclUniqueId u;
  if (!nccl_ok(N.GetUniqueId(&u), "ncclGetUniqueId")) return -1;
  static_assert(sizeof(u) == AB2_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &u, sizeof u);
  return 0;
}

ab2_comm* ab2_comm_init_rank(int nranks, int rank, const char id[AB2_UNIQUE_ID_BYTES], int device)
{
  if (!nccl_load()) return nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) { ab2_set_error("[ab2_comm_init_rank]: bad rank / size"); return nullptr; }
  if (!ab2_check(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
  ab2_comm* c = new ab2_comm();
  memset(c, 0, sizeof *c);
  c->nranks = nranks; c->rank = rank; c->device = device;
  ncclUniqueId u; memcpy(&u, id, sizeof u);
  if (!nccl_ok(N.CommInitRank(&c->comm, nranks, u, rank), "ncclCommInitRank")) { delete c; return nullptr; }
  void* p = nullptr;
  if (!ab2_check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "comm stream") ||
      !ab2_check(cudaMalloc(&p, 64*sizeof(double)), "comm buffer")) { delete c; return nullptr; }
  c->d_buf = (double*)p;
  if (!ab2_check(cudaHostAlloc(&p, 64*sizeof(double), cudaHostAllocDefault), "comm host buffer")) { delete c; return nullptr; }
  c->h_buf = (double*)p;
  return c;
}

void ab2_comm_destroy(ab2_comm* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm && N.so) N.CommDestroy(c->comm);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->h_buf) cudaFreeHost(c->h_buf);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int ab2_comm_rank(const ab2_comm* c) { return c->rank; }
int ab2_comm_size(const ab2_comm* c) { return c->nranks; }
long long ab2_comm_exchanges(const ab2_comm* c) { return c->n_exchanges; }
long long ab2_comm_reductions(const ab2_comm* c) { return c->n_reductions; }

// MPI_Allreduce of n <= 64 host doubles (op 0 MIN, 1 MAX, 2 SUM): new_dt.c:175-179 for callers that combine
// several Grids on the host (SMR meshes), main.c's zone-cycle totals, etc.
int ab2_comm_allreduce(ab2_comm* c, double* vals, int n, int op)
{
  if (n < 1 || n > 64 || op < 0 || op > 2) { ab2_set_error("[ab2_comm_allreduce]: n in 1..64, op in 0..2"); return -1; }
  if (c->nranks == 1) return 0;
  if (!ab2_check(cudaSetDevice(c->device), "cudaSetDevice")) return -1;
  memcpy(c->h_buf, vals, n*sizeof(double));
  if (!ab2_check(cudaMemcpyAsync(c->d_buf, c->h_buf, n*sizeof(double), cudaMemcpyHostToDevice, c->stream), "allreduce H2D")) return -1;
  const ncclRedOp_t rop = (op == 0) ? ncclMin : (op == 1) ? ncclMax : ncclSum;
  if (!nccl_ok(N.AllReduce(c->d_buf, c->d_buf, (size_t)n, ncclDouble, rop, c->comm, c->stream), "ncclAllReduce")) return -1;
  if (!ab2_check(cudaMemcpyAsync(c->h_buf, c->d_buf, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream), "allreduce D2H")) return -1;
  if (!ab2_check(cudaStreamSynchronize(c->stream), "allreduce sync")) return -1;
  memcpy(vals, c->h_buf, n*sizeof(double));
  c->n_reductions++;
  return 0;
}

// MPI_Bcast of raw bytes from `root` (par_dist_mpi-style payloads are the driver's business; this is here so a
// driver without MPI can still distribute small blobs, e.g. a second communicator's unique id)
int ab2_comm_bcast(ab2_comm* c, void* data, int nbytes, int root)
{
  if (nbytes < 0 || nbytes > (int)(64*sizeof(double))) { ab2_set_error("[ab2_comm_bcast]: at most 512 bytes"); return -1; }
  if (c->nranks == 1) return 0;
  if (!ab2_check(cudaSetDevice(c->device), "cudaSetDevice")) return -1;
  // every rank contributes its bytes as u64 words, zeros on the non-roots: a SUM leaves the root's words on every rank
  unsigned long long w[64]; memset(w, 0, sizeof w);
  if (c->rank == root) memcpy(w, data, nbytes);
  memcpy(c->h_buf, w, sizeof w);
  if (!ab2_check(cudaMemcpyAsync(c->d_buf, c->h_buf, sizeof w, cudaMemcpyHostToDevice, c->stream), "bcast H2D")) return -1;
  if (!nccl_ok(N.AllReduce(c->d_buf, c->d_buf, 64, ncclUint64, ncclSum, c->comm, c->stream), "ncclAllReduce")) return -1;
  if (!ab2_check(cudaMemcpyAsync(c->h_buf, c->d_buf, sizeof w, cudaMemcpyDeviceToHost, c->stream), "bcast D2H")) return -1;
  if (!ab2_check(cudaStreamSynchronize(c->stream), "bcast sync")) return -1;
  memcpy(data, c->h_buf, nbytes);
  return 0;
}

int ab2_grid_set_comm(ab2_grid* h, ab2_comm* c, int lo_rank, int hi_rank)
{
  if (c && (lo_rank < -1 || hi_rank < -1 || lo_rank >= c->nranks || hi_rank >= c->nranks || lo_rank == c->rank || hi_rank == c->rank)) {
    ab2_set_error("[ab2_grid_set_comm]: neighbour ranks must lie in the communicator (-1 = none)"); return -1;
  }
  if (c && c->device != h->device) { ab2_set_error("[ab2_grid_set_comm]: communicator and Grid live on different devices"); return -1; }
  h->comm = c; h->lo_rank = c ? lo_rank : -1; h->hi_rank = c ? hi_rank : -1;
  if (c && !h->d_gather) {
    void* p = nullptr;
    if (!ab2_check(cudaMalloc(&p, (size_t)c->nranks*SC_NRED*sizeof(unsigned long long)), "gather buffer")) return -1;
    h->d_gather = (unsigned long long*)p; h->allocs.push_back(p);
    if (!ab2_check(cudaHostAlloc(&p, (size_t)c->nranks*SC_NRED*sizeof(unsigned long long), cudaHostAllocDefault), "gather host buffer")) return -1;
    h->h_gather = (unsigned long long*)p;
  }
  return 0;
}

int ab2_exchange_x3(ab2_grid* h) { return comm_exchange_x3(h); }

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// ghost exchange (bvals_mhd.c:423-446): lower neighbour's top active layers -> my layers [0,4), upper neighbour's bottom
// active layers -> my layers [n3-4, n3).  Slabs span all of x1,x2 incl. ghosts, so the corners follow from the x1 -> x2
// -> x3 order exactly as in bvals_mhd.c:170-172.  With periodic x3 the slabs form a ring (:768-774, 814-820); on a
// 2-rank ring both messages travel between the same pair, and NCCL matches point-to-point calls by order, so every rank
// posts the upward transfers before the downward ones.
// ---------------------------------------------------------------------------------------------------------------------
int comm_exchange_x3(ab2_grid* h)
{
  ab2_comm* c = h->comm;
  if (!c || (h->lo_rank < 0 && h->hi_rank < 0)) return 0;
  const GridDev& g = h->g;
  const size_t slab = (size_t)NG*g.st[2];
  const int nvar = 5 + h->nscal;
  const long n3 = g.n[2];
  if (!nccl_ok(N.GroupStart(), "ncclGroupStart")) return -1;
  for (int v=0; v<nvar; v++) {
    double* U = g.U[v];
    if (h->hi_rank >= 0 && !nccl_ok(N.Send(U + (size_t)(n3-2*NG)*g.st[2], slab, ncclDouble, h->hi_rank, c->comm, h->stream), "ncclSend")) return -1;
    if (h->lo_rank >= 0 && !nccl_ok(N.Recv(U, slab, ncclDouble, h->lo_rank, c->comm, h->stream), "ncclRecv")) return -1;
  }
  for (int v=0; v<nvar; v++) {
    double* U = g.U[v];
    if (h->lo_rank >= 0 && !nccl_ok(N.Send(U + (size_t)NG*g.st[2], slab, ncclDouble, h->lo_rank, c->comm, h->stream), "ncclSend")) return -1;
    if (h->hi_rank >= 0 && !nccl_ok(N.Recv(U + (size_t)(n3-NG)*g.st[2], slab, ncclDouble, h->hi_rank, c->comm, h->stream), "ncclRecv")) return -1;
  }
  if (!nccl_ok(N.GroupEnd(), "ncclGroupEnd")) return -1;
  c->n_exchanges++;
  return 0;
}

// The loop scalars d_scal[SC_MAXV1 .. SC_MAXV1+SC_NRED) of every rank -> h_scal, combined: MAX for the three wave speeds
// and SC_MAX_DTI, MIN for dt_chem / dt_therm, SUM for the check_range count, OR for the error flags.  One all-gather,
// one D2H copy, one synchronisation; with a single rank just the copy.  Values are never NaN (the kernels' atomics
// drop NaN operands), so min/max are plain comparisons.
int comm_read_scalars(ab2_grid* h)
{
  ab2_comm* c = h->comm;
  unsigned long long* hs = h->h_scal;
  if (!c || c->nranks == 1) {
    if (!ab2_check(cudaMemcpyAsync(hs + SC_MAXV1, h->d_scal + SC_MAXV1, SC_NRED*sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream), "scalars D2H")) return -1;
    return ab2_check(cudaStreamSynchronize(h->stream), "scalars sync") ? 0 : -1;
  }
  if (!nccl_ok(N.AllGather(h->d_scal + SC_MAXV1, h->d_gather, SC_NRED, ncclUint64, c->comm, h->stream), "ncclAllGather")) return -1;
  if (!ab2_check(cudaMemcpyAsync(h->h_gather, h->d_gather, (size_t)c->nranks*SC_NRED*sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream), "scalars D2H")) return -1;
  if (!ab2_check(cudaStreamSynchronize(h->stream), "scalars sync")) return -1;
  auto as_d = [](unsigned long long b) { double v; memcpy(&v, &b, sizeof v); return v; };
  for (int s=0; s<SC_NRED; s++) {
    const int id = SC_MAXV1 + s;
    unsigned long long acc = h->h_gather[s];
    for (int r=1; r<c->nranks; r++) {
      const unsigned long long x = h->h_gather[(size_t)r*SC_NRED + s];
      if (id == SC_COUNT) acc += x;
      else if (id == SC_ERR) acc |= x;
      else if (id == SC_DT_CHEM || id == SC_DT_THERM) { if (as_d(x) < as_d(acc)) acc = x; }
      else { if (as_d(x) > as_d(acc)) acc = x; }
    }
    hs[id] = acc;
  }
  c->n_reductions++;
  return 0;
}
