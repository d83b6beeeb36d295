// engine.cuh -- device-side grid descriptor shared by the kernels, and the host-side handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define AB2_NG 4                 // nghost, defs.h.in:129-140
#define AB2_NVAR 6               // d, M1, M2, M3, E, s0 (s0 unused when nscal == 0)

// Static-potential tables.  Entry = Phi evaluated by the HOST callback at exactly the argument the
// reference passes for cell (i,j,k) (integrate_3d_ctu.c:334-336, 1172-1215, 1468-1522, 1878-1933,
// 2746-2777): offsets per direction are 0, -dx ("M"), -dx/2 ("MH"), +dx/2 ("PH").
enum {
  PHI_C = 0,
  PHI_M0, PHI_M1, PHI_M2,          // (x_d - dx_d)
  PHI_MH0, PHI_MH1, PHI_MH2,       // (x_d - dx_d/2)
  PHI_PH0, PHI_PH1, PHI_PH2,       // (x_d + dx_d/2)
  PHI_MIX0,                        // 12 tables: MIX(d,t,s): -dx_d in d and (s? +:-) dx_t/2 in t, t != d
  PHI_NTAB = PHI_MIX0 + 12
};
// index of MIX(d,t,sign): d in 0..2, t = one of the two other directions in ascending order (slot 0/1), sign 0:-,1:+
__host__ __device__ inline int phi_mix(int d, int t, int plus)
{
  int slot = (t > d) ? (t - 1) : t;           // position of t among the directions != d  (0 or 1)
  return PHI_MIX0 + d*4 + slot*2 + plus;
}

#define AB2_RIEMANN_ROE 0
#define AB2_RIEMANN_HLLC 1
#define AB2_RIEMANN_ROE_NOH 2     // Roe, reference built without --enable-h-correction: etah = 0 in every solve

struct GridDev {
  int nx[3];                 // active zones
  int n[3];                  // with ghosts
  long st[3];                // element strides: st[0]=1, st[1]=n1, st[2]=n1*n2
  double dx[3];
  double Gamma, Gamma_1;
  int order;                 // lr_states(): 2 = PLM (lr_states_plm.c), 3 = PPM (lr_states_ppm.c), the reference's --with-order
  int riemann;               // fluxes() of the build being replaced: 0 Roe + H-correction (roe.c), 1 HLLC (hllc.c), 2 Roe without H-correction
  double* U[AB2_NVAR];       // conserved state, SoA [n3][n2][n1]
  const double* Uin[AB2_NVAR]; // what the integrator READS: U itself, or (x3-chunked scratch) a copy of the chunk's planes of U^n
  // CTU scratch.  Whole-grid arrays, or -- when the scratch is chunked along x3 -- arrays that hold only the planes of the
  // current chunk (+4 either side) with their base pointers shifted so that the cell index c of U addresses them too.
  double* Ul[3][AB2_NVAR];   // face states, PHYSICAL component order (d,M1,M2,M3,E,s), one set per face direction
  double* Ur[3][AB2_NVAR];
  double* F[3][AB2_NVAR];    // first-pass fluxes, physical component order
  double* Ff[3][AB2_NVAR];   // final fluxes (integrate_3d_ctu.c step 9; what step 11-12 and the SMR flux correction read)
  double* eta[3];            // only the unfused correct/final pair goes through these
  const double* phi[PHI_NTAB];
};

// scalars block in device memory (mirrored to pinned host memory on demand)
enum {
  SC_MAXV1 = 0, SC_MAXV2, SC_MAXV3,   // new_dt: max |v|+a per direction (bit patterns, atomicMax on u64)
  SC_DT_CHEM, SC_DT_THERM,            // ion: min over cells (atomicMin on u64 patterns)
  SC_MAX_DTI,                         // ion: compute_dt_hydro max (atomicMax)
  SC_COUNT,                           // ion: check_range cell count (u64)
  SC_ERR,                             // sticky error flags
  // sparse ion sub-cycle, written by k_ion_walk: reductions over the re-used chunks (folded into the pass scalars by the
  // first warp of the kernels that run over the task list), number of listed rays, number of re-used chunks
  SC_W_DTC, SC_W_DTT, SC_W_M1, SC_W_M2, SC_W_M3, SC_W_CNT, SC_NTASK, SC_W_NSKIP,
  SC_N = 16
};
#define SC_NRED 8                     // SC_MAXV1 .. SC_ERR: the scalars one cross-rank exchange combines (comm.cu)

struct IonParDev {
  double sigma_ph, m_H, mu, e_gamma, alpha_C, k_B, time_unit;
  double max_de_iter, max_de_therm_iter, max_dx_iter;
  double max_de_step, max_de_therm_step, max_dx_step;
  double tfloor, tceil;
  double min_area, d_nlo;
};

// Sub-cycle memo, one record per (ray, 32-cell chunk along x1) -- ion.cu "sparse sub-cycle".  Within one call of
// ion_radtransfer_3d a chunk whose cells no update touches (fully ionised ambient gas at the neutral floor, the bulk
// of an irradiated-planet box) sees the same inputs in every iteration, so its rates, its reductions and its
// outgoing flux are those of the previous iteration bit for bit: they are re-used instead of re-evaluated.
#define CH_RFIX      1u      // rates in memory were computed with every cell's sign state at its fixed point, no cell meeting the update's condition
#define CH_UINERT    2u      // the update after those rates left E and s of every cell of the chunk untouched (condition false, no floor fired)
#define CH_ALIVE_IN  4u      // the ray was still uncut where it entered the chunk
#define CH_ALIVE_OUT 8u      // ... where it left it
#define CH_STAMP_SHIFT 4     // flags >> 4: number of the rates pass (1, 2, ... within a call) that last EVALUATED the chunk
struct IonTask { int ray, ch0; double carry; int alive, pad; };   // evaluate `ray` from chunk ch0 on, entered with this flux / cut state
struct IonChunks {
  unsigned* flags;            // NULL: memo off (AB2_ION_SKIP=0), every chunk is evaluated in every iteration
  unsigned* uflags;           // CH_UINERT, written by k_ion_update (flags: the other bits, written by k_ion_rates)
  double* rec;                // 8 doubles per chunk: flux in, flux out, min dt_chem, min dt_therm | max |v_d|+a (d = 1..3), check_range count
  IonTask* tasks;             // rays with a chunk to evaluate, listed by k_ion_walk
  int nch;                    // chunks per ray = ceil(nx1 / 32)
};

struct IonDev {
  double *ph_rate, *edot, *nHdot, *e_init, *e_th_init, *x_init;
  int* sign_state;            // (sign_count << 2) | (last_sign + 1)
  double* edgeflux;           // [nx3+1][nx2+1][nx1+1], current
  double* edgeflux_alt;       // written by a fused update+rates launch; becomes current when the sub-cycle continues
  IonChunks ch;
};

typedef void (*ab2_reduce_fn_t)(double* vals, int n, int op, void* ctx);
struct ab2_comm;

// per-kernel CUDA-event timing (bench.py's roofline): ids index KERNEL_NAMES in api.cu
enum {
  KID_BVALS = 0, KID_MAX_SPEEDS, KID_FIRST_X1, KID_FIRST_X2, KID_FIRST_X3, KID_DHALF,
  KID_CORRECT_X1, KID_CORRECT_X2, KID_CORRECT_X3, KID_FINAL_X1, KID_FINAL_X2, KID_FINAL_X3, KID_UPDATE,
  KID_APPLY_FIXED, KID_ION_PREP, KID_ION_RATES, KID_ION_UPDATE, KID_BVALS_IONRAD, KID_TRANSPOSE,
  KID_SMR_RESTRICT, KID_SMR_FLUXCORR, KID_SMR_PROLONG, KID_ION_FUSED, KID_CFINAL, KID_CORRECT3,
  KID_ION_RATES_MEMO, KID_ION_UPDATE_MEMO, KID_ION_WALK, KID_ION_RATES_FIRST, KID_N
};
struct ProfPending { int id; cudaEvent_t a, b; };

struct ab2_grid {
  GridDev g;
  int nscal;
  int device;
  double xmin[3];              // GridS.MinX
  cudaStream_t stream;
  bool own_stream, own_U;
  std::vector<void*> allocs;
  std::vector<void*> phi_allocs;   // static-potential tables (replaced as a set by ab2_set_static_grav_pot)
  unsigned long long* d_scal;      // device scalars
  unsigned long long* h_scal;      // pinned host mirror
  bool has_grav;
  bool has_ion;
  IonParDev ionpar;
  IonDev ion;
  int maxiter;
  unsigned* ion_ch_flags;          // the memo's flag array (ion.ch.flags is NULL while AB2_ION_SKIP=0 turns the memo off)
  size_t ion_nchunks;
  int ion_last_ntask;              // the latest k_ion_walk: rays listed, chunks re-used (ab2_ion_memo_stats)
  long long ion_last_nskip;
  double tcoarse;                  // ionrad_3d.c:1028: dt_done of the last root-level radiation call
  double* scratch;                 // one allocation behind Ul/Ur/F/Ff/eta (and Uin, carry when chunked)
  size_t scratch_doubles;
  int chunk_planes;                // 0: scratch spans the whole Grid; else active x3 planes per chunk (scratch depth = that + 8)
  int fuse;                        // 1: k_cfinal (correct + eta + final fused), 0: k_correct / k_final pair
  int first_split;                 // pass A, PLM: bit d set = direction d as two kernels (k_plm_states + k_ppm_flux) instead of the fused one
  int correct3;                    // unfused pair: 1 = the three directions of pass C in one launch (k_correct3)
  // problem-file boundary functions (bvals_mhd_fun) in table form, one slot per face (ix1, ox1, ix2, ox2, ix3, ox3):
  // entry t sets U[v][dst[t]] = (src[v][t] >= 0) ? coef[v][t]*U[v][src[v][t]] : coef[v][t]
  long bc_n[6];
  long* bc_dst[6];
  long* bc_src[6];                 // [nvar][n]
  double* bc_coef[6];              // [nvar][n]
  long n_fixed;                    // Userwork_in_loop replay table
  long* d_fixed_cells;
  double* d_fixed_vals;
  size_t bytes_allocated;
  long long launches;              // kernels launched so far (bench.py's gpu_launches)
  bool prof_on;
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[KID_N];
  long long prof_cnt[KID_N];
  ab2_comm* comm;                  // NCCL communicator of the Grids that exchange / reduce with this one (comm.cu), or NULL
  int lo_rank, hi_rank;            // x3 neighbours in that communicator, -1 = none (physical boundary or single Grid)
  unsigned long long* d_gather;    // all-gathered loop scalars, nranks x SC_NRED
  unsigned long long* h_gather;    // pinned host mirror
};

// launch bookkeeping: counts every kernel; with profiling on, brackets it with events on the launching stream
cudaEvent_t ab2_prof_event(ab2_grid* h);
#define AB2_LAUNCH(h, id, ...) do { \
    cudaEvent_t _a = nullptr, _b = nullptr; \
    if ((h)->prof_on) { _a = ab2_prof_event(h); _b = ab2_prof_event(h); cudaEventRecord(_a, (h)->stream); } \
    __VA_ARGS__; \
    if ((h)->prof_on) { cudaEventRecord(_b, (h)->stream); (h)->prof_pending.push_back(ProfPending{(id), _a, _b}); } \
    (h)->launches++; \
  } while (0)

// error plumbing (api.cu)
void ab2_set_error(const std::string& s);
bool ab2_check(cudaError_t e, const char* what);

// comm.cu
int comm_exchange_x3(ab2_grid* h);
int comm_read_scalars(ab2_grid* h);
// api.cu: (re)allocate the CTU scratch for `planes` active x3 planes per chunk (0 = whole Grid)
int ab2_alloc_scratch(ab2_grid* h, int planes);
// hydro.cu
int hydro_bvals(ab2_grid* h, const int bc[6]);
int hydro_max_speeds(ab2_grid* h, double out[3]);
int hydro_integrate(ab2_grid* h, double dt);
int hydro_apply_fixed(ab2_grid* h);
// ion.cu
int ion_bvals_ionrad(ab2_grid* h, double flux_i);
int ion_radtransfer(ab2_grid* h, double flux_in, double cour_no, double* dt_io, int* niter, int* nchem, int* ntherm,
                    ab2_reduce_fn_t reduce, void* ctx);
int ion_memo_count(ab2_grid* h, long long out[3]);
int ion_radtransfer_fine(ab2_grid* h, double tcoarse, double* dt_io, int* niter, int* nchem, int* ntherm,
                         ab2_reduce_fn_t reduce, void* ctx);
