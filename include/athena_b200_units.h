/* athena_b200_units.h -- function-level entry points of libathena_b200_units.so: TEST HOOKS, not part of the interface
 * the reference driver binds (include/athena_b200.h).  They run single routines of csrc/physics.cuh / glibc_math.cuh on
 * host arrays so tests/test_gpu_parity.py can compare them with vectors from the reference's own objects. */
#ifndef ATHENA_B200_UNITS_H
#define ATHENA_B200_UNITS_H
#ifdef __cplusplus
extern "C" {
#endif

/* rows of 6 doubles, host arrays */
/* fluxes() of rsolvers/roe.c:59-336 (+ hlle.c fallback) on n independent interfaces */
int ab2_unit_roe_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                      int nscal, double gamma, const double* etah, double* F);
/* Cons1D_to_Prim1D / Prim1D_to_Cons1D / cfast (convert_var.c:389-499); cfast_out may be NULL */
int ab2_unit_convert(int n, const double* in, int nscal, double gamma, int to_prim, double* out, double* cfast_out);
/* a/b and sqrt(a) through the engine's branch-free IEEE refinement (csrc/physics.cuh Arith<true>, plain operators where
 * its range test rejects the operands) next to the plain operators; nrejected[2] = rejected divisions, square roots */
int ab2_unit_ieee(int n, const double* a, const double* b, double* q_fast, double* q_exact, double* s_fast, double* s_exact, int* nrejected);
/* exp(x[i]) and pow(y[i], x[i]) as the ion kernels evaluate them (csrc/glibc_math.cuh: the host glibc's algorithms) */
int ab2_unit_libm(int n, const double* x, const double* y, double* exp_x, double* pow_y_x);
/* lr_states() of reconstruction/lr_states_plm.c:62-374 on npencil pencils of n cells */
int ab2_unit_lr_states(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                       int il, int iu, double* Wl, double* Wr);
/* the same with steps 8-9 in the form an HLLE/HLLC/HLLD build compiles (lr_states_plm.c:257-293) */
int ab2_unit_lr_states_hll(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                           int il, int iu, double* Wl, double* Wr);
/* lr_states() of reconstruction/lr_states_ppm.c:67-651; hll != 0: steps 18-19 in the form an HLL-family build compiles */
int ab2_unit_lr_states_ppm(int npencil, int n, const double* W, int nscal, double gamma, double dt, double dx,
                           int il, int iu, int hll, double* Wl, double* Wr);
/* fluxes() of rsolvers/hllc.c:50-261 on n independent interfaces */
int ab2_unit_hllc_flux(int n, const double* Ul, const double* Ur, const double* Wl, const double* Wr,
                       int nscal, double gamma, double* F);

#ifdef __cplusplus
}
#endif
#endif
