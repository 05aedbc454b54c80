/* libvgm_b200 -- C ABI of the B200-native variational gradient-matching hot path.
 *
 * The reference (ngorbach/Variational_Gradient_Matching_for_Dynamical_Systems) has no FFI:
 * its boundary is the Python function surface of VGM_modules/.  Each entry point below is
 * what a Python (ctypes) binding for one reference site calls; the site it replaces is
 * cited as file:line relative to the reference's VGM_modules/ directory.
 *
 * Conventions
 *   - every function returns VGM_OK (0) or a negative VGM_E* code; vgm_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - pointers are DEVICE pointers unless the name ends in _host; all matrices are FP64,
 *     row-major; state arrays are [n_states][ld] with one state's trajectory contiguous
 *     (ld >= T, ld even, base 16-byte aligned);
 *   - nothing allocates persistent device memory: scratch comes in through explicit
 *     workspace pointers sized by the matching *_ws_bytes() function;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); only the functions
 *     documented as synchronising (they return host scalars) wait for the device;
 *   - there is no CPU fallback anywhere in this library;
 *   - ONE PROCESS DRIVES ONE GPU: the library keeps a few process-wide resources (internal streams and events,
 *     cached CUDA graphs of the correction sequence, kernel attributes) that belong to the device it was first
 *     used on; the sweep / iteration entry points return VGM_EINVAL when called with another current device, and
 *     they are not re-entrant (call them from one thread at a time).  vgm_shutdown() drops the cached graphs.
 */
#ifndef VGM_B200_H
#define VGM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vgm_stream_t; /* cudaStream_t */

enum {
  VGM_OK = 0,
  VGM_EINVAL = -1,   /* bad argument (shape, alignment, null pointer)          */
  VGM_ECUDA = -2,    /* CUDA runtime error                                      */
  VGM_ENOTCONV = -3, /* iterative solve hit max_iter before reaching tolerance  */
  VGM_ENOTSPD = -4,  /* Cholesky pivot <= 0                                     */
  VGM_EDRIVER = -5,  /* CUDA driver API / module loading failure                */
  VGM_ENCCL = -6     /* NCCL not loadable, or an NCCL call failed (vgm_comm_*)  */
};

const char* vgm_last_error(void);
int vgm_version(void);
int vgm_shutdown(void); /* drop cached CUDA graphs (they point into caller workspaces) and the device binding */

/* ------------------------------------------------------------------------------------
 * G1  GP_regression.kernel_function, GP_regression.py:160-260
 * Fused build of C, dC = dk/dt0, ddC = d2k/dt0dt1 on t (T x T, leading dimension ld),
 * C_obs on t2 (T2 x T2, ld2) and C_state_obs (T x T2, ld2).  Any output pointer may be null.
 * kernel_type: VGM_KERNEL_RBF  param_host = [lengthscale, sigma]         (GP_regression.py:176-179)
 *              VGM_KERNEL_RQ   param_host = [alpha, lengthscale, sigma]  (GP_regression.py:205-209)
 *              VGM_KERNEL_PERIODIC         reads param[0], param[1]       (:181-184)
 *              VGM_KERNEL_LOCALLY_PERIODIC reads param[1], param[2]       (:185-188)
 *              VGM_KERNEL_RBF_LIN          reads param[1..4]              (:189-192)
 *              VGM_KERNEL_SIGMOID          reads param[1], param[2]       (:193-195)
 *              VGM_KERNEL_SIGMOID2         reads param[0], param[1]       (:196-198)
 *              VGM_KERNEL_EXP_SIN_SQUARED  reads param[0], param[1]       (:199-204)
 *              VGM_KERNEL_MATERN           param = [lengthscale, nu, ...], nu in {0.5, 1.5, 2.5} (:210-223; the
 *                                          reference's general-nu branch cannot run)
 * The result is scaled by sigma^2 = param[-1]^2 (GP_regression.py:232).  The three types built on
 * sqrt((t0-t1)^2) (periodic, locally periodic, exp_sin_squared) have NaN derivative grids on t0 == t1, exactly
 * as the reference's lambdified derivatives do.
 */
enum {
  VGM_KERNEL_RBF = 0, VGM_KERNEL_RQ = 1, VGM_KERNEL_PERIODIC = 2, VGM_KERNEL_LOCALLY_PERIODIC = 3,
  VGM_KERNEL_RBF_LIN = 4, VGM_KERNEL_SIGMOID = 5, VGM_KERNEL_SIGMOID2 = 6, VGM_KERNEL_EXP_SIN_SQUARED = 7,
  VGM_KERNEL_MATERN = 8
};
int vgm_kernel_build(int kernel_type, const double* param_host, int n_param,
                     const double* t, int T, const double* t2, int T2, int ld, int ld2,
                     double* C, double* dC, double* ddC, double* Cobs, double* Cso,
                     vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Dense FP64 tensor-core GEMM, C[n][m] = alpha * sum_k A[n][k] B[m][k] + epilogue.
 * Replaces every np.dot with a T x T operator on the hot path
 * (proxies_for_ode_parameters_and_states.py:76,225,231,232; GP_regression.py:306,317).
 * Epilogue: C = alpha*acc + beta*C0 + g1*(U1 ? U1 : 1).V1 + g2*(U2 ? U2 : 1).V2 (elementwise,
 * operands laid out like C); optional per-row partial dot products
 * dot_out[row*dot_ld + column_tile] = sum_{m in tile} dotw[row][m] * C[row][m].
 * dotw == C requests the squared row norms of the result.
 * rowsA / rowsC (optional) map logical row n to a physical row of A / of C and of every
 * epilogue operand; n_rows_dev (optional) is a device-side row count (<= N).
 * band > 0 declares B banded: B[m][k] is treated as zero for |m - k| > band and the k range outside
 * the band is skipped (vgm_band_halfwidth measures it; band <= 0 means dense).
 */
typedef struct {
  const double* A;
  const double* B;
  double* C;
  int lda, ldb, ldc;
  int N, M, K;
  const int* rowsA;
  const int* rowsC;
  const int* n_rows_dev;
  double alpha, beta;
  const double* C0;
  const double* U1;
  const double* V1;
  double g1;
  const double* U2;
  const double* V2;
  double g2;
  const double* dotw;
  double* dot_out;
  int dot_ld;
  int band;
} vgm_gemm_args;
int vgm_dgemm_nt(const vgm_gemm_args* args, vgm_stream_t stream);
int vgm_dgemm_set_variant(int variant); /* tile configuration of the large-N GEMM: 12 (default) = 64x64 x 4 CTAs/SM,
                                           2 = 128x64 x 2, 0 = 128x128 x 1 */
int vgm_dgemm_dot_tiles(int M, int N); /* number of column tiles written per row of dot_out */
/* Smallest b such that |B[i][j]| <= rel_threshold * max|B| whenever |i - j| > b (synchronises the stream).
 * With rel_threshold = 2^-64 the dropped part of any product is below half an ulp of the kept part's
 * normwise rounding error, i.e. the banded GEMM is FP64-exact in the backward sense; the engine uses 2^-55, where
 * the dropped part of a row of geometrically decaying operators stays below two units of round-off of
 * max|B| max|x| (less than the rounding error of the kept part's accumulation). */
int vgm_band_halfwidth(const double* B, int rows, int cols, int ld, double rel_threshold, int* band_host,
                       vgm_stream_t stream);
/* Per-launch CUDA-event timing of the kernel families (used by bench.py for the roofline figures):
 * family 0 = FP64 DMMA GEMM, 1 = tcgen05 BF16 GEMM, 2 = other solver vector kernels (init / finish / prep),
 * 3 = coefficient assembly + theta statistics, 4 = ir_update_kernel, 5 = ir_dir_kernel. */
enum { VGM_PROF_DGEMM = 0, VGM_PROF_TCGEMM = 1, VGM_PROF_VEC = 2, VGM_PROF_ASSEMBLE = 3, VGM_PROF_IR_UPDATE = 4,
       VGM_PROF_IR_DIR = 5 };
int vgm_profile_enable(int on);
int vgm_profile_collect(int family, double* flops, double* bytes, double* ms, long long* launches); /* synchronises */
int vgm_transpose(const double* in, int ld_in, double* out, int ld_out, int rows, int cols, vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * G2  GP_regression.kernel_function, GP_regression.py:306,317
 *   D = dC C^-1  (reference: LU solve on the transposes), A = ddC - D dC^T.
 * Cholesky route: C = L L^T, C^-1 = L^-T L^-1, then two DMMA GEMMs.  `batch` hyper-parameter
 * sets are processed back to back; matrices of set b start at base + b*T*ld.
 * Also exposes the SPD inverse used for the shared-operator preconditioner.
 * Returns VGM_ENOTSPD (after a device sync) if a pivot is not positive.
 */
size_t vgm_gm_operators_ws_bytes(int T, int ld);
int vgm_gm_operators(const double* C, const double* dC, const double* ddC, int T, int ld, int batch,
                     double* D, double* A, double* Cinv, void* ws, size_t ws_bytes, vgm_stream_t stream);
/* Same operators through LU with partial pivoting (the reference's own route, np.linalg.solve at :306) for a C
 * that is not positive definite -- the 'sigmoid2' kernel and the reference's Matern 1.5 / 2.5 on the squared
 * distance are indefinite.  One CTA factorises; use it when vgm_gm_operators returns VGM_ENOTSPD.  Same workspace
 * size.  Returns VGM_ENOTSPD if a column has no non-zero pivot. */
int vgm_gm_operators_lu(const double* C, const double* dC, const double* ddC, int T, int ld, double* D, double* A,
                        void* ws, size_t ws_bytes, vgm_stream_t stream);
int vgm_spd_inverse(const double* S, int T, int ld, double* Sinv, void* ws, size_t ws_bytes, vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * F1  GP_regression.fitting_state_observations, GP_regression.py:33-51 (the initial state proxy)
 *   for observed state i (in column order):  cov_obs += sigma_i^2 I   (the reference REASSIGNS cov_obs inside its
 *   loop, :42, so state i sees the noise of all states before it);  mean_i = Cso solve(cov_obs, y_i)        (:43)
 *   pred_cov = cov - Cso solve(cov_obs, Cso^T) with the final accumulated cov_obs                            (:45)
 * Cholesky solves (no explicit inverse) + DMMA GEMMs.  Cobs [T2][ld2], Cso [T][ld2], C [T][ld] (may be null when
 * pred_cov is null), Y [n_obs][ld2] (one observed state's series per row), obs_var_host [n_obs] = sigma_i^2 on the
 * HOST; outputs mean [n_obs][ld], pred_cov [T][ld] (optional).  Synchronises the stream (pivot checks);
 * VGM_ENOTSPD if cov_obs is not numerically positive definite. */
size_t vgm_gp_fit_ws_bytes(int T, int T2, int ld2);
int vgm_gp_fit(const double* Cobs, const double* Cso, const double* C, int T, int T2, int ld, int ld2, const double* Y,
               const double* obs_var_host, int n_obs, double* mean, double* pred_cov, void* ws, size_t ws_bytes,
               vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * G3/G4  rewrite_odes_as_local_linear_combinations.py:28-58,76-124
 * A "program" holds the model-specific coefficient kernels (B_k, b_k per ODE; R_uk, r_uk per
 * state x ODE) generated from the SymPy rewrite as CUDA device functions.  Programs are
 * either compiled ahead of time into this library (Lorenz96) or loaded from a cubin that
 * the Python code generator produced with NVRTC for sm_100a.
 */
typedef struct vgm_program vgm_program;
int vgm_program_builtin(const char* name, vgm_program** out); /* "lorenz96" */
int vgm_program_load(const void* cubin_host, size_t size, int n_param, int arity, vgm_program** out);
int vgm_program_free(vgm_program* prog);
int vgm_program_n_param(const vgm_program* prog);

/* ------------------------------------------------------------------------------------
 * P2/P3  proxy_for_ode_parameters, proxies_for_ode_parameters_and_states.py:58-93
 * Sufficient statistics  m = sum_k B_k^T (D x_k - b_k),  S = sum_k B_k^T B_k  over n_odes
 * ODE rows.  ODE row j belongs to structural class ode_class[j] and reads the states
 * ode_idx[j*arity .. j*arity+arity) (rows of X); its own state (row of DX) is ode_state[j].
 * Output stats = [m (p) | S (p*p, row-major)] on the device (so NCCL can all-reduce it in
 * place); vgm_param_solve solves S theta = m (constraints=None, proxies...py:93).
 */
size_t vgm_param_stats_ws_bytes(int n_odes, int T, int n_param);
int vgm_param_stats(const vgm_program* prog, const double* X, const double* DX, int ld, int T,
                    int n_odes, const int* ode_class, const int* ode_idx, const int* ode_state,
                    double* stats, void* ws, size_t ws_bytes, vgm_stream_t stream);
int vgm_param_solve(const double* stats, int n_param, double* theta, vgm_stream_t stream);

/* 'minimizer' parameter update (opt-in beyond the closed-form path; proxies...py:114-128 hands the squared
 * gradient-matching loss of :379-392 to scipy.optimize.minimize).  One evaluation on the device:
 *   out[0]     = sum_{k,t} (f_k(x(t), theta) - (D x_k)(t))^2
 *   out[1 + i] = 2 sum_{k,t} (f_k - D x_k) d f_k / d theta_i          (the L2 penalty alpha |theta|^2 is the caller's)
 * with f_k and its parameter gradient from the generated program, so models that are not linear in a parameter
 * (Protein Transduction's K_m) have a GPU path too.  theta, out: device.  The optimiser stays SciPy's on the host. */
size_t vgm_param_loss_ws_bytes(int n_odes, int n_param);
int vgm_param_loss(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int n_odes,
                   const int* ode_class, const int* ode_idx, const int* ode_state, const double* theta, double* out,
                   void* ws, size_t ws_bytes, vgm_stream_t stream);

/* Opt-in parameter "covariance"  cov = sum_k B_k^T A B_k  (p x p, row-major, device), the quantity the
 * reference computes at proxies...py:79 and never returns.  A = eps_cov [T][ld].  The B rows are
 * materialised in chunks of ODE rows, multiplied by A with the DMMA GEMM and reduced deterministically. */
size_t vgm_param_cov_ws_bytes(int n_odes, int ld, int n_param);
int vgm_param_cov(const vgm_program* prog, const double* X, const double* A, int ld, int T, int n_odes,
                  const int* ode_class, const int* ode_idx, double* cov, void* ws, size_t ws_bytes,
                  vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * P4  proxy_for_ind_states (optimizer='analytical'), proxies...py:185-262
 * The sweep plan (index tables on the device, built once by the host from state_couplings).
 * Hidden "slot" i is the i-th hidden state in the reference's update order.
 */
typedef struct {
  int T, ld;                 /* time points, leading dimension of every [rows][ld] array     */
  int n_hidden;              /* K_h                                                          */
  const int* slot_state;     /* [n_hidden] state index (row of X) of each slot               */
  const int* pair_ptr;       /* [n_hidden+1] CSR over (slot, coupled ODE) pairs, c(u) order  */
  const int* pair_ode;       /* [n_pairs] ODE row j (index into ode_idx / ode_state)         */
  const int* pair_code;      /* [n_pairs] generated-function selector (class, position)      */
  const int* pair_live;      /* [n_pairs] -1: D.x_k from the sweep-start snapshot; >=0: slot
                                in the live buffer (state already updated in this sweep,
                                proxies...py:225 reads the live column)                      */
  const int* pair_self;      /* [n_pairs] 1 if ODE k == u (proxies...py:228-232)             */
  const int* ode_idx;        /* [n_odes*arity] gather table (shared with vgm_param_stats)    */
  const int* ode_state;      /* [n_odes]                                                     */
  const int* slot_has_self;  /* [n_hidden] 0 => u not in c(u): reference doubles diag (:201,:245) */
  int n_levels;              /* dependency wavefronts that reproduce the sequential order    */
  const int* level_ptr_host; /* [n_levels+1] HOST array                                      */
  const int* level_slots;    /* [n_hidden] slots grouped by level (device)                   */
  int n_live;                /* states whose updated D.x is read later in the same sweep     */
  const int* live_state;     /* [n_live] state index of each live slot (device)              */
  const int* live_level_ptr_host; /* [n_levels+1] HOST: live slots grouped by producing level */
  int n_live_pairs;          /* pairs with pair_live >= 0                                    */
  const int* live_pair_slot; /* [n_live_pairs] hidden slot of each live pair, grouped by the level of the slot */
  const int* live_pair_src;  /* [n_live_pairs] live-buffer slot it reads                     */
  const int* live_pair_level_ptr_host; /* [n_levels+1] HOST                                  */
  int ruu_const;             /* 1 => R_uu is the same constant for every hidden state        */
  double ruu_value;          /*      that constant (Lorenz96: -1)                            */
  /* Optional overlap of a tiny tail with the bulk (vgm_state_sweep, vgm_state_sweep_levels).  When every level after
   * the first holds a handful of systems (Lorenz96 with every second state observed: the wrap-around state alone,
   * which reads the updated x of the first hidden state live, proxies...py:225) the chain "early level-0 systems ->
   * level 1 -> ..." is solved on a high-priority side stream, one single-CTA launch per link, while the other
   * level-0 systems run as one batch.  n_early = 0 (or null pointers) disables it; levels then run one after the
   * other.  At most 8 early systems, 8 tail systems, 8 levels. */
  int n_early;               /* level-0 slots whose updated x a later level reads live       */
  const int* early_slots;    /* [n_early] device                                             */
  const int* bulk_slots;     /* [level_ptr_host[1] - n_early] device: the other level-0 slots */
} vgm_sweep_plan;

typedef struct {
  const double* D;   /* dC C^-1, T x ld                                 */
  const double* DT;  /* its transpose                                   */
  const double* M;   /* (c I - D)^T (c I - D) if plan.ruu_const else null */
  const double* G;   /* optional preconditioner (M + shift I)^-1, else null */
  const void* G_bf16;   /* optional BF16 copy of G (vgm_convert_bf16), [T][ld]                                  */
  const void* M_bf16x2; /* optional BF16 split of M (vgm_split_bf16), [2][T][ld]: hi = bf16(M), lo = bf16(M - hi).
                           With G_bf16 it enables the mixed-precision solver: FP64 residuals on the DMMA pipe,
                           corrections by FP32 PCG whose operator products run on the tcgen05 tensor cores  */
  double G_shift;       /* the shift G was built with                                                          */
  double precond_c;     /* > 0: the preconditioner is S G S with S = diag sqrt((G_shift + c) / (Lambda + c))    */
  int band_D, band_M;   /* half band widths of D (and D^T) and M at the engine's FP64 threshold 2^-55
                           (vgm_band_halfwidth); <= 0: dense                                                  */
  int band_M_lp, band_G; /* half band widths used on the low-precision side: M at ~2^-20 (the bf16-split product
                           resolves 2^-16), G at ~2^-7 (a preconditioner only has to stay SPD); <= 0: dense   */
  const double* DtD;    /* optional D^T D, [T][ld]: with T <= 160 it enables the direct per-CTA solver for systems
                           whose R_uu varies in time (each CTA forms D^T D - diag(R) D - D^T diag(R) +
                           diag(R^2 + Lambda) in shared memory and factorises it); constant R_uu uses M          */
  const double* obs_rhs; /* optional observation term (opt-in; the statement the reference keeps commented out at
                           proxies...py:265-266, active in the MATLAB original, proxy_for_ind_states.m:92-93):
                           x_u = solve(local_scaling + Q, local_mean + Q mu_u) with Q = state_pred_inv_cov.  The
                           caller adds Q to M (or to DtD) and passes obs_rhs[slot] = Q mu_u, [n_hidden][ld]      */
} vgm_operators;
int vgm_convert_bf16(const double* in, void* out_bf16, int rows, int ld, vgm_stream_t stream);
int vgm_split_bf16(const double* in, void* out_bf16x2, int rows, int ld, vgm_stream_t stream);
/* tcgen05/TMEM/TMA BF16 GEMM (FP32 accumulation in tensor memory), raw FP32 output:
 *   out[j][m] = sum_k A0[j][k] B0[m][k]  (+ A1[j][k] B0[m][k] + A0[j][k] B1[m][k]  if A1 and B1 are given:
 *   the three leading terms of the product of two bf16-split operands, ~2^-16 relative accuracy).
 * A*: [n_rows][lda] bf16, B*: [M][ldb] bf16, out: [n_rows][ldo] fp32; band as in vgm_gemm_args. */
int vgm_tc_gemm_bf16(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1, int M, int K,
                     int ldb, int band, float* out, int ldo, vgm_stream_t stream);
/* The same product with the FP32 accumulators rounded to bf16 in the epilogue, out: [n_rows][ldo] bf16 (half the
 * output bytes; the solver stores the preconditioner product z' = (s o r) G^T this way).  The output moves in
 * 16-byte pieces: columns M .. round_up(M, 8) - 1 of a row are overwritten with zeros when M is not a multiple of 8. */
int vgm_tc_gemm_bf16_out(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1, int M, int K,
                         int ldb, int band, void* out_bf16, int ldo, vgm_stream_t stream);
/* ... and, from the same epilogue, the per-row partial dots of the ROUNDED output with a bf16 weight matrix
 * W [n_rows][ldw]:  dots[j][ct] = sum over the 128 columns m of column tile ct of bf16(out[j][m]) * W[j][m],
 * ct < ceil(M / 128) <= ld_dots; M % 8 == 0.  The solver gets r.z of the preconditioned CG this way (W = the A operand
 * s o r, an L2 hit), so the direction update that follows is a single streaming pass. */
int vgm_tc_gemm_bf16_out_dots(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1, int M,
                              int K, int ldb, int band, void* out_bf16, int ldo, const void* W, int ldw, float* dots,
                              int ld_dots, vgm_stream_t stream);

typedef struct {
  double tol;        /* relative residual ||r|| <= tol ||rhs|| per state (default 1e-13)                     */
  int max_iter;      /* FP64 PCG: iteration cap (default 4 T).  The mixed-precision solver allows 12 corrections; systems
                        that stagnate (too ill-conditioned for FP32 / bf16-split corrections) are finished by FP64 PCG */
  int check_every;   /* FP64 PCG: host polls the device-side active count every this many iterations         */
  int inner_iters;   /* mixed precision: FP32 PCG steps per correction (default 10)                           */
  int optimistic_passes; /* mixed precision: corrections enqueued before the first convergence poll (default 4;
                            < 0: poll after every pass).  Kernels read the active count on the device, so passes
                            over converged rows are free; fewer polls keep the queue full on small problems  */
  int small_level_solver; /* 1: dependency levels of <= 8 systems are solved by one CTA per system (FP32 banded Cholesky
                            + FP64 refinement in a single launch, csrc/small_solve.cu) instead of the batched solver.
                            Off by default: at T=1000 it takes 1.4 ms against 1.1 ms for the ~190 tiny launches  */
  int dense_short_grids;  /* 0 (default): systems with T <= 160 are formed and factorised directly, one CTA per system
                             (FP64 Cholesky in shared memory), falling back to the iterative solvers on a non-positive
                             pivot; -1: never */
} vgm_solver_opts;

typedef struct {
  int iterations;    /* CG iterations executed (max over states)    */
  int unconverged;   /* states that hit max_iter                    */
  int gemm_calls;    /* DGEMM launches inside the sweep             */
  int kernel_launches;
} vgm_sweep_info;

size_t vgm_state_sweep_ws_bytes(const vgm_sweep_plan* plan);
/* X is updated in place (hidden rows only).  DX0 must hold D.x_k of the sweep-start X for every
 * state (it is what vgm_param_stats consumed).  theta_star: device pointer to the p parameter
 * values plugged into R_uk, r_uk (the reference hard-codes [8], proxies...py:211).
 * Synchronises the stream (convergence polling). */
int vgm_state_sweep(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops,
                    const vgm_solver_opts* opts, double* X, const double* DX0, const double* theta_star,
                    void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream);

/* The same sweep in two phases, for callers that must exchange halo rows of X between
 * dependency levels (multi-GPU state-block partition): prepare once, then call
 * vgm_state_sweep_level for level = 0 .. n_levels-1 in order.  Before level L > 0 the rows of X
 * named by plan.live_state for producer level L-1 must hold their updated values (locally
 * solved, or received from the owning rank).  info accumulates across the calls. */
int vgm_state_sweep_prepare(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops,
                            const double* X, const double* DX0, const double* theta_star, void* ws, size_t ws_bytes,
                            vgm_sweep_info* info, vgm_stream_t stream);
int vgm_state_sweep_level(const vgm_sweep_plan* plan, const vgm_operators* ops, const vgm_solver_opts* opts,
                          double* X, int level, void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream);
/* Every level in order after vgm_state_sweep_prepare (what vgm_state_sweep does), including the tail overlap
 * described in vgm_sweep_plan.  For callers that own all live sources of their systems (a state-block partition
 * whose cuts no live read crosses) and so need no exchange between levels. */
int vgm_state_sweep_levels(const vgm_sweep_plan* plan, const vgm_operators* ops, const vgm_solver_opts* opts,
                           double* X, void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream);

/* One building block of the sweep, exposed for tests: snapshot coefficients
 * Lambda_u = sum_{k!=u} R_uk^2, rhs_u (without the -B_uu^T r_uu term and without live terms),
 * r_uu, R_uu and the R_uk of live pairs (all [n_hidden or n_live_pairs][ld]). */
int vgm_state_assemble(const vgm_program* prog, const vgm_sweep_plan* plan, const double* X, const double* DX0,
                       const double* theta_star, double* Lambda, double* rhs, double* ruu, double* Ruu,
                       double* Rlive, vgm_stream_t stream);

/* Batched (preconditioned) conjugate gradients on rows:  (Op + diag(Lambda_i)) x_i = rhs_i with
 * Op = M (shared, one GEMM per iteration) or Op_i = (diag(Ruu_i) - D)^T (diag(Ruu_i) - D)
 * (matrix-free, two GEMMs).  Ruu / Lambda / rhs are [n_hidden][ld] indexed by hidden slot;
 * rows[] lists the n_rows slots to solve; x lives in X rows slot_state[slot] and its current
 * content is the initial guess.  Synchronises the stream. */
size_t vgm_pcg_ws_bytes(int n_hidden, int T, int ld);
int vgm_pcg_solve(const vgm_operators* ops, int ruu_const, const double* Ruu, const double* Lambda, const double* rhs,
                  double* X, const int* slot_state, int n_hidden, const int* rows, int n_rows, int T, int ld,
                  const vgm_solver_opts* opts, void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * P5  state "covariance" proxies...py:190,239 (opt-in; the reference computes and drops it)
 *   Cov_u = sum_{k in c(u)} B_uk^T A B_uk  for one hidden slot, T x ld output.
 */
size_t vgm_state_cov_ws_bytes(int T, int ld);
int vgm_state_cov(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops, const double* A,
                  const double* X, const double* theta_star, int slot, double* cov, void* ws, size_t ws_bytes,
                  vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole coordinate-ascent iteration with HOST buffers (the end-to-end path):
 *   H2D X -> D.X GEMM -> theta statistics + solve -> state sweep -> D2H X, theta.
 * X_host_in / X_host_out are [n_states][ld] pinned or pageable host buffers (may alias).
 */
typedef struct {
  int n_states, n_odes, n_param;
  const int* ode_class;
  const int* ode_idx;
  const int* ode_state;
  int theta_mode; /* 0: plug theta_literal into R,r (reference, proxies...py:211); 1: plug the fresh theta */
  double theta_literal[8];
  /* vgm_iteration_host only: when out_n_rows > 0 the updated states are the rows out_row0 + i * out_row_stride
   * (i < out_n_rows) and only those are copied back to the host (one strided copy); the rows of states that the
   * sweep does not touch (observed + clamped, proxies...py:175-179) are then left as they are in X_host_out. */
  int out_row0, out_row_stride, out_n_rows;
  /* vgm_iteration_host only, optional pipelining of the host<->device copies with the compute (pipe_chunks >= 2).
   * The states are cut into pipe_chunks contiguous ranges [pipe_state_ptr_host[c], ..[c+1]) (boundaries multiples
   * of 8); range c is copied in, then STAGED (D.X, theta partial sums, coefficients, rhs GEMM for its hidden slots
   * [pipe_slot_ptr_host[c], ..[c+1])), then the systems of dependency level 0 whose state lies in
   * [state_ptr[c] - halo_x, state_ptr[c+1] - halo_x) are SOLVED (positions [pipe_row_ptr_host[c], ..[c+1]) of the
   * level, hidden slots [pipe_out_slot_ptr_host[c], ..[c+1])) and copied out while the next range is in flight.
   * The shift by the halo keeps every coefficient on the sweep-start snapshot (proxies...py:191).  Row/slot ranges
   * start at the "seam" (states < halo_x, which the last range reads cyclically): rows [0, row_ptr[0]) are solved
   * last; every slot of a later level lies in the last range.  Coefficients of a state read X rows within
   * +-pipe_halo_x (cyclically) and D.X rows within +-pipe_halo_dx.  Requires theta_mode 0, ODE row j == state j,
   * hidden slots in ascending state order (VGMEngine.iteration_host checks all of it and falls back otherwise). */
  int pipe_chunks, pipe_halo_x, pipe_halo_dx;
  const int* pipe_state_ptr_host;
  const int* pipe_slot_ptr_host;
  const int* pipe_row_ptr_host;
  const int* pipe_out_slot_ptr_host;
  const int* pipe_last_rows; /* optional DEVICE array: the level-0 slots of the last solve range followed by the seam's
                                (row_ptr[n] - row_ptr[n-1] + row_ptr[0] entries): one solve instead of two */
} vgm_model_tables;
size_t vgm_iteration_ws_bytes(const vgm_model_tables* model, const vgm_sweep_plan* plan);
int vgm_iteration_device(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                         const vgm_operators* ops, const vgm_solver_opts* opts, double* X, double* DX,
                         double* theta, double* stats, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                         vgm_stream_t stream);
int vgm_iteration_host(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                       const vgm_operators* ops, const vgm_solver_opts* opts, const double* X_host_in,
                       double* X_host_out, double* theta_host, double* X_dev, double* DX_dev, double* theta_dev, double* stats_dev,
                       void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Synthetic Lorenz96 trajectories on the device (data generation; replaces setup_simulation's odeint,
 * simulate_state_dynamics.py:79-92, for benchmarks -- not a parity target, LSODA is third-party).
 * Classical RK4 with step h on the ring f_k = (x_{k+1} - x_{k-2}) x_{k-1} - x_k + alpha (Lorenz96_ODEs.txt:1-10):
 * x0 [K] is the initial state, n_spinup steps are discarded, then X[k][i] = x_k(i * n_sub * h) for i < n_grid
 * (state-major, leading dimension ld).  n_sub <= 8; scratch holds 2 K doubles.  One launch per grid interval. */
int vgm_lorenz96_rk4(const double* x0, int K, double alpha, double h, int n_sub, int n_spinup, int n_grid, double* X,
                     int ld, double* scratch, vgm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU exchange steps, one process per GPU.  The path has two exchanges per iteration and these
 * are thin NCCL wrappers for exactly those; everything else a rank does is the single-GPU entry points above on its
 * own block of states (or trajectories).  NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the
 * process, else from the loader path): the library loads without it and these functions then return VGM_ENCCL.
 *
 *   vgm_comm_unique_id       rank 0 creates the 128-byte id and hands it to the other ranks out of band (file, MPI,
 *                            torch.distributed ... -- the same bytes an `ncclUniqueId` holds)
 *   vgm_comm_init            collective: every rank calls it with its rank on its current device (ncclCommInitRank)
 *   vgm_comm_allreduce_stats in-place sum over ranks of the parameter statistics [m | S] (p + p^2 doubles) that
 *                            vgm_param_stats wrote for this rank's ODE rows: the sum over k of
 *                            proxies_for_ode_parameters_and_states.py:64-79 (`local_mean += ...`, `local_scaling += ...`)
 *                            split over ranks; vgm_param_solve then runs redundantly on every rank
 *   vgm_comm_halo_exchange   snapshot halo of the state sweep (proxies...py:191, one copy of the states per sweep; the
 *                            coefficients of state u read its coupled neighbours, u-3 .. u+3 for Lorenz96): rank r
 *                            publishes rows pub_rows[0 .. n_pub_mine) of its X (n_pub = the maximum count over ranks,
 *                            identical on every rank), ONE all-gather, then X[halo_rows[i]] = gathered[halo_src[i]]
 *                            for i < n_halo with halo_src = owner_rank * n_pub + position in the owner's pub_rows.
 *                            pub_rows / halo_rows / halo_src are DEVICE int arrays; ws holds
 *                            vgm_comm_halo_ws_bytes(world, n_pub, ld) bytes.  n_pub == 0 (trajectory batches): no-op.
 * All calls enqueue on `stream` and return; ranks must issue them in the same order. */
#define VGM_COMM_ID_BYTES 128
typedef struct vgm_comm vgm_comm;
int vgm_comm_nccl_version(int* version_host);
int vgm_comm_unique_id(void* id_host);
int vgm_comm_init(const void* id_host, int rank, int world, vgm_comm** comm);
int vgm_comm_free(vgm_comm* comm);
int vgm_comm_rank(const vgm_comm* comm, int* rank_host, int* world_host);
int vgm_comm_allreduce_stats(vgm_comm* comm, double* stats, int n, vgm_stream_t stream);
size_t vgm_comm_halo_ws_bytes(int world, int n_pub, int ld);
int vgm_comm_halo_exchange(vgm_comm* comm, double* X, int ld, const int* pub_rows, int n_pub_mine, int n_pub,
                           const int* halo_rows, const int* halo_src, int n_halo, void* ws, size_t ws_bytes,
                           vgm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VGM_B200_H */
