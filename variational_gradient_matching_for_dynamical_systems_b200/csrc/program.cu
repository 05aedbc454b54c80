// Model "programs": the built-in Lorenz96 coefficient kernels and the loader for programs
// that the Python code generator compiled with NVRTC (G3/G4,
// rewrite_odes_as_local_linear_combinations.py:28-58,76-124), plus the generic parts of the
// theta update (P2/P3, proxies_for_ode_parameters_and_states.py:58-93).
#include <cuda.h>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"

// ---------------------------------------------------------------------------------------
// Built-in program: Lorenz96, f_k = (x_{k+1} - x_{k-2}) x_{k-1} - x_k + alpha  (Lorenz96_ODEs.txt:1-10)
// Gathered states in order of first appearance in the ODE line:
//   s0 = x_{k+1}, s1 = x_{k-2}, s2 = x_{k-1}, s3 = x_k.      code = position of x_u among them.
// ---------------------------------------------------------------------------------------
namespace lorenz96_prog {
#define VGM_PROG(name) vgm_lorenz96_##name
#define VGM_NPARAM 1
#define VGM_ARITY 4
__device__ __forceinline__ void vgm_theta_coeffs(int, const double (&s)[4], double (&B)[1], double& b) {
  B[0] = 1.0;                                  // d f_k / d alpha; rewrite...py:45
  b = s[0] * s[2] - s[1] * s[2] - s[3];        // f_k - alpha, expanded
}
__device__ __forceinline__ void vgm_state_coeffs(int code, const double* th, const double (&s)[4], double& R, double& r) {
  const double a = th[0];
  switch (code) {
    case 0:  R = s[2];        r = a - s[1] * s[2] - s[3]; break;          // u = k+1
    case 1:  R = -s[2];       r = a + s[0] * s[2] - s[3]; break;          // u = k-2
    case 2:  R = s[0] - s[1]; r = a - s[3]; break;                        // u = k-1
    default: R = -1.0;        r = a + s[0] * s[2] - s[1] * s[2]; break;   // u = k
  }
}
__device__ __forceinline__ void vgm_ode_eval(int, const double* th, const double (&s)[4], double& f, double (&df)[1]) {
  f = (s[0] - s[1]) * s[2] - s[3] + th[0];
  df[0] = 1.0;
}
#include "vgm_program_template.cuh"
#undef VGM_PROG
#undef VGM_NPARAM
#undef VGM_ARITY
}  // namespace lorenz96_prog

struct vgm_program {
  int n_param;
  int arity;
  bool builtin;
  cudaLibrary_t lib;
  cudaKernel_t k_stats;
  cudaKernel_t k_assemble;
  cudaKernel_t k_pair_R;
  cudaKernel_t k_param_B;
  cudaKernel_t k_param_loss;
};

namespace vgm {

typedef CUresult (*cuLaunchKernel_t)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                     CUstream, void**, void**);
static cuLaunchKernel_t g_cuLaunchKernel = nullptr;

static int driver_launch(cudaKernel_t k, unsigned grid, unsigned block, void** args, cudaStream_t s) {
  if (!g_cuLaunchKernel) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    VGM_CUDA_OK(cudaGetDriverEntryPoint("cuLaunchKernel", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      set_error("cuLaunchKernel entry point not available");
      return VGM_EDRIVER;
    }
    g_cuLaunchKernel = (cuLaunchKernel_t)fn;
  }
  CUresult r = g_cuLaunchKernel((CUfunction)k, grid, 1, 1, block, 1, 1, 0, (CUstream)s, args, nullptr);
  if (r != CUDA_SUCCESS) {
    set_error("cuLaunchKernel failed with CUresult %d", (int)r);
    return VGM_EDRIVER;
  }
  return VGM_OK;
}

static int driver_launch_grid(cudaKernel_t k, unsigned grid, unsigned block, void** args, cudaStream_t s) {
  return driver_launch(k, grid, block, args, s);
}

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);

static int prog_threads(int T) { return max(64, min(256, (T + 31) / 32 * 32)); }

int program_param_stats(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int n_odes,
                        const int* ode_class, const int* ode_idx, const int* ode_state, double* partial,
                        cudaStream_t s) {
  const int grid = ceil_div(n_odes, 8);
  const int block = prog_threads(T);
  ProfScope prof(PROF_ASSEMBLE, 0.0, 16.0 * n_odes * T, s);       // X and D.X once
  if (prog->builtin) {
    lorenz96_prog::vgm_lorenz96_param_stats<<<grid, block, 0, s>>>(X, DX, ld, T, n_odes, ode_class, ode_idx, ode_state,
                                                                   partial);
    VGM_LAUNCH_OK();
    return VGM_OK;
  }
  void* args[] = {(void*)&X, (void*)&DX, (void*)&ld, (void*)&T, (void*)&n_odes, (void*)&ode_class,
                  (void*)&ode_idx, (void*)&ode_state, (void*)&partial};
  return driver_launch(prog->k_stats, grid, block, args, s);
}

int program_param_loss(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int n_odes,
                       const int* ode_class, const int* ode_idx, const int* ode_state, const double* theta,
                       double* partial, cudaStream_t s) {
  const int grid = ceil_div(n_odes, 8);
  const int block = prog_threads(T);
  ProfScope prof(PROF_ASSEMBLE, 0.0, 16.0 * n_odes * T, s);
  if (prog->builtin) {
    lorenz96_prog::vgm_lorenz96_param_loss<<<grid, block, 0, s>>>(X, DX, ld, T, n_odes, ode_class, ode_idx, ode_state,
                                                                  theta, partial);
    VGM_LAUNCH_OK();
    return VGM_OK;
  }
  void* args[] = {(void*)&X, (void*)&DX, (void*)&ld, (void*)&T, (void*)&n_odes, (void*)&ode_class,
                  (void*)&ode_idx, (void*)&ode_state, (void*)&theta, (void*)&partial};
  return driver_launch(prog->k_param_loss, grid, block, args, s);
}

int program_param_B(const vgm_program* prog, const double* X, int ld, int T, int n_rows, const int* ode_class,
                    const int* ode_idx, double* Bout, cudaStream_t s) {
  const int block = prog_threads(T);
  if (prog->builtin) {
    lorenz96_prog::vgm_lorenz96_param_B<<<n_rows, block, 0, s>>>(X, ld, T, n_rows, ode_class, ode_idx, Bout);
    VGM_LAUNCH_OK();
    return VGM_OK;
  }
  void* args[] = {(void*)&X, (void*)&ld, (void*)&T, (void*)&n_rows, (void*)&ode_class, (void*)&ode_idx, (void*)&Bout};
  return driver_launch(prog->k_param_B, n_rows, block, args, s);
}

int program_state_assemble(const vgm_program* prog, const vgm_sweep_plan* pl, const double* X, const double* DX0,
                           const double* theta, double* Lambda, double* rhs, double* ruu, double* Ruu, double* Rlive,
                           cudaStream_t s) {
  if (pl->n_hidden == 0) return VGM_OK;
  const int grid = pl->n_hidden, block = prog_threads(pl->T);
  int ld = pl->ld, T = pl->T, nh = pl->n_hidden;
  // writes Lambda, rhs, r_uu (24 B) and reads ~arity+pairs rows of X / D.X through L1/L2 (~8 distinct rows: 64 B)
  ProfScope prof(PROF_ASSEMBLE, 0.0, 88.0 * nh * T, s);
  if (prog->builtin) {
    lorenz96_prog::vgm_lorenz96_state_assemble<<<grid, block, 0, s>>>(
        X, DX0, ld, T, nh, pl->pair_ptr, pl->pair_ode, pl->pair_code, pl->pair_live, pl->pair_self, pl->ode_idx,
        pl->ode_state, theta, Lambda, rhs, ruu, Ruu, Rlive);
    VGM_LAUNCH_OK();
    return VGM_OK;
  }
  const int *pp = pl->pair_ptr, *po = pl->pair_ode, *pc = pl->pair_code, *pv = pl->pair_live, *ps = pl->pair_self,
            *oi = pl->ode_idx, *os = pl->ode_state;
  void* args[] = {(void*)&X, (void*)&DX0, (void*)&ld, (void*)&T, (void*)&nh, (void*)&pp, (void*)&po, (void*)&pc,
                  (void*)&pv, (void*)&ps, (void*)&oi, (void*)&os, (void*)&theta, (void*)&Lambda, (void*)&rhs,
                  (void*)&ruu, (void*)&Ruu, (void*)&Rlive};
  return driver_launch(prog->k_assemble, grid, block, args, s);
}

int program_state_pair_R(const vgm_program* prog, const vgm_sweep_plan* pl, const double* X, const double* theta,
                         int slot, double* Rout, cudaStream_t s) {
  int ld = pl->ld, T = pl->T;
  const int block = prog_threads(T), grid = ceil_div(T, block);
  if (prog->builtin) {
    lorenz96_prog::vgm_lorenz96_state_pair_R<<<grid, block, 0, s>>>(X, ld, T, slot, pl->pair_ptr, pl->pair_ode,
                                                                    pl->pair_code, pl->ode_idx, theta, Rout);
    VGM_LAUNCH_OK();
    return VGM_OK;
  }
  const int *pp = pl->pair_ptr, *po = pl->pair_ode, *pc = pl->pair_code, *oi = pl->ode_idx;
  void* args[] = {(void*)&X, (void*)&ld, (void*)&T, (void*)&slot, (void*)&pp, (void*)&po, (void*)&pc, (void*)&oi,
                  (void*)&theta, (void*)&Rout};
  return driver_launch_grid(prog->k_pair_R, grid, block, args, s);
}

// Deterministic final reduction of the per-CTA partials -> stats = [m (p) | S (p x p)]
__global__ void __launch_bounds__(256) stats_reduce_kernel(const double* __restrict__ partial, int n_part, int p,
                                                           double* __restrict__ stats) {
  __shared__ double red[40];
  const int nstat = p + p * (p + 1) / 2;
  for (int q = 0; q < nstat; ++q) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) v += partial[(size_t)i * nstat + q];
    v = block_sum(v, red);
    if (threadIdx.x == 0) {
      if (q < p) {
        stats[q] = v;
      } else {
        // unpack the upper triangle (i <= l) into the full symmetric p x p matrix
        int rem = q - p, i = 0;
        while (rem >= p - i) { rem -= p - i; ++i; }
        const int l = i + rem;
        stats[p + i * p + l] = v;
        stats[p + l * p + i] = v;
      }
    }
  }
}

// Parameter covariance (proxies...py:79, computed and dropped by the reference): per-CTA partials of
//   cov[i][l] = sum_j sum_t B_{j,i}(t) (A B_{j,l})(t)   over COV_ROWS ODE rows per CTA, all p x p pairs
constexpr int COV_ROWS = 8;
__global__ void __launch_bounds__(256) cov_partial_kernel(const double* __restrict__ Bm, const double* __restrict__ AB,
                                                          int ld, int T, int n_rows, int p,
                                                          double* __restrict__ partial) {
  __shared__ double red[40];
  const int j0 = blockIdx.x * COV_ROWS, j1 = min(n_rows, j0 + COV_ROWS);
  for (int q = 0; q < p * p; ++q) {
    const int i = q / p, l = q % p;
    double v = 0.0;
    for (int j = j0; j < j1; ++j) {
      const double* bi = Bm + ((size_t)j * p + i) * ld;
      const double* al = AB + ((size_t)j * p + l) * ld;
      for (int t = threadIdx.x; t < T; t += blockDim.x) v += bi[t] * al[t];
    }
    v = block_sum(v, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * p * p + q] = v;
  }
}

__global__ void __launch_bounds__(256) cov_reduce_kernel(const double* __restrict__ partial, int n_part, int nq,
                                                         double* __restrict__ cov) {
  __shared__ double red[40];
  for (int q = 0; q < nq; ++q) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) v += partial[(size_t)i * nq + q];
    v = block_sum(v, red);
    if (threadIdx.x == 0) cov[q] = v;
  }
}

// P3: theta = solve(S, m) for constraints=None (proxies...py:93): Gaussian elimination with
// partial pivoting (what LAPACK gesv does) in one thread; p is a handful.
__global__ void param_solve_kernel(const double* __restrict__ stats, int p, double* __restrict__ theta) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  constexpr int PMAX = 32;
  double A[PMAX][PMAX + 1];
  for (int i = 0; i < p; ++i) {
    for (int j = 0; j < p; ++j) A[i][j] = stats[p + i * p + j];
    A[i][p] = stats[i];
  }
  for (int c = 0; c < p; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < p; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (piv != c)
      for (int j = c; j <= p; ++j) { double tmp = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = tmp; }
    for (int r = c + 1; r < p; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int j = c; j <= p; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = p - 1; i >= 0; --i) {
    double v = A[i][p];
    for (int j = i + 1; j < p; ++j) v -= A[i][j] * theta[j];
    theta[i] = v / A[i][i];
  }
}

}  // namespace vgm

using namespace vgm;

extern "C" int vgm_program_builtin(const char* name, vgm_program** out) {
  VGM_REQUIRE(name && out, "null argument");
  if (strcmp(name, "lorenz96") != 0) {
    set_error("unknown built-in program '%s'", name);
    return VGM_EINVAL;
  }
  vgm_program* p = new vgm_program();
  p->n_param = 1;
  p->arity = 4;
  p->builtin = true;
  p->lib = nullptr;
  p->k_stats = nullptr;
  p->k_assemble = nullptr;
  p->k_pair_R = nullptr;
  p->k_param_B = nullptr;
  p->k_param_loss = nullptr;
  *out = p;
  return VGM_OK;
}

extern "C" int vgm_program_load(const void* cubin_host, size_t size, int n_param, int arity, vgm_program** out) {
  VGM_REQUIRE(cubin_host && size > 0 && out, "null argument");
  VGM_REQUIRE(n_param >= 1 && n_param <= 32 && arity >= 1, "n_param/arity");
  VGM_CUDA_OK(cudaFree(0));
  vgm_program* p = new vgm_program();
  p->n_param = n_param;
  p->arity = arity;
  p->builtin = false;
  cudaError_t e = cudaLibraryLoadData(&p->lib, cubin_host, nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) {
    set_error("cudaLibraryLoadData: %s", cudaGetErrorString(e));
    delete p;
    return VGM_EDRIVER;
  }
  e = cudaLibraryGetKernel(&p->k_stats, p->lib, "vgm_prog_param_stats");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&p->k_assemble, p->lib, "vgm_prog_state_assemble");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&p->k_pair_R, p->lib, "vgm_prog_state_pair_R");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&p->k_param_B, p->lib, "vgm_prog_param_B");
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&p->k_param_loss, p->lib, "vgm_prog_param_loss");
  if (e != cudaSuccess) {
    set_error("cudaLibraryGetKernel: %s", cudaGetErrorString(e));
    cudaLibraryUnload(p->lib);
    delete p;
    return VGM_EDRIVER;
  }
  *out = p;
  return VGM_OK;
}

extern "C" int vgm_program_free(vgm_program* prog) {
  if (!prog) return VGM_OK;
  if (!prog->builtin && prog->lib) cudaLibraryUnload(prog->lib);
  delete prog;
  return VGM_OK;
}

extern "C" int vgm_program_n_param(const vgm_program* prog) { return prog ? prog->n_param : -1; }

// 'minimizer' parameter update: out = [loss | gradient (p)] on the device (deterministic two-pass reduction)
extern "C" size_t vgm_param_loss_ws_bytes(int n_odes, int n_param) {
  return align_up((size_t)ceil_div(n_odes, 8) * (n_param + 1) * sizeof(double), 256);
}

extern "C" int vgm_param_loss(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int n_odes,
                              const int* ode_class, const int* ode_idx, const int* ode_state, const double* theta,
                              double* out, void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(prog && X && DX && ode_class && ode_idx && ode_state && theta && out && ws, "null argument");
  VGM_REQUIRE(n_odes > 0 && T > 0 && ld >= T, "shape");
  VGM_REQUIRE(ws_bytes >= vgm_param_loss_ws_bytes(n_odes, prog->n_param), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  double* partial = (double*)ws;
  VGM_TRY(program_param_loss(prog, X, DX, ld, T, n_odes, ode_class, ode_idx, ode_state, theta, partial, s));
  cov_reduce_kernel<<<1, 256, 0, s>>>(partial, ceil_div(n_odes, 8), prog->n_param + 1, out);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

extern "C" size_t vgm_param_stats_ws_bytes(int n_odes, int T, int n_param) {
  (void)T;
  const size_t nstat = n_param + (size_t)n_param * (n_param + 1) / 2;
  return align_up((size_t)ceil_div(n_odes, 8) * nstat * sizeof(double), 256);
}

namespace vgm {
// ODE rows [j0, j1) only (j0 a multiple of 8): partials land where the full launch would put them, so a later
// param_stats_reduce over all of them gives bit-identical statistics.
int param_stats_range(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int j0, int j1,
                      const int* ode_class, const int* ode_idx, const int* ode_state, double* partial, cudaStream_t s) {
  if (j1 <= j0) return VGM_OK;
  const int nstat = prog->n_param + prog->n_param * (prog->n_param + 1) / 2;
  return program_param_stats(prog, X, DX, ld, T, j1 - j0, ode_class + j0, ode_idx + (size_t)j0 * prog->arity,
                             ode_state + j0, partial + (size_t)(j0 / 8) * nstat, s);
}
int param_stats_reduce(const vgm_program* prog, const double* partial, int n_odes, double* stats, cudaStream_t s) {
  stats_reduce_kernel<<<1, 256, 0, s>>>(partial, ceil_div(n_odes, 8), prog->n_param, stats);
  VGM_LAUNCH_OK();
  return VGM_OK;
}
}  // namespace vgm

extern "C" int vgm_param_stats(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int n_odes,
                               const int* ode_class, const int* ode_idx, const int* ode_state, double* stats,
                               void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(prog && X && DX && stats && ode_class && ode_idx && ode_state, "null argument");
  VGM_REQUIRE(n_odes > 0 && T > 0 && ld >= T, "shape");
  VGM_REQUIRE(ws && ws_bytes >= vgm_param_stats_ws_bytes(n_odes, T, prog->n_param), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  double* partial = (double*)ws;
  VGM_TRY(program_param_stats(prog, X, DX, ld, T, n_odes, ode_class, ode_idx, ode_state, partial, s));
  stats_reduce_kernel<<<1, 256, 0, s>>>(partial, ceil_div(n_odes, 8), prog->n_param, stats);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// ODE rows handled per pass of vgm_param_cov: bounds the two [rows * p][ld] scratch matrices
static int cov_chunk_rows(int n_odes, int n_param) {
  const int cap = max(COV_ROWS, (8192 / n_param) / COV_ROWS * COV_ROWS);
  return min(cap, ceil_div(n_odes, COV_ROWS) * COV_ROWS);
}

extern "C" size_t vgm_param_cov_ws_bytes(int n_odes, int ld, int n_param) {
  const size_t partial = align_up((size_t)ceil_div(n_odes, COV_ROWS) * n_param * n_param * sizeof(double), 256);
  const size_t mat = align_up((size_t)cov_chunk_rows(n_odes, n_param) * n_param * ld * sizeof(double), 256);
  return partial + 2 * mat;
}

extern "C" int vgm_param_cov(const vgm_program* prog, const double* X, const double* A, int ld, int T, int n_odes,
                             const int* ode_class, const int* ode_idx, double* cov, void* ws, size_t ws_bytes,
                             vgm_stream_t stream) {
  VGM_REQUIRE(prog && X && A && cov && ode_class && ode_idx, "null argument");
  VGM_REQUIRE(n_odes > 0 && T > 0 && ld >= T, "shape");
  const int p = prog->n_param;
  VGM_REQUIRE(ws && ws_bytes >= vgm_param_cov_ws_bytes(n_odes, ld, p), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int ch = cov_chunk_rows(n_odes, p);
  const size_t partial_bytes = align_up((size_t)ceil_div(n_odes, COV_ROWS) * p * p * sizeof(double), 256);
  const size_t mat_bytes = align_up((size_t)ch * p * ld * sizeof(double), 256);
  double* partial = (double*)ws;
  double* Bm = (double*)((char*)ws + partial_bytes);
  double* AB = (double*)((char*)ws + partial_bytes + mat_bytes);
  for (int j0 = 0; j0 < n_odes; j0 += ch) {
    const int n = min(ch, n_odes - j0);
    VGM_TRY(program_param_B(prog, X, ld, T, n, ode_class + j0, ode_idx + (size_t)j0 * prog->arity, Bm, s));
    vgm_gemm_args g;                              // AB[r][t] = sum_s Bm[r][s] A[t][s] = (A B_r)(t)
    memset(&g, 0, sizeof(g));
    g.A = Bm; g.B = A; g.C = AB;
    g.lda = g.ldb = g.ldc = ld;
    g.N = n * p; g.M = T; g.K = T;
    g.alpha = 1.0;
    VGM_TRY(dgemm_nt(g, s));
    cov_partial_kernel<<<ceil_div(n, COV_ROWS), 256, 0, s>>>(Bm, AB, ld, T, n, p,
                                                           partial + (size_t)(j0 / COV_ROWS) * p * p);
    VGM_LAUNCH_OK();
  }
  cov_reduce_kernel<<<1, 256, 0, s>>>(partial, ceil_div(n_odes, COV_ROWS), p * p, cov);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

extern "C" int vgm_param_solve(const double* stats, int n_param, double* theta, vgm_stream_t stream) {
  VGM_REQUIRE(stats && theta && n_param >= 1 && n_param <= 32, "param_solve arguments");
  param_solve_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(stats, n_param, theta);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

extern "C" int vgm_state_assemble(const vgm_program* prog, const vgm_sweep_plan* plan, const double* X,
                                  const double* DX0, const double* theta_star, double* Lambda, double* rhs, double* ruu,
                                  double* Ruu, double* Rlive, vgm_stream_t stream) {
  VGM_REQUIRE(prog && plan && X && DX0 && theta_star && Lambda && rhs && ruu, "null argument");
  VGM_REQUIRE(plan->n_live_pairs == 0 || Rlive, "Rlive buffer required when the plan has live pairs");
  return program_state_assemble(prog, plan, X, DX0, theta_star, Lambda, rhs, ruu, Ruu, Rlive, (cudaStream_t)stream);
}
