// P4: the per-state Gaussian update of proxy_for_ind_states(optimizer='analytical'),
// proxies_for_ode_parameters_and_states.py:185-262, for ALL hidden states at once.
//
// For hidden state u the reference forms a dense T x T system
//     (B_uu^T B_uu + diag(Lambda_u)) x_u = rhs_u,   B_uu = diag(R_uu) - D        (:229-232, :245, :262)
// and calls LAPACK gesv on it.  Here every state of a dependency level is solved together:
//   * shared system matrix (R_uu constant, e.g. -1 for Lorenz96: M = (cI - D)^T (cI - D)): mixed-precision
//     iterative refinement, ir_solve -- FP64 residuals r = rhs - x M^T - Lambda o x as ONE banded DMMA GEMM per
//     pass, corrections by FP32 preconditioned CG whose operator products are bf16-split tcgen05 GEMMs
//     (tc_gemm.cu) and whose vector work lives in refine.cu; systems it cannot contract (too ill-conditioned)
//     are finished by
//   * FP64 preconditioned CG on the DMMA pipe, pcg_solve -- operator Q = P M^T + Lambda o P (one GEMM), or two
//     GEMMs with D and D^T when R_uu varies per state; optional preconditioner G = (M + shift I)^-1.
// Residuals are driven to ||r|| <= tol ||rhs|| per state (default 1e-13), i.e. far below the 1e-9 parity
// tolerance, and the sequential Gauss-Seidel semantics of the reference (:185 alias, :225 live read) are
// reproduced by solving dependency levels in order and patching rhs with the freshly updated D x_k between
// levels.
#include <cuda_bf16.h>
#include <stdlib.h>

#include <utility>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"
#include "tc_gemm.cuh"
#include "refine.cuh"
#include "small_solve.cuh"
#include "dense_solve.cuh"

struct vgm_program;

namespace vgm {

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);
int gemm_dot_tiles(int M, int N);
constexpr int GEMM_MIN_BN = 32;   // narrowest column tile of any DMMA GEMM configuration (dgemm.cu)


int program_state_assemble(const vgm_program* prog, const vgm_sweep_plan* pl, const double* X, const double* DX0,
                           const double* theta, double* Lambda, double* rhs, double* ruu, double* Ruu, double* Rlive,
                           cudaStream_t s);

int program_state_pair_R(const vgm_program* prog, const vgm_sweep_plan* pl, const double* X, const double* theta,
                         int slot, double* Rout, cudaStream_t s);
int transpose(const double* in, int ld_in, double* out, int ld_out, int rows, int cols, cudaStream_t s);

static inline int row_threads(int T) { return max(64, min(256, (T + 31) / 32 * 32)); }

// ------------------------------------------------------------------------------------
// CG vector kernels: one CTA per active row.  act[] lists hidden slots; *n_act rows are live.
// ------------------------------------------------------------------------------------
struct CgBuf {
  double *R, *P, *Q, *Z, *W;      // [n_hidden][ld]
  double *pq_part, *rz_part;      // [n_hidden][n_tiles]
  double *rz, *rr, *bb;           // [n_hidden]
  int *done, *iters;              // [n_hidden]
  int *act[2];                    // ping-pong active lists
  int *n_act;                     // device count (2 ints: [0] current, [1] scratch)
  int n_tiles;                    // partial dots per row written by the operator GEMM
  int n_tiles_z;                  // ... by the preconditioner GEMM
};

__global__ void cg_gather_x_kernel(const double* __restrict__ X, const int* __restrict__ slot_state,
                                   const int* __restrict__ act, const int* __restrict__ n_act, int T, int ld,
                                   double* __restrict__ P) {
  if ((int)blockIdx.x >= *n_act) return;
  const int slot = act[blockIdx.x];
  const double* x = X + (size_t)slot_state[slot] * ld;
  double* p = P + (size_t)slot * ld;
  for (int t = threadIdx.x; t < T; t += blockDim.x) p[t] = x[t];
}

// r = b - q (q = Op x0 already includes the Lambda term); rr, bb; rows without a self pair are
// solved directly: x = rhs / (2 Lambda)  (alias quirk, proxies...py:201,245).
__global__ void cg_init_kernel(CgBuf cb, const double* __restrict__ rhs, const double* __restrict__ Lambda,
                               double* __restrict__ X, const int* __restrict__ slot_state,
                               const int* __restrict__ slot_has_self, const int* __restrict__ act, int T, int ld,
                               double tol2, int precond) {
  __shared__ double red[40];
  if ((int)blockIdx.x >= *cb.n_act) return;
  const int slot = act[blockIdx.x];
  const size_t o = (size_t)slot * ld;
  if (slot_has_self && !slot_has_self[slot]) {
    double* x = X + (size_t)slot_state[slot] * ld;
    for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] = rhs[o + t] / (2.0 * Lambda[o + t]);
    if (threadIdx.x == 0) { cb.done[slot] = 1; cb.iters[slot] = 0; }
    return;
  }
  double rr = 0.0, bb = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double b = rhs[o + t];
    const double r = b - cb.Q[o + t];
    cb.R[o + t] = r;
    if (!precond) cb.P[o + t] = r;
    rr += r * r;
    bb += b * b;
  }
  rr = block_sum(rr, red);
  bb = block_sum(bb, red);
  if (bb == 0.0) {  // zero right-hand side: the solution is zero
    double* x = X + (size_t)slot_state[slot] * ld;
    for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] = 0.0;
  }
  if (threadIdx.x == 0) {
    cb.rr[slot] = rr;
    cb.bb[slot] = bb;
    cb.rz[slot] = rr;
    cb.iters[slot] = 0;
    cb.done[slot] = (bb == 0.0 || rr <= tol2 * bb) ? 1 : 0;
  }
}

// p = z, rz = r.z (from the GEMM's partial dots)
__global__ void cg_init_precond_kernel(CgBuf cb, const int* __restrict__ act, int T, int ld) {
  if ((int)blockIdx.x >= *cb.n_act) return;
  const int slot = act[blockIdx.x];
  if (cb.done[slot]) return;
  const size_t o = (size_t)slot * ld;
  for (int t = threadIdx.x; t < T; t += blockDim.x) cb.P[o + t] = cb.Z[o + t];
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < cb.n_tiles_z; ++j) s += cb.rz_part[(size_t)slot * cb.n_tiles_z + j];
    cb.rz[slot] = s;
  }
}

// alpha = rz / (p.q); x += alpha p; r -= alpha q; rr = r.r; convergence test.
// Plain CG (precond == 0) also finishes the iteration: beta = rr_new / rr_old, p = r + beta p.
__global__ void cg_update_kernel(CgBuf cb, double* __restrict__ X, const int* __restrict__ slot_state,
                                 const int* __restrict__ act, int T, int ld, double tol2, int precond) {
  __shared__ double red[40];
  if ((int)blockIdx.x >= *cb.n_act) return;
  const int slot = act[blockIdx.x];
  if (cb.done[slot]) return;
  const size_t o = (size_t)slot * ld;
  double pq = 0.0;
  for (int j = 0; j < cb.n_tiles; ++j) pq += cb.pq_part[(size_t)slot * cb.n_tiles + j];
  const double rz = cb.rz[slot];
  const double alpha = rz / pq;
  double* x = X + (size_t)slot_state[slot] * ld;
  double rr = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double p = cb.P[o + t];
    x[t] += alpha * p;
    const double r = cb.R[o + t] - alpha * cb.Q[o + t];
    cb.R[o + t] = r;
    rr += r * r;
  }
  rr = block_sum(rr, red);
  const bool conv = (rr <= tol2 * cb.bb[slot]);
  if (!precond && !conv) {
    const double beta = rr / rz;
    for (int t = threadIdx.x; t < T; t += blockDim.x) cb.P[o + t] = cb.R[o + t] + beta * cb.P[o + t];
  }
  if (threadIdx.x == 0) {
    cb.rr[slot] = rr;
    cb.iters[slot] += 1;
    if (!precond) cb.rz[slot] = rr;
    if (conv) cb.done[slot] = 1;
  }
}

// beta = (r.z)_new / (r.z)_old ; p = z + beta p
__global__ void cg_precond_dir_kernel(CgBuf cb, const int* __restrict__ act, int T, int ld) {
  if ((int)blockIdx.x >= *cb.n_act) return;
  const int slot = act[blockIdx.x];
  if (cb.done[slot]) return;
  const size_t o = (size_t)slot * ld;
  double rz_new = 0.0;
  for (int j = 0; j < cb.n_tiles_z; ++j) rz_new += cb.rz_part[(size_t)slot * cb.n_tiles_z + j];
  const double beta = rz_new / cb.rz[slot];
  for (int t = threadIdx.x; t < T; t += blockDim.x) cb.P[o + t] = cb.Z[o + t] + beta * cb.P[o + t];
  __syncthreads();
  if (threadIdx.x == 0) cb.rz[slot] = rz_new;
}

// Stable compaction of the active list (single CTA): keeps rows with done == 0.  Each thread takes 4 consecutive
// entries, offsets come from a warp-shuffle scan + a scan of the 32 warp totals by warp 0 (two barriers per 4096
// entries; the first version -- one entry per thread, every thread adding up all warp totals -- took 63 us for the
// 50 000 rows of config C5, 7 times per sweep).
__global__ void __launch_bounds__(1024) cg_compact_kernel(const int* __restrict__ act_in, int* __restrict__ act_out,
                                                          int* __restrict__ n_act, const int* __restrict__ done) {
  __shared__ int warp_off[32];
  __shared__ int chunk_tot;
  pdl_wait();
  const int n = n_act[0];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int base = 0;
  for (int start = 0; start < n; start += 4096) {
    const int i0 = start + 4 * tid;
    int slot[4], keep[4], mine = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + e;
      slot[e] = (i < n) ? act_in[i] : -1;
      keep[e] = (i < n && !done[slot[e]]) ? 1 : 0;
      mine += keep[e];
    }
    int incl = mine;                                   // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_off[w] = incl;
    __syncthreads();
    if (w == 0) {
      const int tot = warp_off[lane];
      int sc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += v;
      }
      warp_off[lane] = sc - tot;                       // exclusive offsets of the warps
      if (lane == 31) chunk_tot = sc;
    }
    __syncthreads();
    int pos = base + warp_off[w] + incl - mine;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (keep[e]) act_out[pos++] = slot[e];
    base += chunk_tot;
    __syncthreads();                                   // warp_off / chunk_tot are rewritten by the next chunk
  }
  if (tid == 0) n_act[1] = base;
}

// n_act = { current count, scratch (new count), "active set changed since the previous correction", "fresh list" }
__global__ void cg_commit_count_kernel(int* n_act) {
  pdl_wait();
  n_act[2] = (n_act[1] != n_act[0]) || n_act[3];
  n_act[3] = 0;
  n_act[0] = n_act[1];
}

__global__ void cg_set_count_kernel(int* n_act, int n) { n_act[0] = n; n_act[1] = n; n_act[2] = 1; n_act[3] = 1; }

__global__ void cg_count_unconverged_kernel(const int* __restrict__ rows, int n_rows, const int* __restrict__ done,
                                            const int* __restrict__ iters, int* __restrict__ out) {
  // out[0] += #rows not done, out[1] = max iterations (out is zeroed by the caller)
  int bad = 0, mx = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
    const int s = rows[i];
    bad += done[s] ? 0 : 1;
    mx = max(mx, iters[s]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(out, bad);
    atomicMax(out + 1, mx);
  }
}

// rhs[slot] += Rlive[q] o DXlive[src]  for the live pairs of one dependency level
__global__ void add_live_terms_kernel(const int* __restrict__ live_pair_slot, const int* __restrict__ live_pair_src,
                                      int q0, const double* __restrict__ Rlive, const double* __restrict__ DXlive,
                                      int T, int ld, double* __restrict__ rhs) {
  const int q = q0 + blockIdx.x;
  const int slot = live_pair_slot[q];
  const double* R = Rlive + (size_t)q * ld;
  const double* dx = DXlive + (size_t)live_pair_src[q] * ld;
  double* b = rhs + (size_t)slot * ld;
  for (int t = threadIdx.x; t < T; t += blockDim.x) atomicAdd(b + t, R[t] * dx[t]);
}

// ------------------------------------------------------------------------------------
struct PcgWs {
  CgBuf cb;
  IrBuf ir;                      // low-precision side of the mixed-precision solver
  float* small_scratch;          // factor scratch of the tiny-level solver (+ 64 status ints at its end)
  int* host_poll_dev;            // 2 ints
};

static size_t pcg_ws_bytes(int n_hidden, int T, int ld) {
  const size_t mat = align_up((size_t)n_hidden * ld * sizeof(double), 256);
  const int n_tiles = ceil_div(T, GEMM_MIN_BN);  // upper bound over GEMM configurations
  const size_t part = align_up((size_t)n_hidden * n_tiles * sizeof(double), 256);
  const size_t vec = align_up((size_t)n_hidden * sizeof(double), 256);
  return 5 * mat + ir_ws_bytes(n_hidden, ld) + small_solve_scratch_bytes(T) + 2 * part + 3 * vec +
         4 * vec /*done, iters, act x2 (ints fit)*/ + 1024;
}

static int pcg_carve(void* ws, size_t ws_bytes, int n_hidden, int T, int ld, PcgWs& w) {
  VGM_REQUIRE(ws && ws_bytes >= pcg_ws_bytes(n_hidden, T, ld), "PCG workspace too small");
  VGM_REQUIRE(((uintptr_t)ws % 256) == 0, "workspace must be 256-byte aligned");
  char* p = (char*)ws;
  const size_t mat = align_up((size_t)n_hidden * ld * sizeof(double), 256);
  const int n_tiles = ceil_div(T, GEMM_MIN_BN);
  const size_t part = align_up((size_t)n_hidden * n_tiles * sizeof(double), 256);
  const size_t vec = align_up((size_t)n_hidden * sizeof(double), 256);
  CgBuf& c = w.cb;
  c.R = (double*)p; p += mat;
  c.P = (double*)p; p += mat;
  c.Q = (double*)p; p += mat;
  c.Z = (double*)p; p += mat;
  c.W = (double*)p; p += mat;
  ir_carve(p, n_hidden, ld, w.ir);
  w.small_scratch = (float*)p; p += small_solve_scratch_bytes(T);
  c.pq_part = (double*)p; p += part;
  c.rz_part = (double*)p; p += part;
  c.rz = (double*)p; p += vec;
  c.rr = (double*)p; p += vec;
  c.bb = (double*)p; p += vec;
  c.done = (int*)p; p += vec;
  c.iters = (int*)p; p += vec;
  c.act[0] = (int*)p; p += vec;
  c.act[1] = (int*)p; p += vec;
  c.n_act = (int*)p; p += 256;
  w.host_poll_dev = (int*)p; p += 256;
  c.n_tiles = 0;
  c.n_tiles_z = 0;
  return VGM_OK;
}

struct Counters {
  int gemm = 0, launches = 0;
};
static const bool g_debug = getenv("VGM_DEBUG") != nullptr;

// Q = Op P (+ Lambda o P), partial dots p.q -> pq_part
static int apply_operator(const vgm_operators* ops, int ruu_const, const double* Ruu, const double* Lambda, CgBuf& cb,
                          const int* act, int n_upper, int T, int ld, bool want_dot, Counters& cnt, cudaStream_t s) {
  vgm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.lda = g.ldb = g.ldc = ld;
  g.N = n_upper; g.M = T; g.K = T;
  g.rowsA = act; g.rowsC = act; g.n_rows_dev = cb.n_act;
  if (ruu_const) {
    g.A = cb.P; g.B = ops->M; g.C = cb.Q;
    g.alpha = 1.0;
    g.band = ops->band_M;
    g.U1 = Lambda; g.V1 = cb.P; g.g1 = 1.0;
    if (want_dot) { g.dotw = cb.P; g.dot_out = cb.pq_part; g.dot_ld = cb.n_tiles; }
    VGM_TRY(dgemm_nt(g, s));
    cnt.gemm += 1; cnt.launches += 1;
  } else {
    // W = Ruu o P - D P
    g.A = cb.P; g.B = ops->D; g.C = cb.W;
    g.alpha = -1.0;
    g.band = ops->band_D;
    g.U1 = Ruu; g.V1 = cb.P; g.g1 = 1.0;
    VGM_TRY(dgemm_nt(g, s));
    // Q = Ruu o W - D^T W + Lambda o P
    g.A = cb.W; g.B = ops->DT; g.C = cb.Q;
    g.U1 = Ruu; g.V1 = cb.W; g.g1 = 1.0;
    g.U2 = Lambda; g.V2 = cb.P; g.g2 = 1.0;
    if (want_dot) { g.dotw = cb.P; g.dot_out = cb.pq_part; g.dot_ld = cb.n_tiles; }
    VGM_TRY(dgemm_nt(g, s));
    cnt.gemm += 2; cnt.launches += 2;
  }
  return VGM_OK;
}

// FP64 preconditioner: Z = R G^T on the DMMA pipe, partial dots r.z -> rz_part.
static int apply_preconditioner(const vgm_operators* ops, CgBuf& cb, const int* act, int n_upper, int T, int ld,
                                Counters& cnt, cudaStream_t s) {
  vgm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = cb.R; g.B = ops->G; g.C = cb.Z;
  g.lda = g.ldb = g.ldc = ld;
  g.N = n_upper; g.M = T; g.K = T;
  g.rowsA = act; g.rowsC = act; g.n_rows_dev = cb.n_act;
  g.alpha = 1.0;
  g.dotw = cb.R; g.dot_out = cb.rz_part; g.dot_ld = cb.n_tiles_z;
  VGM_TRY(dgemm_nt(g, s));
  cnt.gemm += 1; cnt.launches += 1;
  return VGM_OK;
}

// Drop converged rows from the active list and fetch the new count (one 4-byte D2H poll).
static int compact_and_poll(CgBuf& cb, int& cur, const int*& act, int& n_upper, Counters& cnt, cudaStream_t s) {
  cg_compact_kernel<<<1, 1024, 0, s>>>(cb.act[cur], cb.act[cur ^ 1], cb.n_act, cb.done);
  cg_commit_count_kernel<<<1, 1, 0, s>>>(cb.n_act);
  VGM_LAUNCH_OK();
  cnt.launches += 2;
  cur ^= 1;
  act = cb.act[cur];
  int host_cnt = 0;
  VGM_CUDA_OK(cudaMemcpyAsync(&host_cnt, cb.n_act, sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  n_upper = host_cnt;
  return VGM_OK;
}

// `final` = false: a non-converged result is only reported through the return code (the caller has a fallback)
static int report_convergence(CgBuf& cb, PcgWs& w, const int* rows, int n_rows, double tol, int cap,
                              vgm_sweep_info* info, Counters& cnt, cudaStream_t s, bool final = true) {
  VGM_CUDA_OK(cudaMemsetAsync(w.host_poll_dev, 0, 2 * sizeof(int), s));
  cg_count_unconverged_kernel<<<min(148, ceil_div(n_rows, 256)), 256, 0, s>>>(rows, n_rows, cb.done, cb.iters,
                                                                              w.host_poll_dev);
  VGM_LAUNCH_OK();
  int res[2] = {0, 0};
  VGM_CUDA_OK(cudaMemcpyAsync(res, w.host_poll_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  cnt.launches += 1;
  if (info) {
    info->iterations = max(info->iterations, res[1]);
    if (final) info->unconverged += res[0];
  }
  if (res[0] > 0) {
    if (final) set_error("solver: %d of %d systems did not reach tol=%g within %d iterations", res[0], n_rows, tol, cap);
    return VGM_ENOTCONV;
  }
  return VGM_OK;
}

// ---- mixed-precision iterative refinement ----------------------------------------------------------------
// bb = ||rhs||^2; rows without a self pair are solved directly, x = rhs / (2 Lambda) (alias quirk,
// proxies...py:201,245); a zero right-hand side gives x = 0.
__global__ void ir_prep_kernel(CgBuf cb, const double* __restrict__ rhs, const double* __restrict__ Lambda,
                               double* __restrict__ X, const int* __restrict__ slot_state,
                               const int* __restrict__ slot_has_self, const int* __restrict__ act, int T, int ld) {
  __shared__ double red[40];
  const int slot = act[blockIdx.x];
  const size_t o = (size_t)slot * ld;
  double* x = X + (size_t)slot_state[slot] * ld;
  if (slot_has_self && !slot_has_self[slot]) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] = rhs[o + t] / (2.0 * Lambda[o + t]);
    if (threadIdx.x == 0) { cb.done[slot] = 1; cb.iters[slot] = 0; }
    return;
  }
  double bb = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) bb += rhs[o + t] * rhs[o + t];
  bb = block_sum(bb, red);
  if (bb == 0.0)
    for (int t = threadIdx.x; t < T; t += blockDim.x) x[t] = 0.0;
  if (threadIdx.x == 0) {
    cb.bb[slot] = bb;
    cb.iters[slot] = 0;
    cb.done[slot] = (bb == 0.0) ? 1 : 0;
  }
}

// rr = ||r||^2 from the residual GEMM's partial dots; convergence test of the outer iteration.
__global__ void ir_check_kernel(CgBuf cb, const int* __restrict__ act, int n_tiles, int dot_ld, double tol2, int outer,
                                int inner) {
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *cb.n_act) return;
  const int slot = act[j];
  if (cb.done[slot]) return;
  double rr = 0.0;
  for (int i = 0; i < n_tiles; ++i) rr += cb.pq_part[(size_t)slot * dot_ld + i];
  cb.rr[slot] = rr;
  cb.iters[slot] = outer * inner;
  if (rr <= tol2 * cb.bb[slot]) cb.done[slot] = 1;
}

// Solve (M + diag Lambda_u) x_u = rhs_u for the rows of one dependency level:
//   outer (FP64, DMMA):  r = rhs - (x M^T + Lambda o x), test ||r|| <= tol ||rhs||, x += sigma d
//   inner (FP32 vectors): `inner` steps of PCG on A d = r / sigma with the operator applied as a bf16-split
//                         tcgen05 GEMM (3 products, ~2^-16) and the preconditioner S G S (G = (M + shift I)^-1
//                         in bf16, S the diagonal scaling of vgm_operators.precond_c).
// Each correction contracts the error by ~1e-3..1e-4, so 4-5 FP64 GEMMs replace the ~60 of FP64 PCG.
// A level can be split into independent row chunks driven on separate streams (VGM_IR_CHUNKS, default 1): that
// overlapped the L2-bound tensor-core GEMMs of one chunk with the HBM-bound vector kernels of another while
// the GEMM epilogue was inefficient; with the TMA-store epilogue every kernel is HBM-bound and one chunk wins.
struct IrChunk {
  int row0 = 0, n_rows = 0;        // slice of the level's row list
  int cur = 0, n_upper = 0;
  int* act_buf[2] = {nullptr, nullptr};
  const int* act = nullptr;
  int* n_act = nullptr;            // device: [0] current count, [1] scratch
  IrBuf ir;
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
  volatile int* host_cnt = nullptr;  // pinned
  bool finished = false;
};

constexpr int IR_MAX_CHUNKS = 4;
struct IrAsync {                   // created once, kept for the life of the process
  cudaStream_t side[IR_MAX_CHUNKS] = {};       // side[0] unused: chunk 0 runs on the caller's stream
  cudaEvent_t ev[IR_MAX_CHUNKS] = {}, join[IR_MAX_CHUNKS] = {}, fork = nullptr;
  int* host_cnt = nullptr;         // pinned, IR_MAX_CHUNKS x 16 ints
};
static IrAsync g_ir_async;

static int ir_async_init() {
  IrAsync& a = g_ir_async;
  if (a.host_cnt) return VGM_OK;
  for (int i = 0; i < IR_MAX_CHUNKS; ++i) {
    if (i > 0) VGM_CUDA_OK(cudaStreamCreateWithFlags(&a.side[i], cudaStreamNonBlocking));
    VGM_CUDA_OK(cudaEventCreateWithFlags(&a.ev[i], cudaEventDisableTiming));
    VGM_CUDA_OK(cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming));
  }
  VGM_CUDA_OK(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
  VGM_CUDA_OK(cudaHostAlloc((void**)&a.host_cnt, IR_MAX_CHUNKS * 16 * sizeof(int), cudaHostAllocDefault));
  return VGM_OK;
}

static IrBuf ir_view(const IrBuf& b, int row0, int ld) {
  IrBuf v = b;
  const size_t o = (size_t)row0 * ld;
  v.r32 += o; v.d32 += o; v.acc32 += o; v.lam32 += o;
  v.p_hi += o; v.p_lo += o; v.q_hi += o; v.q_lo += o; v.rsb += o; v.s16 += o;
  v.sigma += row0; v.rz += row0; v.rz_part += (size_t)row0 * b.rz_tiles;
  return v;
}

constexpr int IR_MAX_OUTER = 12;
// A correction over a small active set (straggler passes, the wrap-around level, small per-GPU blocks) is ~45
// launch-bound kernels; it is captured once into a CUDA graph (kernels read the exact row count on the device, so
// the same graph serves every pass over at most IR_GRAPH_MAX_ROWS rows) and replayed.
constexpr int IR_GRAPH_MAX_ROWS = 1024;
struct IrGraphKey {
  const void *r32, *act, *n_act, *G, *M, *X, *Lambda, *R, *P, *rr, *slot_state;
  int n, T, ld, inner, band_G, band_M;
  double shift, c;
  bool operator==(const IrGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; }
};
struct IrGraphEntry {
  IrGraphKey key;
  cudaGraphExec_t exec;
};
constexpr int IR_GRAPH_CACHE = 16;
static IrGraphEntry g_ir_graphs[IR_GRAPH_CACHE];
static int g_ir_graph_count = 0, g_ir_graph_next = 0;
static const bool g_no_graph = getenv("VGM_NO_GRAPH") != nullptr;
constexpr int IR_CHUNK_MIN_ROWS = 4096;   // a chunk is at least this many rows; smaller levels run as one chunk
void ir_graph_cache_clear() {
  for (int i = 0; i < g_ir_graph_count; ++i)
    if (g_ir_graphs[i].exec) { cudaGraphExecDestroy(g_ir_graphs[i].exec); g_ir_graphs[i].exec = nullptr; }
  g_ir_graph_count = g_ir_graph_next = 0;
}
static int ir_chunk_target() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("VGM_IR_CHUNKS");
    n = e ? atoi(e) : 1;
    n = max(1, min(IR_MAX_CHUNKS, n));
  }
  return n;
}

static int ir_solve(const vgm_operators* ops, const double* Lambda, const double* rhs, double* X,
                    const int* slot_state, const int* slot_has_self, const int* rows, int n_rows, int T, int ld,
                    const vgm_solver_opts* opts, PcgWs& w, vgm_sweep_info* info, Counters& cnt, cudaStream_t s) {
  const double tol = (opts && opts->tol > 0) ? opts->tol : 1e-13;
  // a correction normally gains 3-4 digits; a system that needs more than IR_MAX_OUTER of them is too ill-conditioned
  // for FP32 / bf16-split corrections (cond(A) 2^-16 >~ 1) and is handed to FP64 PCG by the caller
  const int max_outer = IR_MAX_OUTER;
  const int inner = (opts && opts->inner_iters > 0) ? opts->inner_iters : 10;
  // corrections enqueued back to back before the host looks at a convergence count for the first time: the
  // kernels read the exact active count on the device, so a pass over already converged rows costs nothing
  const int optimistic = (opts && opts->optimistic_passes > 0) ? opts->optimistic_passes : (opts && opts->optimistic_passes < 0 ? 0 : 4);
  CgBuf& cb = w.cb;
  const int th = row_threads(T);
  const __nv_bfloat16* M_hi = (const __nv_bfloat16*)ops->M_bf16x2;
  const __nv_bfloat16* M_lo = M_hi + (size_t)T * ld;
  VGM_TRY(ir_async_init());
  IrAsync& as = g_ir_async;
  const int dot_ld = ceil_div(T, GEMM_MIN_BN);           // row stride of pq_part (its allocation bound)

  const int n_chunks = max(1, min(ir_chunk_target(), n_rows / IR_CHUNK_MIN_ROWS));
  IrChunk ch[IR_MAX_CHUNKS];
  // chunk boundaries on multiples of 128 rows (whole GEMM row tiles)
  const int per = ((ceil_div(n_rows, n_chunks) + 127) / 128) * 128;
  for (int c = 0; c < n_chunks; ++c) {
    IrChunk& k = ch[c];
    k.row0 = min(c * per, n_rows);
    k.n_rows = min(per, n_rows - k.row0);
    k.act_buf[0] = cb.act[0] + k.row0;
    k.act_buf[1] = cb.act[1] + k.row0;
    k.n_act = cb.n_act + 4 * c;
    k.ir = ir_view(w.ir, k.row0, ld);
    k.st = (c == 0) ? s : as.side[c];
    k.ev = as.ev[c];
    k.host_cnt = as.host_cnt + 16 * c;
    k.n_upper = k.n_rows;
    k.act = k.act_buf[0];
  }
  if (n_chunks > 1) {
    VGM_CUDA_OK(cudaEventRecord(as.fork, s));
    for (int c = 1; c < n_chunks; ++c) VGM_CUDA_OK(cudaStreamWaitEvent(as.side[c], as.fork, 0));
  }
  for (int c = 0; c < n_chunks; ++c) {
    IrChunk& k = ch[c];
    if (k.n_rows <= 0) { k.finished = true; continue; }
    CgBuf cbk = cb;
    cbk.n_act = k.n_act;
    VGM_CUDA_OK(cudaMemcpyAsync(k.act_buf[0], rows + k.row0, (size_t)k.n_rows * sizeof(int), cudaMemcpyDeviceToDevice, k.st));
    cg_set_count_kernel<<<1, 1, 0, k.st>>>(k.n_act, k.n_rows);
    {
      ProfScope prof(PROF_VEC, 0.0, (double)k.n_rows * T * 24.0, k.st);
      ir_prep_kernel<<<k.n_rows, th, 0, k.st>>>(cbk, rhs, Lambda, X, slot_state, slot_has_self, k.act, T, ld);
      cg_gather_x_kernel<<<k.n_rows, th, 0, k.st>>>(X, slot_state, k.act, k.n_act, T, ld, cb.P);
    }
    VGM_LAUNCH_OK();
    cnt.launches += 4;
  }

  for (int outer = 0; outer <= max_outer; ++outer) {
    // --- FP64 residual, convergence test, compaction; the new count travels to pinned host memory ---
    for (int c = 0; c < n_chunks; ++c) {
      IrChunk& k = ch[c];
      if (k.finished) continue;
      CgBuf cbk = cb;
      cbk.n_act = k.n_act;
      const int n_tiles = gemm_dot_tiles(T, k.n_upper);
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = cb.P; g.B = ops->M; g.C = cb.R;
      g.lda = g.ldb = g.ldc = ld;
      g.N = k.n_upper; g.M = T; g.K = T;
      g.rowsA = k.act; g.rowsC = k.act; g.n_rows_dev = k.n_act;
      g.alpha = -1.0; g.beta = 1.0; g.C0 = rhs;
      g.U1 = Lambda; g.V1 = cb.P; g.g1 = -1.0;
      g.dotw = cb.R; g.dot_out = cb.pq_part; g.dot_ld = dot_ld;   // fixed stride: chunks may tile differently
      g.band = ops->band_M;
      VGM_TRY(dgemm_nt(g, k.st));
      VGM_CUDA_OK(launch_pdl(ir_check_kernel, dim3(ceil_div(k.n_upper, 256)), dim3(256), 0, k.st, cbk, k.act, n_tiles, dot_ld,
                             tol * tol, outer, inner));
      VGM_CUDA_OK(launch_pdl(cg_compact_kernel, dim3(1), dim3(1024), 0, k.st, (const int*)k.act_buf[k.cur], k.act_buf[k.cur ^ 1],
                             k.n_act, (const int*)cb.done));
      VGM_CUDA_OK(launch_pdl(cg_commit_count_kernel, dim3(1), dim3(1), 0, k.st, k.n_act));
      k.cur ^= 1;
      k.act = k.act_buf[k.cur];
      VGM_CUDA_OK(cudaMemcpyAsync((void*)k.host_cnt, k.n_act, sizeof(int), cudaMemcpyDeviceToHost, k.st));
      VGM_CUDA_OK(cudaEventRecord(k.ev, k.st));
      cnt.gemm += 1; cnt.launches += 4;
    }
    // --- corrections: enqueue each chunk's inner solve as soon as its count is known ---
    bool any = false;
    for (int c = 0; c < n_chunks; ++c) {
      IrChunk& k = ch[c];
      if (k.finished) continue;
      if (outer >= optimistic || outer == max_outer) {
        VGM_CUDA_OK(cudaEventSynchronize(k.ev));
        k.n_upper = *k.host_cnt;
        if (g_debug) fprintf(stderr, "[vgm] refinement %d chunk %d: %d of %d systems above tol\n", outer, c, k.n_upper, k.n_rows);
        if (k.n_upper == 0 || outer == max_outer) { k.finished = true; continue; }
      }
      any = true;
      // n: host upper bound of the rows (k.n_act is exact); st: the stream to enqueue on / capture from
      // experiment hook: VGM_IR_SCHEDULE="10,10,10,7" sets the CG steps of correction 0,1,2,... (last value repeats)
      static int sched[16], n_sched = -1;
      if (n_sched < 0) {
        n_sched = 0;
        if (const char* e = getenv("VGM_IR_SCHEDULE")) {
          for (const char* q = e; *q && n_sched < 16;) { sched[n_sched++] = atoi(q); while (*q && *q != ',') ++q; if (*q) ++q; }
        }
      }
      // Straggler passes (after the optimistic full-width corrections: small, launch-bound active sets) take six more
      // CG steps: at K = 100 000 the 5 % of the systems that are still above 1e-13 after four corrections then finish
      // in ONE pass instead of two (a pass costs a residual GEMM, the check, the compaction and a host poll).
      const int inner_now = n_sched > 0 ? max(1, sched[min(outer, n_sched - 1)])
                                        : (outer >= optimistic && optimistic > 0 ? inner + 6 : inner);
      auto enqueue_correction = [&](int n, cudaStream_t st) -> int {
        const int inner = inner_now;
        IrBuf ib = k.ir;                                    // ir_dir reads p and writes q: swap after every call
        auto swap_pq = [&]() { std::swap(ib.p_hi, ib.q_hi); std::swap(ib.p_lo, ib.q_lo); };
        VGM_TRY(ir_inner_init(ib, cb.R, Lambda, cb.rr, k.act, n, k.n_act, T, ld, ops->G_shift, ops->precond_c, st));
        TcGemmArgs tg;
        memset(&tg, 0, sizeof(tg));
        tg.n_rows = n; tg.n_rows_dev = k.n_act; tg.M = T; tg.K = T; tg.lda = ld; tg.ldb = ld; tg.out = ib.acc32; tg.ldo = ld;
        TcGemmArgs tp = tg;                                 // z' = (s o r) G^T, stored as bf16 when ir_dir can read that
        // (no effect on the convergence: profiles/solver_study.md); 2: the GEMM epilogue also leaves (s o r).z' per tile
        const int zb16 = ir_dir_takes_bf16(T) ? (ir_dir_takes_dots() ? 2 : 1) : 0;
        tp.A0 = ib.rsb; tp.B0 = ops->G_bf16; tp.band = ops->band_G; tp.out_bf16 = zb16 ? 1 : 0;
        if (zb16 == 2) { tp.dot_w = ib.rsb; tp.dot_ldw = ld; tp.dot_out = ib.rz_part; tp.dot_ld = ib.rz_tiles; }
        TcGemmArgs to = tg;                                 // q' = p M^T (bf16 split)
        to.B0 = M_hi; to.B1 = M_lo; to.band = ops->band_M_lp;
        VGM_TRY(tc_gemm(tp, st));
        VGM_TRY(ir_dir(ib, n, k.n_act, T, ld, 1, zb16, st));
        swap_pq();
        for (int it = 0; it < inner; ++it) {
          const int last = (it == inner - 1) ? 1 : 0;
          to.A0 = ib.p_hi; to.A1 = ib.p_lo;
          VGM_TRY(tc_gemm(to, st));
          VGM_TRY(ir_update(ib, n, k.n_act, T, ld, last, it == 0 ? 1 : 0, ops->G_shift, ops->precond_c, st));
          if (last) break;
          VGM_TRY(tc_gemm(tp, st));
          VGM_TRY(ir_dir(ib, n, k.n_act, T, ld, 0, zb16, st));
          swap_pq();
        }
        return ir_finish(ib, cb.P, X, slot_state, k.act, n, k.n_act, T, ld, st);
      };
      cnt.launches += 4 * inner + 2; cnt.gemm += 2 * inner;
      const bool use_graph = !g_no_graph && !profile_enabled() && k.n_upper <= IR_GRAPH_MAX_ROWS;
      if (!use_graph) {
        VGM_TRY(enqueue_correction(k.n_upper, k.st));
        continue;
      }
      IrGraphKey key;
      memset(&key, 0, sizeof(key));
      key.r32 = k.ir.r32; key.act = k.act; key.n_act = k.n_act; key.G = ops->G_bf16; key.M = ops->M_bf16x2; key.X = X;
      key.Lambda = Lambda; key.R = cb.R; key.P = cb.P; key.rr = cb.rr; key.slot_state = slot_state;
      key.n = min(k.n_rows, IR_GRAPH_MAX_ROWS); key.T = T; key.ld = ld; key.inner = inner_now;
      key.band_G = ops->band_G; key.band_M = ops->band_M_lp; key.shift = ops->G_shift; key.c = ops->precond_c;
      cudaGraphExec_t exec = nullptr;
      for (int i = 0; i < g_ir_graph_count; ++i)
        if (g_ir_graphs[i].key == key) exec = g_ir_graphs[i].exec;
      if (!exec) {
        cudaGraph_t graph = nullptr;
        // the caller's stream may be the legacy default stream, which cannot be captured: record the sequence on
        // an internal stream (nothing executes during capture) and replay it wherever the solve runs
        cudaStream_t cap = as.side[IR_MAX_CHUNKS - 1];
        VGM_CUDA_OK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
        const int rc_cap = enqueue_correction(key.n, cap);
        const cudaError_t e_cap = cudaStreamEndCapture(cap, &graph);
        if (rc_cap != VGM_OK) { if (graph) cudaGraphDestroy(graph); return rc_cap; }
        VGM_CUDA_OK(e_cap);
        VGM_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        VGM_CUDA_OK(cudaGraphDestroy(graph));
        IrGraphEntry& slot_e = g_ir_graphs[g_ir_graph_next];
        if (g_ir_graph_count == IR_GRAPH_CACHE && slot_e.exec) cudaGraphExecDestroy(slot_e.exec);
        slot_e.key = key; slot_e.exec = exec;
        g_ir_graph_next = (g_ir_graph_next + 1) % IR_GRAPH_CACHE;
        g_ir_graph_count = min(g_ir_graph_count + 1, IR_GRAPH_CACHE);
      }
      VGM_CUDA_OK(cudaGraphLaunch(exec, k.st));
    }
    if (!any) break;
  }
  for (int c = 1; c < n_chunks; ++c) {
    VGM_CUDA_OK(cudaEventRecord(as.join[c], as.side[c]));
    VGM_CUDA_OK(cudaStreamWaitEvent(s, as.join[c], 0));
  }
  return report_convergence(cb, w, rows, n_rows, tol, max_outer * inner, info, cnt, s, /*final=*/false);
}

static int pcg_solve(const vgm_operators* ops, int ruu_const, const double* Ruu, const double* Lambda,
                     const double* rhs, double* X, const int* slot_state, const int* slot_has_self, const int* rows,
                     int n_rows, int n_hidden, int T, int ld, const vgm_solver_opts* opts, PcgWs& w,
                     vgm_sweep_info* info, Counters& cnt, cudaStream_t s) {
  if (n_rows == 0) return VGM_OK;
  VGM_REQUIRE(ops && ops->D && ops->DT, "operators D and DT are required");
  VGM_REQUIRE(!ruu_const || ops->M, "shared operator M is required when R_uu is constant");
  VGM_REQUIRE(ruu_const || Ruu, "R_uu array is required when it is not constant");
  // short time grids: one CTA per system forms its matrix in shared memory and solves it directly
  if (!(opts && opts->dense_short_grids < 0) && dense_solve_applicable(ops, ruu_const, T)) {
    DenseSolveArgs da;
    memset(&da, 0, sizeof(da));
    da.M = ruu_const ? ops->M : nullptr;
    da.D = ops->D; da.DT = ops->DT; da.DtD = ops->DtD; da.T = T; da.ld = ld;
    da.Ruu = Ruu; da.Lambda = Lambda; da.rhs = rhs; da.X = X;
    da.slot_state = slot_state; da.slot_has_self = slot_has_self; da.rows = rows;
    da.bad = w.cb.n_act + 32;
    VGM_CUDA_OK(cudaMemsetAsync(da.bad, 0, sizeof(int), s));
    VGM_TRY(dense_solve(da, n_rows, s));
    int bad = 0;
    VGM_CUDA_OK(cudaMemcpyAsync(&bad, da.bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGM_CUDA_OK(cudaStreamSynchronize(s));
    cnt.launches += 1;
    if (bad == 0) {
      if (info) info->iterations = max(info->iterations, 1);
      return VGM_OK;
    }
    if (g_debug) fprintf(stderr, "[vgm] dense solver: non-positive pivot in slot %d, iterative solver takes over\n", bad - 1);
  }
  // a handful of systems (the wrap-around state, a level of an all-hidden chain): one CTA per system does the
  // whole banded-Cholesky + refinement solve in a single launch
  if (opts && opts->small_level_solver && ops->M_bf16x2 && small_solve_applicable(ops, ruu_const, n_rows, T)) {
    SmallSolveArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.M = ops->M; sa.T = T; sa.ld = ld; sa.band_M = ops->band_M;
    sa.Lambda = Lambda; sa.rhs = rhs; sa.X = X; sa.slot_state = slot_state; sa.slot_has_self = slot_has_self;
    sa.rows = rows; sa.Lscratch = w.small_scratch;
    sa.tol = (opts && opts->tol > 0) ? opts->tol : 1e-13;
    sa.max_pass = 8;
    sa.done = w.cb.done; sa.iters = w.cb.iters;
    sa.status = (int*)((char*)w.small_scratch + small_solve_scratch_bytes(T) - 256);
    cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;
    if (g_debug) { cudaEventCreate(&dbg0); cudaEventCreate(&dbg1); cudaEventRecord(dbg0, s); }
    VGM_TRY(small_solve(sa, n_rows, s));
    if (g_debug) cudaEventRecord(dbg1, s);
    int st[SMALL_SOLVE_MAX_SYSTEMS];
    VGM_CUDA_OK(cudaMemcpyAsync(st, sa.status, n_rows * sizeof(int), cudaMemcpyDeviceToHost, s));
    VGM_CUDA_OK(cudaStreamSynchronize(s));
    cnt.launches += 1;
    bool ok = true;
    for (int i = 0; i < n_rows; ++i) ok = ok && st[i] == 0;
    if (g_debug) {
      int h_rows0 = 0, h_it = -1;
      cudaMemcpy(&h_rows0, rows, sizeof(int), cudaMemcpyDeviceToHost);
      cudaMemcpy(&h_it, w.cb.iters + h_rows0, sizeof(int), cudaMemcpyDeviceToHost);
      float dbg_ms = 0.f;
      cudaEventElapsedTime(&dbg_ms, dbg0, dbg1);
      cudaEventDestroy(dbg0); cudaEventDestroy(dbg1);
      fprintf(stderr, "[vgm] tiny level: %d system(s), single-CTA solver %s (status %d, %d passes, %.3f ms)\n", n_rows,
              ok ? "converged" : "fell back", st[0], h_it, dbg_ms);
    }
    if (ok) {
      if (info) info->iterations = max(info->iterations, 1);
      return VGM_OK;
    }
    // fall through to the batched solver (x was left untouched for the systems that did not converge)
  }
  // mixed precision needs the bf16 operators and tiles that the TMA boxes (128 x 64) fit into
  bool after_mixed = false;
  if (ruu_const && ops->G_bf16 && ops->M_bf16x2 && T >= 128) {
    int r = ir_solve(ops, Lambda, rhs, X, slot_state, slot_has_self, rows, n_rows, T, ld, opts, w, info, cnt, s);
    if (r != VGM_ENOTCONV) return r;
    // stagnated rows: finish them with FP64 preconditioned CG, warm-started from the refined x (rows that did
    // converge are recognised by their residual in cg_init_kernel and cost one operator application)
    if (g_debug) fprintf(stderr, "[vgm] mixed-precision refinement stagnated on some systems: FP64 PCG takes over\n");
    after_mixed = true;
  }
  const double tol = (opts && opts->tol > 0) ? opts->tol : 1e-13;
  const int max_iter = (opts && opts->max_iter > 0 && !after_mixed) ? opts->max_iter : 4 * T;
  const int check_every = (opts && opts->check_every > 0) ? opts->check_every : 8;
  const int precond = (ruu_const && ops->G) ? 1 : 0;
  const double tol2 = tol * tol;
  CgBuf& cb = w.cb;
  cb.n_tiles = gemm_dot_tiles(T, n_rows);
  cb.n_tiles_z = cb.n_tiles;
  const int th = row_threads(T);

  VGM_CUDA_OK(cudaMemcpyAsync(cb.act[0], rows, (size_t)n_rows * sizeof(int), cudaMemcpyDeviceToDevice, s));
  cg_set_count_kernel<<<1, 1, 0, s>>>(cb.n_act, n_rows);
  int cur = 0;
  const int* act = cb.act[0];
  cg_gather_x_kernel<<<n_rows, th, 0, s>>>(X, slot_state, act, cb.n_act, T, ld, cb.P);
  VGM_LAUNCH_OK();
  VGM_TRY(apply_operator(ops, ruu_const, Ruu, Lambda, cb, act, n_rows, T, ld, false, cnt, s));
  cg_init_kernel<<<n_rows, th, 0, s>>>(cb, rhs, Lambda, X, slot_state, slot_has_self, act, T, ld, tol2, precond);
  VGM_LAUNCH_OK();
  cnt.launches += 4;
  if (precond) {
    VGM_TRY(apply_preconditioner(ops, cb, act, n_rows, T, ld, cnt, s));
    cg_init_precond_kernel<<<n_rows, th, 0, s>>>(cb, act, T, ld);
    VGM_LAUNCH_OK();
    cnt.launches += 1;
  }

  int n_upper = n_rows;  // host-side upper bound of the active count
  int it = 0;
  while (it < max_iter && n_upper > 0) {
    const int burst = min(check_every, max_iter - it);
    for (int b = 0; b < burst; ++b, ++it) {
      VGM_TRY(apply_operator(ops, ruu_const, Ruu, Lambda, cb, act, n_upper, T, ld, true, cnt, s));
      cg_update_kernel<<<n_upper, th, 0, s>>>(cb, X, slot_state, act, T, ld, tol2, precond);
      VGM_LAUNCH_OK();
      cnt.launches += 1;
      if (precond) {
        VGM_TRY(apply_preconditioner(ops, cb, act, n_upper, T, ld, cnt, s));
        cg_precond_dir_kernel<<<n_upper, th, 0, s>>>(cb, act, T, ld);
        VGM_LAUNCH_OK();
        cnt.launches += 1;
      }
    }
    // drop converged rows, then poll the new count
    VGM_TRY(compact_and_poll(cb, cur, act, n_upper, cnt, s));
    // the GEMM tile configuration (hence the number of partial dots per row) follows the host bound
    cb.n_tiles = gemm_dot_tiles(T, n_upper > 0 ? n_upper : 1);
    cb.n_tiles_z = cb.n_tiles;
  }
  return report_convergence(cb, w, rows, n_rows, tol, max_iter, info, cnt, s);
}

// ------------------------------------------------------------------------------------
struct SweepWs {
  double *Lambda, *rhs, *ruu, *Ruu, *Rlive, *DXlive;
  double* tail_rhs;  // [SMALL_SOLVE_MAX_SYSTEMS][ld] right-hand sides of the tail systems before the live terms
  int* ident;  // identity row map for the live buffer
  PcgWs pcg;
};

static size_t sweep_ws_bytes(const vgm_sweep_plan* pl) {
  const size_t mat = align_up((size_t)pl->n_hidden * pl->ld * sizeof(double), 256);
  const size_t live_pairs = align_up((size_t)max(pl->n_live_pairs, 1) * pl->ld * sizeof(double), 256);
  const size_t live = align_up((size_t)max(pl->n_live, 1) * pl->ld * sizeof(double), 256);
  const size_t tail = align_up((size_t)SMALL_SOLVE_MAX_SYSTEMS * pl->ld * sizeof(double), 256);
  return (pl->ruu_const ? 3 : 4) * mat + live_pairs + live + tail + pcg_ws_bytes(pl->n_hidden, pl->T, pl->ld) + 512;
}

static int sweep_carve(void* ws, size_t ws_bytes, const vgm_sweep_plan* pl, SweepWs& w) {
  VGM_REQUIRE(ws && ws_bytes >= sweep_ws_bytes(pl), "sweep workspace too small");
  VGM_REQUIRE(((uintptr_t)ws % 256) == 0, "workspace must be 256-byte aligned");
  char* p = (char*)ws;
  const size_t mat = align_up((size_t)pl->n_hidden * pl->ld * sizeof(double), 256);
  const size_t live_pairs = align_up((size_t)max(pl->n_live_pairs, 1) * pl->ld * sizeof(double), 256);
  const size_t live = align_up((size_t)max(pl->n_live, 1) * pl->ld * sizeof(double), 256);
  w.Lambda = (double*)p; p += mat;
  w.rhs = (double*)p; p += mat;
  w.ruu = (double*)p; p += mat;
  if (pl->ruu_const) {
    w.Ruu = nullptr;
  } else {
    w.Ruu = (double*)p; p += mat;
  }
  w.Rlive = (double*)p; p += live_pairs;
  w.DXlive = (double*)p; p += live;
  w.tail_rhs = (double*)p; p += align_up((size_t)SMALL_SOLVE_MAX_SYSTEMS * pl->ld * sizeof(double), 256);
  const size_t used = (size_t)(p - (char*)ws);
  return pcg_carve(p, ws_bytes - used, pl->n_hidden, pl->T, pl->ld, w.pcg);
}

// Phase 1 of a sweep: snapshot coefficients + the -B_uu^T r_uu term of every hidden state.
static int state_sweep_prepare(const vgm_program* prog, const vgm_sweep_plan* pl, const vgm_operators* ops,
                               const double* X, const double* DX0, const double* theta_star, void* ws,
                               size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_REQUIRE(prog && pl && ops && X && DX0 && theta_star, "null argument");
  VGM_REQUIRE(pl->T > 0 && pl->ld >= pl->T && (pl->ld % 2) == 0, "plan shape");
  if (info) memset(info, 0, sizeof(*info));
  if (pl->n_hidden == 0) return VGM_OK;
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  const int T = pl->T, ld = pl->ld;
  // snapshot coefficients for every hidden state (R, r evaluated on the sweep-start X, :191)
  VGM_TRY(program_state_assemble(prog, pl, X, DX0, theta_star, w.Lambda, w.rhs, w.ruu, w.Ruu, w.Rlive, s));
  // rhs += -B_uu^T r_uu = D^T r_uu - R_uu o r_uu                                    (:231)
  vgm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = w.ruu; g.B = ops->DT; g.C = w.rhs; g.C0 = w.rhs;
  g.lda = g.ldb = g.ldc = ld;
  g.N = pl->n_hidden; g.M = T; g.K = T;
  g.alpha = 1.0; g.beta = 1.0;
  g.U1 = pl->ruu_const ? nullptr : w.Ruu;
  g.V1 = w.ruu;
  g.g1 = pl->ruu_const ? -pl->ruu_value : -1.0;
  if (ops->obs_rhs) { g.V2 = ops->obs_rhs; g.g2 = 1.0; }      // + Q mu_u (opt-in observation term)
  g.band = ops->band_D;
  VGM_TRY(dgemm_nt(g, s));
  if (info) { info->gemm_calls += 1; info->kernel_launches += 2; }
  return VGM_OK;
}

// Phase 1 for the hidden slots [s0, s1) only (pipelined host iteration: rows arrive chunk by chunk).
int state_sweep_prepare_range(const vgm_program* prog, const vgm_sweep_plan* pl, const vgm_operators* ops,
                              const double* X, const double* DX0, const double* theta_star, int s0, int s1, void* ws,
                              size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_REQUIRE(prog && pl && ops && X && DX0 && theta_star && 0 <= s0 && s0 <= s1 && s1 <= pl->n_hidden, "slot range");
  if (s1 == s0) return VGM_OK;
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  const int T = pl->T, ld = pl->ld;
  const size_t o = (size_t)s0 * ld;
  vgm_sweep_plan sub = *pl;
  sub.n_hidden = s1 - s0;
  sub.pair_ptr = pl->pair_ptr + s0;
  VGM_TRY(program_state_assemble(prog, &sub, X, DX0, theta_star, w.Lambda + o, w.rhs + o, w.ruu + o,
                                 w.Ruu ? w.Ruu + o : nullptr, w.Rlive, s));
  vgm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = w.ruu + o; g.B = ops->DT; g.C = w.rhs + o; g.C0 = w.rhs + o;
  g.lda = g.ldb = g.ldc = ld;
  g.N = s1 - s0; g.M = T; g.K = T;
  g.alpha = 1.0; g.beta = 1.0;
  g.U1 = pl->ruu_const ? nullptr : w.Ruu + o;
  g.V1 = w.ruu + o;
  g.g1 = pl->ruu_const ? -pl->ruu_value : -1.0;
  if (ops->obs_rhs) { g.V2 = ops->obs_rhs + o; g.g2 = 1.0; }
  g.band = ops->band_D;
  VGM_TRY(dgemm_nt(g, s));
  if (info) { info->gemm_calls += 1; info->kernel_launches += 2; }
  return VGM_OK;
}

// Solve rows [r0, r1) of dependency level `lev` (positions within the level); no live-term refresh, so only
// meaningful for level 0 or after state_sweep_level has patched the right-hand sides.
int state_sweep_level_rows(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                           int lev, int r0, int r1, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_REQUIRE(pl && ops && X && lev >= 0 && lev < pl->n_levels, "level out of range");
  const int base = pl->level_ptr_host[lev], n_lev = pl->level_ptr_host[lev + 1] - base;
  VGM_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= n_lev, "row range");
  if (r1 == r0) return VGM_OK;
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  Counters cnt;
  int rc = pcg_solve(ops, pl->ruu_const, w.Ruu, w.Lambda, w.rhs, X, pl->slot_state, pl->slot_has_self,
                     pl->level_slots + base + r0, r1 - r0, pl->n_hidden, pl->T, pl->ld, opts, w.pcg, info, cnt, s);
  if (info) { info->gemm_calls += cnt.gemm; info->kernel_launches += cnt.launches; }
  return rc;
}

// Solve an explicit list of level-0 systems (device array of hidden slots); used by the pipelined host iteration
// to solve its last range together with the cyclic seam.
int state_sweep_row_list(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                         const int* rows_dev, int n_rows, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                         cudaStream_t s) {
  VGM_REQUIRE(pl && ops && X && rows_dev && n_rows >= 0 && n_rows <= pl->n_hidden, "row list");
  if (n_rows == 0) return VGM_OK;
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  Counters cnt;
  int rc = pcg_solve(ops, pl->ruu_const, w.Ruu, w.Lambda, w.rhs, X, pl->slot_state, pl->slot_has_self, rows_dev, n_rows,
                     pl->n_hidden, pl->T, pl->ld, opts, w.pcg, info, cnt, s);
  if (info) { info->gemm_calls += cnt.gemm; info->kernel_launches += cnt.launches; }
  return rc;
}

// Phase 2, once per dependency level (in order): refresh D x_k of the states updated in the
// previous level that this level reads live (:225), patch rhs, solve the level's systems.
static int state_sweep_level(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts,
                             double* X, int lev, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_REQUIRE(pl && ops && X && lev >= 0 && lev < pl->n_levels, "level out of range");
  if (pl->n_hidden == 0) return VGM_OK;
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  Counters cnt;
  const int T = pl->T, ld = pl->ld;
  if (lev > 0 && pl->n_live > 0) {
    const int l0 = pl->live_level_ptr_host[lev - 1], l1 = pl->live_level_ptr_host[lev];
    if (l1 > l0) {
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = X; g.B = ops->D; g.C = w.DXlive + (size_t)l0 * ld;
      g.lda = g.ldb = g.ldc = ld;
      g.N = l1 - l0; g.M = T; g.K = T;
      g.rowsA = pl->live_state + l0;
      g.alpha = 1.0;
      g.band = ops->band_D;
      VGM_TRY(dgemm_nt(g, s));
      cnt.gemm += 1; cnt.launches += 1;
    }
  }
  if (lev > 0 && pl->n_live_pairs > 0) {
    const int q0 = pl->live_pair_level_ptr_host[lev], q1 = pl->live_pair_level_ptr_host[lev + 1];
    if (q1 > q0) {
      add_live_terms_kernel<<<q1 - q0, row_threads(T), 0, s>>>(pl->live_pair_slot, pl->live_pair_src, q0, w.Rlive,
                                                               w.DXlive, T, ld, w.rhs);
      VGM_LAUNCH_OK();
      cnt.launches += 1;
    }
  }
  const int r0 = pl->level_ptr_host[lev], r1 = pl->level_ptr_host[lev + 1];
  int rc = pcg_solve(ops, pl->ruu_const, w.Ruu, w.Lambda, w.rhs, X, pl->slot_state, pl->slot_has_self,
                     pl->level_slots + r0, r1 - r0, pl->n_hidden, T, ld, opts, w.pcg, info, cnt, s);
  if (info) {
    info->gemm_calls += cnt.gemm;
    info->kernel_launches += cnt.launches;
  }
  return rc;
}

int state_sweep_level_full(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                           int lev, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  return state_sweep_level(pl, ops, opts, X, lev, ws, ws_bytes, info, s);
}

// ---- tail overlap (vgm_sweep_plan.n_early) ------------------------------------------------------------------
// dst[i] = src[slot rows[i]] (save) or src[i] -> dst[slot rows[i]] (restore) for a handful of rows
__global__ void tail_rows_copy_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                      const int* __restrict__ rows, int T, int ld, int restore) {
  const size_t a = (size_t)rows[blockIdx.x] * ld, b = (size_t)blockIdx.x * ld;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    if (restore) dst[a + t] = src[b + t]; else dst[b + t] = src[a + t];
  }
}

struct TailAsync {              // created once per process (one process drives one GPU)
  cudaStream_t side = nullptr;  // highest priority: its single-CTA kernels are scheduled ahead of the bulk's CTAs
  cudaEvent_t fork = nullptr, join = nullptr;
  int* status_host = nullptr;   // pinned, 64 ints
};
static TailAsync g_tail;
static int tail_async_init() {
  TailAsync& a = g_tail;
  if (a.side) return VGM_OK;
  int lo = 0, hi = 0;
  VGM_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  VGM_CUDA_OK(cudaStreamCreateWithPriority(&a.side, cudaStreamNonBlocking, hi));
  VGM_CUDA_OK(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
  VGM_CUDA_OK(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
  VGM_CUDA_OK(cudaHostAlloc((void**)&a.status_host, 64 * sizeof(int), cudaHostAllocDefault));
  return VGM_OK;
}

static bool tail_overlap_ok(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts) {
  static const bool off = getenv("VGM_NO_TAIL_OVERLAP") != nullptr;
  if (off || pl->n_early <= 0 || !pl->early_slots || !pl->bulk_slots || pl->n_levels < 2 ||
      pl->n_levels > SMALL_SOLVE_MAX_SYSTEMS || pl->n_early > SMALL_SOLVE_MAX_SYSTEMS)
    return false;
  const int n_l0 = pl->level_ptr_host[1], n_tail = pl->n_hidden - n_l0;
  if (n_tail <= 0 || n_tail > SMALL_SOLVE_MAX_SYSTEMS || n_l0 - pl->n_early <= SMALL_SOLVE_MAX_SYSTEMS) return false;
  // same preconditions as the mixed-precision path whose tiny levels the single-CTA solver replaces
  if (!ops->M_bf16x2 || !ops->G_bf16 || dense_solve_applicable(ops, pl->ruu_const, pl->T)) return false;
  (void)opts;
  return small_solve_applicable(ops, pl->ruu_const, max(pl->n_early, 1), pl->T);
}

// All levels, with the chain "early level-0 systems -> level 1 -> ..." on a side stream (one single-CTA launch per
// link, nothing the host has to wait for) while the rest of level 0 is solved as one batch on `s`.
static int state_sweep_levels_overlapped(const vgm_sweep_plan* pl, const vgm_operators* ops,
                                         const vgm_solver_opts* opts, double* X, void* ws, size_t ws_bytes,
                                         vgm_sweep_info* info, cudaStream_t s) {
  SweepWs w;
  VGM_TRY(sweep_carve(ws, ws_bytes, pl, w));
  VGM_TRY(tail_async_init());
  TailAsync& ta = g_tail;
  const int T = pl->T, ld = pl->ld;
  const int n_l0 = pl->level_ptr_host[1], n_tail = pl->n_hidden - n_l0, n_bulk = n_l0 - pl->n_early;
  const int* tail_rows = pl->level_slots + n_l0;
  Counters cnt;
  VGM_CUDA_OK(cudaEventRecord(ta.fork, s));
  VGM_CUDA_OK(cudaStreamWaitEvent(ta.side, ta.fork, 0));
  cudaStream_t q = ta.side;
  // the live terms are ADDED to rhs: keep the untouched right-hand sides for a clean fallback
  tail_rows_copy_kernel<<<n_tail, row_threads(T), 0, q>>>(w.rhs, w.tail_rhs, tail_rows, T, ld, 0);
  VGM_LAUNCH_OK();
  SmallSolveArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.M = ops->M; sa.T = T; sa.ld = ld; sa.band_M = ops->band_M;
  sa.Lambda = w.Lambda; sa.rhs = w.rhs; sa.X = X; sa.slot_state = pl->slot_state; sa.slot_has_self = pl->slot_has_self;
  sa.Lscratch = w.pcg.small_scratch;
  sa.tol = (opts && opts->tol > 0) ? opts->tol : 1e-13;
  sa.max_pass = 8;
  sa.done = w.pcg.cb.done; sa.iters = w.pcg.cb.iters;
  int* status_dev = (int*)((char*)w.pcg.small_scratch + small_solve_scratch_bytes(T) - 256);   // 64 ints
  VGM_CUDA_OK(cudaMemsetAsync(status_dev, 0, 64 * sizeof(int), q));
  sa.rows = pl->early_slots; sa.status = status_dev;
  VGM_TRY(small_solve(sa, pl->n_early, q));
  cnt.launches += 3;
  for (int lev = 1; lev < pl->n_levels; ++lev) {
    const int l0 = pl->live_level_ptr_host[lev - 1], l1 = pl->live_level_ptr_host[lev];
    if (l1 > l0) {
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = X; g.B = ops->D; g.C = w.DXlive + (size_t)l0 * ld;
      g.lda = g.ldb = g.ldc = ld;
      g.N = l1 - l0; g.M = T; g.K = T;
      g.rowsA = pl->live_state + l0;
      g.alpha = 1.0;
      g.band = ops->band_D;
      VGM_TRY(dgemm_nt(g, q));
      cnt.gemm += 1; cnt.launches += 1;
    }
    const int q0 = pl->live_pair_level_ptr_host[lev], q1 = pl->live_pair_level_ptr_host[lev + 1];
    if (q1 > q0) {
      add_live_terms_kernel<<<q1 - q0, row_threads(T), 0, q>>>(pl->live_pair_slot, pl->live_pair_src, q0, w.Rlive,
                                                               w.DXlive, T, ld, w.rhs);
      VGM_LAUNCH_OK();
      cnt.launches += 1;
    }
    const int r0 = pl->level_ptr_host[lev], r1 = pl->level_ptr_host[lev + 1];
    if (r1 > r0) {
      sa.rows = pl->level_slots + r0; sa.status = status_dev + SMALL_SOLVE_MAX_SYSTEMS * lev;
      VGM_TRY(small_solve(sa, r1 - r0, q));
      cnt.launches += 1;
    }
  }
  VGM_CUDA_OK(cudaMemcpyAsync(ta.status_host, status_dev, 64 * sizeof(int), cudaMemcpyDeviceToHost, q));
  VGM_CUDA_OK(cudaEventRecord(ta.join, q));
  // the bulk of level 0 (polls the device while the chain runs beside it)
  int rc = pcg_solve(ops, pl->ruu_const, w.Ruu, w.Lambda, w.rhs, X, pl->slot_state, pl->slot_has_self, pl->bulk_slots,
                     n_bulk, pl->n_hidden, T, ld, opts, w.pcg, info, cnt, s);
  if (rc != VGM_OK && rc != VGM_ENOTCONV) return rc;
  VGM_CUDA_OK(cudaStreamWaitEvent(s, ta.join, 0));
  VGM_CUDA_OK(cudaEventSynchronize(ta.join));
  bool chain_ok = true;
  for (int i = 0; i < 64; ++i) chain_ok = chain_ok && ta.status_host[i] == 0;
  if (info) { info->gemm_calls += cnt.gemm; info->kernel_launches += cnt.launches; info->iterations = max(info->iterations, 1); }
  if (chain_ok) return rc;
  // the single-CTA solver gave up on a link (it leaves x untouched then): restore the tail's right-hand sides and
  // run the chain through the batched solvers, level by level
  if (g_debug) fprintf(stderr, "[vgm] tail overlap: single-CTA solver did not converge, batched solvers take over\n");
  tail_rows_copy_kernel<<<n_tail, row_threads(T), 0, s>>>(w.tail_rhs, w.rhs, tail_rows, T, ld, 1);
  VGM_LAUNCH_OK();
  Counters c2;
  int r = pcg_solve(ops, pl->ruu_const, w.Ruu, w.Lambda, w.rhs, X, pl->slot_state, pl->slot_has_self, pl->early_slots,
                    pl->n_early, pl->n_hidden, T, ld, opts, w.pcg, info, c2, s);
  if (info) { info->gemm_calls += c2.gemm; info->kernel_launches += c2.launches + 1; }
  if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
  for (int lev = 1; lev < pl->n_levels; ++lev) {
    r = state_sweep_level(pl, ops, opts, X, lev, ws, ws_bytes, info, s);
    if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
  }
  return rc;
}

static int state_sweep_levels(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts,
                              double* X, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_REQUIRE(pl && ops && X, "null argument");
  if (pl->n_hidden == 0) return VGM_OK;
  if (tail_overlap_ok(pl, ops, opts)) return state_sweep_levels_overlapped(pl, ops, opts, X, ws, ws_bytes, info, s);
  int rc = VGM_OK;
  for (int lev = 0; lev < pl->n_levels; ++lev) {
    int r = state_sweep_level(pl, ops, opts, X, lev, ws, ws_bytes, info, s);
    if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
  }
  return rc;
}

static int state_sweep(const vgm_program* prog, const vgm_sweep_plan* pl, const vgm_operators* ops,
                       const vgm_solver_opts* opts, double* X, const double* DX0, const double* theta_star, void* ws,
                       size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  VGM_TRY(state_sweep_prepare(prog, pl, ops, X, DX0, theta_star, ws, ws_bytes, info, s));
  return state_sweep_levels(pl, ops, opts, X, ws, ws_bytes, info, s);
}

// ------------------------------------------------------------------------------------
// P5: Cov_u = sum_{k in c(u)} B_uk^T A B_uk   (proxies...py:239; computed and dropped there)
//   k != u : diag(R_k) A diag(R_k) = A o (R_k R_k^T)
//   k == u : (diag(R_uu) - D)^T A (diag(R_uu) - D), via E = A (diag(R_uu) - D) and B_uu^T E
// ------------------------------------------------------------------------------------
// E[i][j] = A[i][j] * Ruu[j] - AD[i][j]      (in place over AD)
__global__ void cov_e1_kernel(const double* __restrict__ A, const double* __restrict__ Ruu, double* __restrict__ E,
                              int T, int ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < T) E[(size_t)i * ld + j] = A[(size_t)i * ld + j] * Ruu[j] - E[(size_t)i * ld + j];
}
// cov[i][j] = has_self * (Ruu[i] E[i][j] - DtE[i][j]) + A[i][j] * sum_{k != u} R_k[i] R_k[j]
__global__ void cov_e2_kernel(const double* __restrict__ A, const double* __restrict__ Rp, int n_pairs, int self_row,
                              const double* __restrict__ E, double* __restrict__ cov, int T, int ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= T) return;
  double rr = 0.0;
  for (int e = 0; e < n_pairs; ++e)
    if (e != self_row) rr += Rp[(size_t)e * ld + i] * Rp[(size_t)e * ld + j];
  const size_t o = (size_t)i * ld + j;
  double v = A[o] * rr;
  if (self_row >= 0) v += Rp[(size_t)self_row * ld + i] * E[o] - cov[o];
  cov[o] = v;
}

static size_t state_cov_ws_bytes(int T, int ld) {
  return 2 * align_up((size_t)T * ld * sizeof(double), 256) + align_up((size_t)32 * ld * sizeof(double), 256);
}

}  // namespace vgm

using namespace vgm;

extern "C" size_t vgm_state_cov_ws_bytes(int T, int ld) { return state_cov_ws_bytes(T, ld); }

extern "C" int vgm_state_cov(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops,
                             const double* A, const double* X, const double* theta_star, int slot, double* cov,
                             void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(prog && plan && ops && ops->D && ops->DT && A && X && theta_star && cov, "null argument");
  VGM_REQUIRE(slot >= 0 && slot < plan->n_hidden, "slot out of range");
  const int T = plan->T, ld = plan->ld;
  VGM_REQUIRE(ws && ws_bytes >= state_cov_ws_bytes(T, ld) && ((uintptr_t)ws % 256) == 0, "workspace");
  cudaStream_t s = (cudaStream_t)stream;
  // pair range of this slot lives on the device: fetch the two offsets and the self flags
  int ptr2[2];
  VGM_CUDA_OK(cudaMemcpyAsync(ptr2, plan->pair_ptr + slot, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  const int n_pairs = ptr2[1] - ptr2[0];
  VGM_REQUIRE(n_pairs <= 32, "more than 32 coupled ODEs for one state");
  int self_flags[32];
  int self_row = -1;
  if (n_pairs > 0) {
    VGM_CUDA_OK(cudaMemcpyAsync(self_flags, plan->pair_self + ptr2[0], n_pairs * sizeof(int), cudaMemcpyDeviceToHost, s));
    VGM_CUDA_OK(cudaStreamSynchronize(s));
    for (int e = 0; e < n_pairs; ++e)
      if (self_flags[e]) self_row = e;
  }
  char* p = (char*)ws;
  const size_t mat = align_up((size_t)T * ld * sizeof(double), 256);
  double* E = (double*)p; p += mat;
  double* ET = (double*)p; p += mat;
  double* Rp = (double*)p;
  VGM_TRY(program_state_pair_R(prog, plan, X, theta_star, slot, Rp, s));
  dim3 eb(128), eg(ceil_div(T, 128), T);
  if (self_row >= 0) {
    vgm_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.lda = g.ldb = g.ldc = ld;
    g.N = T; g.M = T; g.K = T;
    g.alpha = 1.0;
    g.A = A; g.B = ops->DT; g.C = E;                 // A D
    VGM_TRY(dgemm_nt(g, s));
    cov_e1_kernel<<<eg, eb, 0, s>>>(A, Rp + (size_t)self_row * ld, E, T, ld);
    VGM_LAUNCH_OK();
    VGM_TRY(transpose(E, ld, ET, ld, T, T, s));
    g.A = ops->DT; g.B = ET; g.C = cov;              // D^T E
    VGM_TRY(dgemm_nt(g, s));
  }
  cov_e2_kernel<<<eg, eb, 0, s>>>(A, Rp, n_pairs, self_row, E, cov, T, ld);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

extern "C" size_t vgm_pcg_ws_bytes(int n_hidden, int T, int ld) { return pcg_ws_bytes(n_hidden, T, ld); }

extern "C" int vgm_pcg_solve(const vgm_operators* ops, int ruu_const, const double* Ruu, const double* Lambda,
                             const double* rhs, double* X, const int* slot_state, int n_hidden, const int* rows,
                             int n_rows, int T, int ld, const vgm_solver_opts* opts, void* ws, size_t ws_bytes,
                             vgm_sweep_info* info, vgm_stream_t stream) {
  VGM_REQUIRE(Lambda && rhs && X && slot_state && rows && n_rows >= 0 && n_hidden >= n_rows, "null argument");
  VGM_TRY(device_guard());
  PcgWs w;
  VGM_TRY(pcg_carve(ws, ws_bytes, n_hidden, T, ld, w));
  if (info) memset(info, 0, sizeof(*info));
  Counters cnt;
  int r = pcg_solve(ops, ruu_const, Ruu, Lambda, rhs, X, slot_state, nullptr, rows, n_rows, n_hidden, T, ld, opts, w,
                    info, cnt, (cudaStream_t)stream);
  if (info) { info->gemm_calls = cnt.gemm; info->kernel_launches = cnt.launches; }
  return r;
}

extern "C" int vgm_state_sweep_prepare(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops,
                                       const double* X, const double* DX0, const double* theta_star, void* ws,
                                       size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream) {
  VGM_TRY(device_guard());
  return state_sweep_prepare(prog, plan, ops, X, DX0, theta_star, ws, ws_bytes, info, (cudaStream_t)stream);
}

extern "C" int vgm_state_sweep_level(const vgm_sweep_plan* plan, const vgm_operators* ops, const vgm_solver_opts* opts,
                                     double* X, int level, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                                     vgm_stream_t stream) {
  VGM_TRY(device_guard());
  return state_sweep_level(plan, ops, opts, X, level, ws, ws_bytes, info, (cudaStream_t)stream);
}

extern "C" int vgm_state_sweep_levels(const vgm_sweep_plan* plan, const vgm_operators* ops, const vgm_solver_opts* opts,
                                      double* X, void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream) {
  VGM_TRY(device_guard());
  return state_sweep_levels(plan, ops, opts, X, ws, ws_bytes, info, (cudaStream_t)stream);
}

extern "C" size_t vgm_state_sweep_ws_bytes(const vgm_sweep_plan* plan) { return plan ? sweep_ws_bytes(plan) : 0; }

extern "C" int vgm_state_sweep(const vgm_program* prog, const vgm_sweep_plan* plan, const vgm_operators* ops,
                               const vgm_solver_opts* opts, double* X, const double* DX0, const double* theta_star,
                               void* ws, size_t ws_bytes, vgm_sweep_info* info, vgm_stream_t stream) {
  VGM_TRY(device_guard());
  return state_sweep(prog, plan, ops, opts, X, DX0, theta_star, ws, ws_bytes, info, (cudaStream_t)stream);
}
