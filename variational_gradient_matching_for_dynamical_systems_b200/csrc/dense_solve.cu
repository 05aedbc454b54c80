// P4 for short time grids (config C4: Protein Transduction, T = 99, 131 072 systems per sweep): the reference
// forms  B_uu^T B_uu + diag(Lambda_u)  and calls np.linalg.solve per state (proxies...py:228-262).  For T <= 160 the
// whole system fits one CTA's shared memory, so each CTA forms its matrix and solves it directly:
//
//   M_u = (diag R - D)^T (diag R - D) + diag(Lambda)
//       = D^T D - diag(R) D - D^T diag(R) + diag(R^2 + Lambda)          (D^T D is shared: vgm_operators.DtD)
// (with a constant R_uu = c the shared  M = (c I - D)^T (c I - D)  is read instead)
//
// Only the lower triangle is formed (three L2-resident reads per entry) and kept packed, factorised left-looking in
// shared memory, the right-hand side rides along as row T (forward substitution for free), and a
// right-looking back substitution finishes.  FP64 throughout; a non-positive pivot is reported and the caller falls back to the iterative solver.
#include "dense_solve.cuh"
#include "vgm_common.cuh"
#include "vgm_profile.cuh"

namespace vgm {

// Factorisation: M = W Dg^-1 W^T with W lower triangular and Dg = diag(W) (Cholesky without the square roots, so a
// column needs nothing from the pivot thread before it is complete and ONE barrier per column suffices):
//   W[i][c] = M[i][c] - sum_{k<c} W[i][k] W[c][k] / W[k][k]                        i >= c
// Thread i owns row i (row T carries the right-hand side).  The lower triangle is stored packed, rows padded to an
// even length (16-byte loads): 41 KB at T = 99, so five CTAs share an SM and hide each other's dependent chains.
// The kernel is shared-memory bound, so the scaled pivot row  S[k] = W[c][k] / W[k][k]  of the NEXT column is
// prepared by the threads whose rows are already finished (thread k < c scales entry k while column c is being
// computed): the inner loop is one 16-byte own-row load, one 16-byte broadcast load and two DFMAs per two k.
__device__ __forceinline__ int tri_row(int i) {      // offset of packed row i (rows padded to even length)
  const int m = i >> 1;
  return (i & 1) ? 2 * (m + 1) * (m + 1) : 2 * m * (m + 1);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) dense_spd_solve_kernel(const DenseSolveArgs a) {
  extern __shared__ double sm[];
  const int T = a.T, ld = a.ld, Tp = (T + 2) & ~1;
  double* W = sm;                                 // packed lower triangle rows 0 .. T, row T = right-hand side
  double* S = W + tri_row(T + 1);                 // 2 x Tp  scaled pivot rows (double buffered)
  double* Rv = S + 2 * Tp;                        // Tp    R_uu(t)
  double* inv = Rv + Tp;                          // Tp    1 / W[k][k]
  double* xs = inv + Tp;                          // Tp    solution
  const int tid = threadIdx.x;
  const int slot = a.rows[blockIdx.x];
  const size_t so = (size_t)slot * ld;
  double* xout = a.X + (size_t)a.slot_state[slot] * ld;
  const bool has_self = a.slot_has_self ? a.slot_has_self[slot] != 0 : true;
  if (!has_self) {
    // u not in c(u): the reference's aliasing doubles the diagonal and there is no B_uu term
    for (int t = tid; t < T; t += THREADS) xout[t] = a.rhs[so + t] / (2.0 * a.Lambda[so + t]);
    return;
  }
  const int rowT = tri_row(T);
  for (int t = tid; t < T; t += THREADS) {
    Rv[t] = a.M ? 0.0 : a.Ruu[so + t];
    W[rowT + t] = a.rhs[so + t];
  }
  __syncthreads();
  // lower triangle, row by row: warp lanes run along j (coalesced reads of the L2-resident operators)
  for (int i = tid >> 5; i < T; i += THREADS >> 5) {
    double* wi = W + tri_row(i);
    if (a.M) {
      const double* m = a.M + (size_t)i * ld;
      for (int j = tid & 31; j <= i; j += 32) wi[j] = m[j] + (j == i ? a.Lambda[so + i] : 0.0);
    } else {
      const double ri = Rv[i];
      const double *p = a.DtD + (size_t)i * ld, *d = a.D + (size_t)i * ld, *dt = a.DT + (size_t)i * ld;
      for (int j = tid & 31; j <= i; j += 32) {
        double v = p[j] - ri * d[j] - Rv[j] * dt[j];
        if (j == i) v += ri * ri + a.Lambda[so + i];
        wi[j] = v;
      }
    }
  }
  __syncthreads();
  const int i = tid;
  double* Wi = W + tri_row(min(i, T));
  for (int c = 0; c < T; ++c) {
    const double* Sc = S + (c & 1) * Tp;          // S[k] = W[c][k] inv[k] for k < c - 1 (made during column c - 1)
    if (i >= c && i <= T) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const int n = c - 1;                        // entries served by Sc
      int k = 0;
      for (; k + 4 <= n; k += 4) {
        const double2 w0 = *reinterpret_cast<const double2*>(Wi + k), s0 = *reinterpret_cast<const double2*>(Sc + k);
        const double2 w1 = *reinterpret_cast<const double2*>(Wi + k + 2), s1 = *reinterpret_cast<const double2*>(Sc + k + 2);
        a0 += w0.x * s0.x;
        a1 += w0.y * s0.y;
        a2 += w1.x * s1.x;
        a3 += w1.y * s1.y;
      }
      for (; k < n; ++k) a0 += Wi[k] * Sc[k];
      if (c >= 1) a1 += Wi[c - 1] * (W[tri_row(c) + c - 1] * inv[c - 1]);   // column c - 1 was finished last
      const double v = Wi[c] - ((a0 + a1) + (a2 + a3));
      Wi[c] = v;
      if (i == c) {
        if (!(v > 0.0)) atomicCAS(a.bad, 0, slot + 1);
        inv[c] = 1.0 / v;
      }
    } else if (i < c) {
      // finished rows: scale entry i of the next pivot row (row c + 1 <= T always exists; row T is the rhs)
      S[((c + 1) & 1) * Tp + i] = W[tri_row(c + 1) + i] * inv[i];
    }
    __syncthreads();
  }
  // W Dg^-1 W^T x = b: row T now holds u = (W Dg^-1)^-1 b; back substitution x_j = (u_j - sum_{k>j} W[k][j] x_k) / W[j][j]
  double t = (i < T) ? W[rowT + i] : 0.0;
  for (int c = T - 1; c >= 0; --c) {
    if (i == c) xs[c] = t * inv[c];
    __syncthreads();
    if (i < c) t -= W[tri_row(c) + i] * xs[c];
  }
  if (i < T) xout[i] = xs[i];
}

// ------------------------------------------------------------------------------------------------------------
// Blocked variant (the default): right-looking Cholesky on 8 x 8 blocks.  The lower block triangle lives in shared
// memory block by block (64 contiguous doubles each: 46 KB at T = 99, four CTAs per SM); per panel
//   1. warp 0 factorises the diagonal block in registers (lane = row, pivot rows travel by shuffle, rsqrt instead
//      of sqrt + divide),
//   2. one thread per row below solves its 8 panel entries against it (the right-hand side is one more row, so the
//      forward substitution rides along),
//   3. the trailing blocks get  C -= A B^T  on the FP64 tensor cores: one m8n8k4 DMMA pair per 8 x 8 block, the
//      fragments are single 16-byte loads because a block is stored exactly in fragment order.
// Three barriers per panel instead of one per column, ~4x fewer instructions and ~7x fewer shared-memory
// wavefronts than the column version below.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int blk_off(int I, int J) { return (I * (I + 1) / 2 + J) * 64; }

// Cholesky of one 8 x 8 diagonal block by one warp: lane r (< 8) owns row r, pivot rows travel by shuffle, rsqrt
// replaces sqrt + divide; the strict upper triangle is zeroed and 1 / L[c][c] goes to dinv.
__device__ __forceinline__ void factor_diag_block(double* Lpp, double* dinv, int lane, int* s_bad) {
  const int r = lane & 7;
  double w[8];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const double2 t2 = *reinterpret_cast<const double2*>(Lpp + r * 8 + k);
    w[k] = t2.x;
    w[k + 1] = t2.y;
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    double v = w[cc];
#pragma unroll
    for (int k = 0; k < cc; ++k) v -= w[k] * __shfl_sync(0xffffffffu, w[k], cc);
    const double dd = __shfl_sync(0xffffffffu, v, cc);
    if (!(dd > 0.0) && lane == 0) *s_bad = 1;
    const double si = rsqrt(dd);
    w[cc] = (r == cc) ? dd * si : (r > cc ? v * si : 0.0);
    if (lane == cc) dinv[cc] = si;
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(Lpp + r * 8 + k) = make_double2(w[k], w[k + 1]);
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) dense_spd_solve_blocked_kernel(const DenseSolveArgs a) {
  extern __shared__ double sm[];
  const int T = a.T, ld = a.ld, nb = (T + 7) >> 3, Tp = nb * 8;
  double* Wb = sm;                                // nb (nb + 1) / 2 blocks of 8 x 8
  double* y = Wb + blk_off(nb, 0);                // Tp   right-hand side -> L^-1 b
  double* xs = y + Tp;                            // Tp   solution
  double* Rv = xs + Tp;                           // Tp   R_uu(t)
  double* dinv = Rv + Tp;                         // Tp   1 / L[k][k]
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int slot = a.rows[blockIdx.x];
  const size_t so = (size_t)slot * ld;
  double* xout = a.X + (size_t)a.slot_state[slot] * ld;
  const bool has_self = a.slot_has_self ? a.slot_has_self[slot] != 0 : true;
  if (!has_self) {
    // u not in c(u): the reference's aliasing doubles the diagonal and there is no B_uu term
    for (int t = tid; t < T; t += THREADS) xout[t] = a.rhs[so + t] / (2.0 * a.Lambda[so + t]);
    return;
  }
  if (tid == 0) s_bad = 0;
  for (int t = tid; t < Tp; t += THREADS) {
    Rv[t] = (a.M || t >= T) ? 0.0 : a.Ruu[so + t];
    y[t] = t < T ? a.rhs[so + t] : 0.0;
  }
  __syncthreads();
  // the lower block triangle, one 8 x 8 block per warp pass (two entries per lane, six independent L2 reads in
  // flight); padding rows/columns are the identity
  {
    const int n_blk = nb * (nb + 1) / 2;
    int I = 0, J = 0;                             // block coordinates of flat index b (advanced incrementally)
    for (int b = 0; b < warp; ++b) { if (++J > I) { ++I; J = 0; } }
    for (int b = warp; b < n_blk; b += NW) {
      double* blk = Wb + b * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h, r = e >> 3, cc = e & 7, i = 8 * I + r, j = 8 * J + cc;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < T && j <= i) {
          if (a.M) {
            v = a.M[(size_t)i * ld + j];
          } else {
            const size_t o = (size_t)i * ld + j;
            v = a.DtD[o] - Rv[i] * a.D[o] - Rv[j] * a.DT[o];
            if (j == i) v += Rv[i] * Rv[i];
          }
          if (j == i) v += a.Lambda[so + i];
        }
        blk[e] = v;
      }
      for (int st = 0; st < NW; ++st) { if (++J > I) { ++I; J = 0; } }
    }
  }
  __syncthreads();

  const int g = lane >> 2, q = lane & 3;          // DMMA fragment coordinates
  if (warp == 0) factor_diag_block(Wb, dinv, lane, &s_bad);
  __syncthreads();
  for (int pnl = 0; pnl < nb; ++pnl) {
    const int c0 = 8 * pnl;
    double* Lpp = Wb + blk_off(pnl, pnl);
    // ---- 1. the diagonal block was factorised by warp 0 at the end of the previous panel's step 3 ----
    // ---- 2. panel: rows below the diagonal block, plus the right-hand side as one more row ----
    {
      const int n_below = Tp - c0 - 8;
      if (tid <= n_below) {
        const bool is_rhs = tid == n_below;
        const int i = c0 + 8 + tid;
        double* xp = is_rhs ? y + c0 : Wb + blk_off(i >> 3, pnl) + (i & 7) * 8;
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const double2 t2 = *reinterpret_cast<const double2*>(xp + k);
          x[k] = t2.x;
          x[k + 1] = t2.y;
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          double v = x[cc];
#pragma unroll
          for (int k = 0; k < cc; ++k) v -= x[k] * Lpp[cc * 8 + k];
          x[cc] = v * dinv[c0 + cc];
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(xp + k) = make_double2(x[k], x[k + 1]);
      }
    }
    __syncthreads();
    // ---- 3. trailing update on the tensor cores.  Warp 0 updates the next diagonal block and factorises it right
    //         away (step 1 of the next panel, off the critical path); the other warps share the remaining blocks ----
    {
      if (warp == 0) {
        if (pnl + 1 < nb) {
          const double2 af = *reinterpret_cast<const double2*>(Wb + blk_off(pnl + 1, pnl) + g * 8 + 2 * q);
          double2* cp = reinterpret_cast<double2*>(Wb + blk_off(pnl + 1, pnl + 1) + g * 8 + 2 * q);
          const double2 c2 = *cp;
          double cf[2] = {c2.x, c2.y};
          dmma_8x8x4(cf, -af.x, af.x);
          dmma_8x8x4(cf, -af.y, af.y);
          *cp = make_double2(cf[0], cf[1]);
          __syncwarp();
          factor_diag_block(Wb + blk_off(pnl + 1, pnl + 1), dinv + c0 + 8, lane, &s_bad);
        }
      } else {
        int base = 0;                             // (blocks dealt so far) mod (NW - 1)
        for (int I = pnl + 2; I < nb; ++I) {
          const int n_row = I - pnl;              // blocks (I, pnl+1 .. I)
          int J = pnl + 1 + ((warp - 1 - base + (NW - 1)) % (NW - 1));
          if (J <= I) {
            const double2 af = *reinterpret_cast<const double2*>(Wb + blk_off(I, pnl) + g * 8 + 2 * q);
            const double nx = -af.x, ny = -af.y;
            for (; J <= I; J += NW - 1) {
              const double2 bf = *reinterpret_cast<const double2*>(Wb + blk_off(J, pnl) + g * 8 + 2 * q);
              double2* cp = reinterpret_cast<double2*>(Wb + blk_off(I, J) + g * 8 + 2 * q);
              const double2 c2 = *cp;
              double cf[2] = {c2.x, c2.y};
              dmma_8x8x4(cf, nx, bf.x);
              dmma_8x8x4(cf, ny, bf.y);
              *cp = make_double2(cf[0], cf[1]);
            }
          }
          base = (base + n_row) % (NW - 1);
        }
      }
      // right-hand side: y[j] -= L[j][c0 .. c0+8) . z[c0 .. c0+8)
      const int j = c0 + 8 + tid;
      if (j < Tp) {
        const double* lj = Wb + blk_off(j >> 3, pnl) + (j & 7) * 8;
        double v = y[j];
#pragma unroll
        for (int k = 0; k < 8; ++k) v -= lj[k] * y[c0 + k];
        y[j] = v;
      }
    }
    __syncthreads();
  }
  // ---- back substitution L^T x = z, panel by panel from the bottom.  The 8 entries of a panel are owned by 8
  //      neighbouring threads of one warp, so the thread that solves the next diagonal block only needs a warp-level
  //      sync after that warp's update: one block barrier per panel ----
  for (int pnl = nb - 1; pnl >= 0; --pnl) {
    const int c0 = 8 * pnl;
    const double* Lpp = Wb + blk_off(pnl, pnl);
    if (tid == c0) {
      double x[8];
#pragma unroll
      for (int cc = 7; cc >= 0; --cc) {
        double v = y[c0 + cc];
#pragma unroll
        for (int k = cc + 1; k < 8; ++k) v -= Lpp[k * 8 + cc] * x[k];
        x[cc] = v * dinv[c0 + cc];
        xs[c0 + cc] = x[cc];
      }
    }
    __syncthreads();
    if (tid < c0) {
      const double* col = Wb + blk_off(pnl, tid >> 3) + (tid & 7);
      double v = y[tid];
#pragma unroll
      for (int k = 0; k < 8; ++k) v -= col[k * 8] * xs[c0 + k];
      y[tid] = v;
    }
    __syncwarp();
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicCAS(a.bad, 0, slot + 1);
    return;                                       // leave x untouched for the iterative fallback
  }
  for (int t = tid; t < T; t += THREADS) xout[t] = xs[t];
}

static size_t dense_blocked_smem_bytes(int T) {
  const int nb = (T + 7) / 8;
  return ((size_t)(nb * (nb + 1) / 2) * 64 + 4 * (size_t)nb * 8) * sizeof(double);
}

static size_t dense_smem_bytes(int T) {
  const int n = T + 1, m = n >> 1;                                     // tri_row(T + 1) on the host
  const size_t tri = (n & 1) ? 2 * (size_t)(m + 1) * (m + 1) : 2 * (size_t)m * (m + 1);
  return (tri + 5 * (size_t)((T + 2) & ~1)) * sizeof(double);
}

static const bool g_no_dense = getenv("VGM_NO_DENSE") != nullptr;

bool dense_solve_applicable(const vgm_operators* ops, int ruu_const, int T) {
  if (g_no_dense || !ops || T < 2 || T > DENSE_MAX_T) return false;
  return ruu_const ? ops->M != nullptr : (ops->DtD && ops->D && ops->DT);
}

static const bool g_dense_columns = getenv("VGM_DENSE_COLUMNS") != nullptr;   // A/B switch: the column version

template <int THREADS>
static int launch_blocked(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  const size_t smem = dense_blocked_smem_bytes(a.T);
  static size_t configured = 0;
  if (smem > configured) {
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_blocked_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_blocked_kernel<THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = smem;
  }
  ProfScope prof(PROF_VEC, (double)n_sys * ((double)a.T * a.T * a.T / 3.0 + 4.0 * a.T * a.T), (double)n_sys * a.T * 40.0, s);
  dense_spd_solve_blocked_kernel<THREADS><<<n_sys, THREADS, smem, s>>>(a);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

template <int THREADS>
static int launch(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  const size_t smem = dense_smem_bytes(a.T);
  static size_t configured = 0;
  if (smem > configured) {
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_kernel<THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = smem;
  }
  // forms T^2/2 entries from three operators (L2) and reads/writes 4 vectors; T^3/3 + 2 T^2 flops per system
  ProfScope prof(PROF_VEC, (double)n_sys * ((double)a.T * a.T * a.T / 3.0 + 4.0 * a.T * a.T), (double)n_sys * a.T * 40.0, s);
  dense_spd_solve_kernel<THREADS><<<n_sys, THREADS, smem, s>>>(a);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

int dense_solve(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  if (n_sys <= 0) return VGM_OK;
  VGM_REQUIRE(a.T <= DENSE_MAX_T, "dense solver: T too large");
  if (!g_dense_columns) {
    const int rows = (a.T + 7) / 8 * 8 - 8 + 1;          // panel rows of the first panel + the right-hand side
    if (rows <= 64) return launch_blocked<64>(a, n_sys, s);
    if (rows <= 128) return launch_blocked<128>(a, n_sys, s);
    return launch_blocked<192>(a, n_sys, s);
  }
  if (a.T + 1 <= 64) return launch<64>(a, n_sys, s);
  if (a.T + 1 <= 128) return launch<128>(a, n_sys, s);
  return launch<192>(a, n_sys, s);
}

}  // namespace vgm
