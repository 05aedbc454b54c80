// P4 for short time grids (config C4: Protein Transduction, T = 99, 131 072 systems per sweep): the reference
// forms  B_uu^T B_uu + diag(Lambda_u)  and calls np.linalg.solve per state (proxies...py:228-262).  For T <= 160 the
// whole system fits one CTA's shared memory, so each CTA forms its matrix and solves it directly:
//
//   M_u = (diag R - D)^T (diag R - D) + diag(Lambda)
//       = D^T D - diag(R) D - D^T diag(R) + diag(R^2 + Lambda)          (D^T D is shared: vgm_operators.DtD)
// (with a constant R_uu = c the shared  M = (c I - D)^T (c I - D)  is read instead)
//
// Only the lower block triangle is formed (three L2-resident reads per entry), the right-hand side rides along as
// one more row (forward substitution for free) and a blocked back substitution finishes.  FP64 throughout, the same
// backward-stable direct solve the reference's LAPACK call performs; a non-positive pivot is reported and the
// caller falls back to the iterative solver.
#include "dense_solve.cuh"
#include "vgm_common.cuh"
#include "vgm_profile.cuh"

namespace vgm {

// ------------------------------------------------------------------------------------------------------------
// Right-looking Cholesky on 8 x 8 blocks.  The lower block triangle lives in shared
// memory block by block (64 contiguous doubles each: 46 KB at T = 99, four CTAs per SM); per panel
//   1. warp 0 factorises the diagonal block in registers (lane = row, pivot rows travel by shuffle, rsqrt instead
//      of sqrt + divide),
//   2. one thread per row below solves its 8 panel entries against it (the right-hand side is one more row, so the
//      forward substitution rides along),
//   3. the trailing blocks get  C -= A B^T  on the FP64 tensor cores: one m8n8k4 DMMA pair per 8 x 8 block, the
//      fragments are single 16-byte loads because a block is stored exactly in fragment order.
// Two barriers per panel (the next diagonal block is factorised by warp 0 inside step 3).  A first version with one
// thread per row and one barrier per column (left-looking, packed triangle) needed 3x the instructions and 2.6x the
// shared-memory wavefronts: 26 ms against 10.5 ms for the 131 072 systems of config C4.
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
      if (J < I) {
        // off-diagonal block: every entry is below the diagonal; lane = (row lane / 4, column pair lane % 4)
        const int r = lane >> 2, cc = 2 * (lane & 3), i = 8 * I + r, j = 8 * J + cc;
        double2 v = make_double2(0.0, 0.0);
        if (i < T) {
          const size_t o = (size_t)i * ld + j;
          if (a.M) {
            v = *reinterpret_cast<const double2*>(a.M + o);
          } else {
            const double2 pv = *reinterpret_cast<const double2*>(a.DtD + o), dv = *reinterpret_cast<const double2*>(a.D + o),
                          tv = *reinterpret_cast<const double2*>(a.DT + o);
            const double ri = Rv[i];
            v.x = pv.x - ri * dv.x - Rv[j] * tv.x;
            v.y = pv.y - ri * dv.y - Rv[j + 1] * tv.y;
          }
        }
        *reinterpret_cast<double2*>(blk + r * 8 + cc) = v;
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = lane + 32 * h, r = e >> 3, cc = e & 7, i = 8 * I + r, j = 8 * J + cc;
          double v = (i == j) ? 1.0 : 0.0;
          if (i < T && j <= i) {
            const size_t o = (size_t)i * ld + j;
            v = a.M ? a.M[o] : a.DtD[o] - Rv[i] * a.D[o] - Rv[j] * a.DT[o];
            if (j == i) v += (a.M ? 0.0 : Rv[i] * Rv[i]) + a.Lambda[so + i];
          }
          blk[e] = v;
        }
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
        // column-oriented so that only one multiply and one FMA per step sit on the dependent chain
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          x[cc] *= dinv[c0 + cc];
#pragma unroll
          for (int k = cc + 1; k < 8; ++k) x[k] -= x[cc] * Lpp[k * 8 + cc];
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
      for (int k = 0; k < 8; ++k) x[k] = y[c0 + k];
#pragma unroll
      for (int cc = 7; cc >= 0; --cc) {
        x[cc] *= dinv[c0 + cc];
#pragma unroll
        for (int k = 0; k < cc; ++k) x[k] -= Lpp[cc * 8 + k] * x[cc];
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

static const bool g_no_dense = getenv("VGM_NO_DENSE") != nullptr;

bool dense_solve_applicable(const vgm_operators* ops, int ruu_const, int T) {
  if (g_no_dense || !ops || T < 2 || T > DENSE_MAX_T) return false;
  return ruu_const ? ops->M != nullptr : (ops->DtD && ops->D && ops->DT);
}

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

int dense_solve(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  if (n_sys <= 0) return VGM_OK;
  VGM_REQUIRE(a.T <= DENSE_MAX_T, "dense solver: T too large");
  const int rows = (a.T + 7) / 8 * 8 - 8 + 1;            // panel rows of the first panel + the right-hand side
  // more warps than panel rows on purpose: the trailing blocks are dealt to all warps but warp 0 (256 threads are
  // 6 % faster than 128 at T = 99)
  (void)rows;
  if (rows <= 64) return launch_blocked<128>(a, n_sys, s);
  return launch_blocked<256>(a, n_sys, s);
}

}  // namespace vgm
