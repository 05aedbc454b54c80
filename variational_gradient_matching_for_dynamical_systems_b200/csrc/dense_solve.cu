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

__device__ __forceinline__ int blk_off(int I, int J) { return (I * (I + 1) / 2 + J) * 64; }

// ------------------------------------------------------------------------------------------------------------
// Right-looking Cholesky on 8 x 8 blocks held in shared memory in DMMA fragment order (64 contiguous doubles per
// block, 54 KB at T = 99: four CTAs per SM).  Built around what the first version's ncu capture showed
// (profiles/r01h_ncu_dense_solver.md: 8-way bank conflicts of a row-per-thread panel solve, a serial 8 x 8 back
// substitution; profiles/r02_ncu_dense_solver.md has the comparison):
//   * warp 0 factorises the diagonal block and INVERTS it in place (L_pp itself is never needed again),
//   * the panel solve  X = A L_pp^-T  is the product  A inv(L_pp)^T : one DMMA pair per block, in place, on blocks
//     stored in fragment order (no thread walks a 64-byte row),
//   * the right-hand side is block row nb of the matrix (row 0 of block (nb, J) holds y[8J .. 8J+8), the other rows
//     are zero), so forward substitution is the same two steps and needs no code of its own,
//   * the trailing blocks are dealt to the warps as contiguous runs of the row-major list (one A fragment per row,
//     pointer increments between blocks); warp 0 updates only the next diagonal block and factorises it while the
//     others work -- it skips the barrier between panel solve and trailing update (named barrier, arrive only),
//   * back substitution: x_p = inv(L_pp)^T y_p by eight threads, then one thread per row above subtracts its
//     8 products; one block barrier per panel,
//   * R_uu, Lambda and the right-hand side are fetched once, together, before the operator reads start; the first
//     diagonal block is factorised by warp 0 while the other warps still form the matrix.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_barrier_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_barrier_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// One warp turns an 8 x 8 SPD block (lower triangle valid, row-major in shared memory) into the inverse of its
// Cholesky factor, in place (zeros above the diagonal): Cholesky with rsqrt pivots (lane = row, pivot rows travel by
// shuffle), then lane c forward-substitutes column c of the inverse.  A division-free elimination with the row
// operations applied to an identity alongside (two dependent FP64 operations per column instead of shuffle, rsqrt,
// shuffle) was measured SLOWER (8.91 vs 8.60 ms for config C4): it needs 30 % more instructions and twice the
// shuffles, and the kernel is bound by the LSU pipe and issue slots, not by this chain (profiles/solver_study.md).
__device__ __forceinline__ void factor_invert_diag_block(double* Lpp, int lane, int* s_bad) {
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
    w[cc] = (r == cc) ? si : v * si;              // the diagonal keeps 1 / L[c][c]; entries above it are never read
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(Lpp + r * 8 + k) = make_double2(w[k], w[k + 1]);
  }
  __syncwarp();
  double x[8];                                    // lane c solves L x = e_c (x[k] = 0 for k < c falls out of the recurrence)
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    double row[8];
#pragma unroll
    for (int k = 0; k <= rr; k += 2) {
      const double2 t2 = *reinterpret_cast<const double2*>(Lpp + rr * 8 + k);
      row[k] = t2.x;
      row[k + 1] = t2.y;
    }
    double sacc = (rr == r) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < rr; ++k) sacc -= row[k] * x[k];
    x[rr] = sacc * row[rr];
  }
  __syncwarp();
  if (lane < 8) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) Lpp[rr * 8 + r] = x[rr];
  }
  __syncwarp();
}

// One entry pair of the system matrix: rows i, columns j, j + 1 (j even).
template <bool HAS_M>
__device__ __forceinline__ double2 matrix_pair(const DenseSolveArgs& a, const double* Rv, size_t o, int i, int j) {
  if (HAS_M) return *reinterpret_cast<const double2*>(a.M + o);
  const double2 pv = *reinterpret_cast<const double2*>(a.DtD + o), dv = *reinterpret_cast<const double2*>(a.D + o),
                tv = *reinterpret_cast<const double2*>(a.DT + o);
  const double ri = Rv[i];
  const double2 rj = *reinterpret_cast<const double2*>(Rv + j);
  return make_double2(pv.x - ri * dv.x - rj.x * tv.x, pv.y - ri * dv.y - rj.y * tv.y);
}

template <int NW, bool HAS_M>
__device__ __forceinline__ void form_matrix(const DenseSolveArgs& a, double* Wb, const double* Rv, const double* Lv, int T, int ld,
                                            int nb, int warp, int lane, int* s_bad) {
  const int g = lane >> 2, q = lane & 3, fo = g * 8 + 2 * q;
  const int n_diag0 = NW > 1 ? (nb + 1) / 2 : nb;             // diagonal blocks of warp 0
  // diagonal blocks
  for (int d = warp == 0 ? 0 : n_diag0 + warp - 1; d < (warp == 0 ? n_diag0 : nb); d += (warp == 0 ? 1 : NW - 1)) {
    const int i = 8 * d + g, j = 8 * d + 2 * q;
    double2 v = make_double2(0.0, 0.0);
    if (i < T) {
      if (j <= i) {
        v = matrix_pair<HAS_M>(a, Rv, (size_t)i * ld + j, i, j);
        const double dg = Rv[i] * Rv[i] + Lv[i];
        if (j == i) {
          v.x += dg;
          v.y = 0.0;
        } else if (j + 1 == i) {
          v.y += dg;
        }
      }
    } else {
      if (j == i) v.x = 1.0;
      if (j + 1 == i) v.y = 1.0;
    }
    *reinterpret_cast<double2*>(Wb + blk_off(d, d) + fo) = v;
    if (d == 0) {
      __syncwarp();
      factor_invert_diag_block(Wb, lane, s_bad);
    }
  }
  // blocks below the diagonal: (I, J), J < I, flat index I (I - 1) / 2 + J dealt round-robin to warps 1 .. NW-1
  if (NW > 1 && warp == 0) return;
  const int n_off = nb * (nb - 1) / 2, stride = NW > 1 ? NW - 1 : 1;
  int I = 1, J = NW > 1 ? warp - 1 : 0;
  while (J >= I) { J -= I; ++I; }
  for (int b = NW > 1 ? warp - 1 : 0; b < n_off; b += stride) {
    const int i = 8 * I + g, j = 8 * J + 2 * q;
    double2 v = make_double2(0.0, 0.0);
    if (i < T) v = matrix_pair<HAS_M>(a, Rv, (size_t)i * ld + j, i, j);
    *reinterpret_cast<double2*>(Wb + blk_off(I, J) + fo) = v;
    J += stride;
    while (J >= I) { J -= I; ++I; }
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 4) dense_spd_solve_inv_kernel(const DenseSolveArgs a) {
  extern __shared__ double sm[];
  const int T = a.T, ld = a.ld, nb = (T + 7) >> 3, Tp = nb * 8;
  const int n_tri = nb * (nb + 1) / 2;
  double* Wb = sm;                                // blocks (I, J): I < nb the lower block triangle, I = nb the right-hand side
  double* yb = Wb + n_tri * 64;                   // block row nb: y[j] at yb[(j >> 3) * 64 + (j & 7)]
  double* Rv = yb + nb * 64;                      // Tp   R_uu(t)
  double* Lv = Rv + Tp;                           // Tp   Lambda(t)
  __shared__ int s_bad;
  constexpr int NW = THREADS / 32;
  __shared__ int s_run[(DENSE_MAX_T / 8) * NW];   // [m][warp]: first item and length of a warp's run of trailing blocks
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = a.rows[blockIdx.x];
  const size_t so = (size_t)slot * ld;
  double* xout = a.X + (size_t)a.slot_state[slot] * ld;
  const bool has_self = a.slot_has_self ? a.slot_has_self[slot] != 0 : true;
  if (!has_self) {
    for (int t = tid; t < T; t += THREADS) xout[t] = a.rhs[so + t] / (2.0 * a.Lambda[so + t]);
    return;
  }
  if (tid == 0) s_bad = 0;
  // the trailing blocks of a panel with m block rows left are items 0 .. n-1 of the row-major list (rows pnl+1 .. nb,
  // columns pnl+1 .. min(row, nb-1)); item 0 is warp 0's, the others go to warps 1 .. NW-1 as contiguous runs
  for (int e = tid; e < nb * NW; e += THREADS) {
    const int m = e / NW, w = e % NW;
    int packed = 0;
    if (m >= 1 && w >= 1) {
      const int n_items = m * (m + 1) / 2 + m, chunk = (n_items - 1 + NW - 2) / (NW - 1), s0 = 1 + (w - 1) * chunk;
      const int cnt = min(chunk, n_items - s0);
      if (cnt > 0) {
        int i = (int)((sqrtf(8.0f * (float)s0 + 1.0f) - 1.0f) * 0.5f);
        if (i * (i + 1) / 2 > s0) --i;
        if ((i + 1) * (i + 2) / 2 <= s0) ++i;
        packed = i | ((s0 - i * (i + 1) / 2) << 8) | (cnt << 16);
      }
    }
    s_run[e] = packed;
  }
  // per-system vectors: one round trip to HBM for all three
  for (int e = tid; e < nb * 64; e += THREADS) {
    const int t = (e >> 6) * 8 + (e & 7);
    yb[e] = ((e & 56) == 0 && t < T) ? a.rhs[so + t] : 0.0;
  }
  for (int t = tid; t < Tp; t += THREADS) {
    Rv[t] = (a.M || t >= T) ? 0.0 : a.Ruu[so + t];
    Lv[t] = t < T ? a.Lambda[so + t] : 0.0;
  }
  __syncthreads();
  const int g = lane >> 2, q = lane & 3, fo = g * 8 + 2 * q;   // DMMA fragment coordinates; a lane's double2 of a block
  // ---- the matrix: lane (g, q) forms entries (g, 2q), (g, 2q + 1) of a block; padding rows/columns are the identity.
  //      Warp 0 forms the first half of the diagonal blocks and factorises block (0, 0) at once; the other warps share
  //      the blocks below the diagonal (no triangle or diagonal logic in that loop) and the remaining diagonal blocks.
  if (a.M) form_matrix<NW, true>(a, Wb, Rv, Lv, T, ld, nb, warp, lane, &s_bad);
  else form_matrix<NW, false>(a, Wb, Rv, Lv, T, ld, nb, warp, lane, &s_bad);
  __syncthreads();
  for (int pnl = 0; pnl < nb; ++pnl) {
    // ---- A. panel: X = A inv(L_pp)^T for the blocks below the diagonal block and for the right-hand side ----
    double2 x0 = make_double2(0.0, 0.0);          // warp 0 keeps block (pnl+1, pnl) for its diagonal update
    {
      int I = pnl + 1 + warp;
      if (I <= nb) {
        const double2 bf = *reinterpret_cast<const double2*>(Wb + blk_off(pnl, pnl) + fo);
        bool first = true;
        for (; I <= nb; I += NW) {
          double2* xp = reinterpret_cast<double2*>(Wb + blk_off(I, pnl) + fo);
          const double2 af = *xp;
          double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
          dmma_8x8x4(c0, af.x, bf.x);
          dmma_8x8x4(c1, af.y, bf.y);
          const double2 xv = make_double2(c0[0] + c1[0], c0[1] + c1[1]);
          *xp = xv;
          if (first) x0 = xv;
          first = false;
        }
      }
    }
    const int m = nb - 1 - pnl;                   // block rows left in the triangle
    if (m == 0) break;                            // last panel: only the right-hand side block, solved by warp 0
    // ---- B. trailing update.  Warp 0 needs only its own block (pnl+1, pnl) and skips the wait ----
    if (warp == 0) {
      __syncwarp();
      named_barrier_arrive(1, THREADS);
      double2* cp = reinterpret_cast<double2*>(Wb + blk_off(pnl + 1, pnl + 1) + fo);
      const double2 c2 = *cp;
      double c0[2] = {c2.x, c2.y}, c1[2] = {0.0, 0.0};
      dmma_8x8x4(c0, -x0.x, x0.x);                // the C fragment of the panel solve is the A / B fragment pair here
      dmma_8x8x4(c1, -x0.y, x0.y);
      *cp = make_double2(c0[0] + c1[0], c0[1] + c1[1]);
      __syncwarp();
      factor_invert_diag_block(Wb + blk_off(pnl + 1, pnl + 1), lane, &s_bad);
    } else {
      named_barrier_sync(1, THREADS);
      const int run = s_run[m * NW + warp];
      int cnt = run >> 16;
      if (cnt > 0) {
        int I = pnl + 1 + (run & 255), J = pnl + 1 + ((run >> 8) & 255);
        const double* Ap = Wb + blk_off(I, pnl) + fo;
        double2 af = *reinterpret_cast<const double2*>(Ap);
        double nx = -af.x, ny = -af.y;
        double* Cp = Wb + blk_off(I, J) + fo;
        const double* Bp = Wb + blk_off(J, pnl) + fo;
        int jmax = min(I, nb - 1);
        for (; cnt > 0; --cnt) {
          const double2 bf = *reinterpret_cast<const double2*>(Bp);
          const double2 c2 = *reinterpret_cast<const double2*>(Cp);
          double cf[2] = {c2.x, c2.y};
          dmma_8x8x4(cf, nx, bf.x);
          dmma_8x8x4(cf, ny, bf.y);
          *reinterpret_cast<double2*>(Cp) = make_double2(cf[0], cf[1]);
          ++J;
          if (J > jmax) {
            ++I;
            J = pnl + 1;
            Ap = Wb + blk_off(I, pnl) + fo;
            if (cnt > 1) {
              af = *reinterpret_cast<const double2*>(Ap);
              nx = -af.x;
              ny = -af.y;
            }
            Cp = const_cast<double*>(Ap) + 64;
            Bp = Wb + blk_off(pnl + 1, pnl) + fo;
            jmax = min(I, nb - 1);
          } else {
            Cp += 64;
            Bp += J * 64;
          }
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // ---- back substitution L^T x = z, panel by panel from the bottom; y lives in row 0 of block row nb.  The eight
  //      entries of a panel are owned by eight neighbouring threads of one warp, which also compute x_p, so only
  //      the hand-over of x_p to the rows above needs a block barrier ----
  for (int pnl = nb - 1; pnl >= 0; --pnl) {
    const int c0 = 8 * pnl;
    double* yp = yb + pnl * 64;
    if (tid >= c0 && tid < c0 + 8) {
      const unsigned mask = 0xffu << (c0 & 31);
      const double* inv = Wb + blk_off(pnl, pnl) + (tid - c0);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        s0 += inv[k * 8] * yp[k];
        s1 += inv[k * 8 + 8] * yp[k + 1];
      }
      __syncwarp(mask);
      yp[tid - c0] = s0 + s1;
    }
    __syncthreads();
    if (tid < c0) {
      const double* col = Wb + blk_off(pnl, tid >> 3) + (tid & 7);
      double* yj = yb + (tid >> 3) * 64 + (tid & 7);
      double s0 = *yj, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        s0 -= col[k * 8] * yp[k];
        s1 -= col[k * 8 + 8] * yp[k + 1];
      }
      *yj = s0 + s1;
    }
    __syncwarp();
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicCAS(a.bad, 0, slot + 1);
    return;                                       // leave x untouched for the iterative fallback
  }
  for (int t = tid; t < T; t += THREADS) xout[t] = yb[(t >> 3) * 64 + (t & 7)];
}

static size_t dense_inv_smem_bytes(int T) {
  const int nb = (T + 7) / 8;
  return ((size_t)(nb * (nb + 1) / 2 + nb) * 64 + 2 * (size_t)nb * 8) * sizeof(double);
}

static const bool g_no_dense = getenv("VGM_NO_DENSE") != nullptr;

bool dense_solve_applicable(const vgm_operators* ops, int ruu_const, int T) {
  if (g_no_dense || !ops || T < 2 || T > DENSE_MAX_T) return false;
  return ruu_const ? ops->M != nullptr : (ops->DtD && ops->D && ops->DT);
}

template <int THREADS>
static int launch_inv(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  const size_t smem = dense_inv_smem_bytes(a.T);
  static size_t configured = 0;
  if (smem > configured) {
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_inv_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGM_CUDA_OK(cudaFuncSetAttribute(dense_spd_solve_inv_kernel<THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = smem;
  }
  ProfScope prof(PROF_VEC, (double)n_sys * ((double)a.T * a.T * a.T / 3.0 + 4.0 * a.T * a.T), (double)n_sys * a.T * 40.0, s);
  dense_spd_solve_inv_kernel<THREADS><<<n_sys, THREADS, smem, s>>>(a);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// VGM_DENSE_THREADS in {64, 128, 256} overrides the CTA size.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int dense_solve(const DenseSolveArgs& a, int n_sys, cudaStream_t s) {
  if (n_sys <= 0) return VGM_OK;
  VGM_REQUIRE(a.T <= DENSE_MAX_T, "dense solver: T too large");
  static const int threads_env = env_int("VGM_DENSE_THREADS", 0);
  int threads = threads_env ? threads_env : (a.T <= 32 ? 64 : (a.T <= 64 ? 128 : 256));
  if (threads < (a.T + 7) / 8 * 8) threads = 256;      // the back substitution wants one thread per row
  if (threads <= 64) return launch_inv<64>(a, n_sys, s);
  if (threads <= 128) return launch_inv<128>(a, n_sys, s);
  return launch_inv<256>(a, n_sys, s);
}

}  // namespace vgm
