// FP64 tensor-core GEMM for the gradient-matching operators (sm_100a, DMMA mma.sync m16n8k8).
//
//   C[n][m] = alpha * sum_k A[n][k] * B[m][k]  (+ beta*C0 + g1*U1.V1 + g2*U2.V2)   "NT": both operands K-contiguous
//
// This one form covers every dense contraction on the reference hot path because a state's
// trajectory is contiguous in HBM (X is [K_states][T]):
//   D.x_k for all k      proxies_for_ode_parameters_and_states.py:76,225   A = X,      B = D
//   (diag(R)-D)^T r      proxies...py:231                                  A = r_uu,   B = D^T
//   B^T B p  (CG)        proxies...py:232,262 (solve)                      A = P,      B = M or D, D^T
//   dC C^-1, D dC^T      GP_regression.py:306,317                          A = dC / D, B = C^-1 / dC
//
// tcgen05 has no f64 kind, so FP64 tensor work on sm_100a is warp-level DMMA.  The B operand
// (a T x T operator, 8 MB at T=1000) stays L2 resident; the A operand streams from HBM once
// per column-tile sweep.  Operands are staged through a cp.async ring in shared memory (2 to 4
// stages, per tile configuration) with a +4 double row pad (conflict-free 64-bit fragment loads).
// Banded operators: a CTA loops over the k range its columns need and every 8-column MMA tile
// issues only the DMMAs inside its own band (the "ragged band" below).
#include <stdlib.h>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"

namespace vgm {

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int BK_ = 16, int STAGES_ = 4, int MINB_ = 1>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, LDS = BK + 4, STAGES = STAGES_, MINB = MINB_;
  static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;  // warp tile
  static constexpr int MT = WM / 16, NT = WN / 8;             // mma tiles per warp
  static constexpr int SMEM = STAGES * (BM + BN) * LDS * (int)sizeof(double);
  static_assert(WM % 16 == 0 && WN % 8 == 0, "warp tile must be a multiple of the MMA tile");
};

// VGM_NO_RAGGED_BAND=1 (experiments): every MMA tile runs the k range of its whole CTA tile, as before round 2's end
__device__ int g_no_ragged_dev = 0;

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) dgemm_nt_kernel(const vgm_gemm_args a) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDS = Cfg::LDS, STAGES = Cfg::STAGES;
  constexpr int MT = Cfg::MT, NT = Cfg::NT, THREADS = Cfg::THREADS;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                        // [STAGES][BM][LDS]
  double* Bs = smem + STAGES * BM * LDS;    // [STAGES][BN][LDS]

  pdl_wait();       // launched with programmatic stream serialization (vgm_common.cuh): the predecessor is done from here on
  const int N = a.n_rows_dev ? min(*a.n_rows_dev, a.N) : a.N;
  const int n0 = blockIdx.y * BM;
  if (n0 >= N) return;
  const int m0 = blockIdx.x * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;

  // ---- global -> shared copy plan: 16-byte chunks, BK/2 per tile row --------------
  // Thread tid copies chunk (row r0 + i*RPT, columns c2, c2+1) of the A tile for i < A_CHUNKS and of the B tile for
  // i < B_CHUNKS.  Only the (gathered) A row indices are kept, as 32-bit integers: the kernel lives on a 64-register
  // budget (4 CTAs per SM) and pointers would cost twice as much.
  constexpr int CPR = BK / 2;                       // chunks per row
  constexpr int RPT = THREADS / CPR;                // tile rows covered by one pass of the CTA
  constexpr int A_CHUNKS = BM * CPR / THREADS;      // per thread
  constexpr int B_CHUNKS = BN * CPR / THREADS;
  static_assert((BM * CPR) % THREADS == 0 && (BN * CPR) % THREADS == 0 && THREADS % CPR == 0, "tile/threads mismatch");
  const int r0 = tid / CPR;
  const int kc_of_thread = 2 * (tid % CPR);         // same column offset for every chunk of this thread
  int a_row[A_CHUNKS];
#pragma unroll
  for (int i = 0; i < A_CHUNKS; ++i) {
    int n = min(n0 + r0 + i * RPT, N - 1);
    if (a.rowsA) n = a.rowsA[n];
    a_row[i] = n;
  }
  // Banded operator: B[m][k] is (numerically) zero for |m - k| > band, so this column tile only needs
  // k in [m0 - band, m0 + BN + band).  band <= 0: dense.
  int kt_lo = 0, KT = (a.K + BK - 1) / BK;
  if (a.band > 0) {
    kt_lo = max(0, m0 - a.band) / BK;
    KT = min(KT, (min(a.K, m0 + BN + a.band) + BK - 1) / BK);
  }

  // Ragged band: the k range above is the union over the BN columns of the tile, but the 8 columns of MMA tile j
  // of this warp only need k in [mj - band, mj + 8 + band), mj = m0 + wn*WN + 8j.  In units of 8-wide k steps tile j
  // is active for lo0 + j <= step < hi0 + j (mj advances one step per tile), so at the two ends of the k loop a
  // warp issues only the DMMAs whose operator entries are not negligible: 8 + 2*band instead of BN + 2*band
  // contraction steps per output column (260 instead of 316 at band 126: the DMMA pipe is what bounds this kernel).
  int lo0 = -(1 << 28), hi0 = 1 << 28;
  if (a.band > 0 && !g_no_ragged_dev) {
    const int mw = m0 + wn * Cfg::WN;
    lo0 = (mw - a.band) >> 3;                      // arithmetic shift = floor division (mw - band may be negative)
    hi0 = (mw + 8 + a.band + 7) >> 3;
  }

  auto load_stage = [&](int kt) {
    const int st = kt % STAGES;
    const int k0 = kt * BK;
    int bytes = (a.K - (k0 + kc_of_thread)) * 8;
    bytes = max(0, min(16, bytes));
    const int koff = ((bytes > 0) ? k0 : 0) + kc_of_thread;   // keep the address in bounds when fully masked
    double* as = As + st * BM * LDS + r0 * LDS + kc_of_thread;
    double* bs = Bs + st * BN * LDS + r0 * LDS + kc_of_thread;
#pragma unroll
    for (int i = 0; i < A_CHUNKS; ++i) cp_async16(as + i * RPT * LDS, a.A + (size_t)a_row[i] * a.lda + koff, bytes);
#pragma unroll
    for (int i = 0; i < B_CHUNKS; ++i)
      cp_async16(bs + i * RPT * LDS, a.B + (size_t)min(m0 + r0 + i * RPT, a.M - 1) * a.ldb + koff, bytes);
  };

  double acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (kt_lo + s < KT) load_stage(kt_lo + s);
    cp_async_commit();
  }

  for (int kt = kt_lo; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt+STAGES-1 into the slot freed by iteration kt-1
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * BM * LDS + (wm * Cfg::WM) * LDS;
    const double* bs = Bs + (kt % STAGES) * BN * LDS + (wn * Cfg::WN) * LDS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      const int step = kt * (BK / 8) + kk / 8;
      const int j_hi = step - lo0, j_lo = step - hi0;     // MMA tile j is inside its band iff j_lo < j <= j_hi
      if (j_hi < 0 || j_lo >= NT - 1) continue;           // warp-uniform: nothing of this warp's columns at this step
      double af[MT][4], bf[NT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const double* p = as + (i * 16 + g) * LDS + kk + t;
        af[i][0] = p[0];
        af[i][1] = p[8 * LDS];
        af[i][2] = p[4];
        af[i][3] = p[8 * LDS + 4];
      }
      if (j_hi >= NT - 1 && j_lo < 0) {                   // interior of the band: every tile
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double* p = bs + (j * 8 + g) * LDS + kk + t;
          bf[j][0] = p[0];
          bf[j][1] = p[4];
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma_16x8x8(acc[i][j], af[i], bf[j]);
      } else {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (j > j_lo && j <= j_hi) {
            const double* p = bs + (j * 8 + g) * LDS + kk + t;
            bf[j][0] = p[0];
            bf[j][1] = p[4];
#pragma unroll
            for (int i = 0; i < MT; ++i) dmma_16x8x8(acc[i][j], af[i], bf[j]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ---------------------------------------------------------------------
  // thread owns rows {g, g+8} of each 16-row tile and column pairs (2t, 2t+1) of each 8-col tile
  const bool want_dot = (a.dot_out != nullptr);
  const bool dot_self = (a.dotw == a.C);
  __shared__ double dot_sm[Cfg::WARPS_N][BM];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rl = wm * Cfg::WM + i * 16 + g + 8 * h;   // row within the CTA tile
      const int n = n0 + rl;
      const bool row_ok = n < N;
      int nc = row_ok ? n : N - 1;
      if (a.rowsC) nc = a.rowsC[nc];
      const size_t roff = (size_t)nc * a.ldc;
      double dsum = 0.0;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int m = m0 + wn * Cfg::WN + j * 8 + 2 * t;
        if (row_ok && m < a.M) {
          const bool two = (m + 1 < a.M);
          double v0 = a.alpha * acc[i][j][2 * h], v1 = a.alpha * acc[i][j][2 * h + 1];
          const size_t o = roff + m;
          if (a.C0) {
            v0 += a.beta * a.C0[o];
            if (two) v1 += a.beta * a.C0[o + 1];
          }
          if (a.V1) {
            v0 += a.g1 * (a.U1 ? a.U1[o] : 1.0) * a.V1[o];
            if (two) v1 += a.g1 * (a.U1 ? a.U1[o + 1] : 1.0) * a.V1[o + 1];
          }
          if (a.V2) {
            v0 += a.g2 * (a.U2 ? a.U2[o] : 1.0) * a.V2[o];
            if (two) v1 += a.g2 * (a.U2 ? a.U2[o + 1] : 1.0) * a.V2[o + 1];
          }
          if (two) {
            *reinterpret_cast<double2*>(a.C + o) = make_double2(v0, v1);
          } else {
            a.C[o] = v0;
          }
          if (want_dot) {
            if (dot_self) {          // dotw == C: squared row norm of the result
              dsum += v0 * v0;
              if (two) dsum += v1 * v1;
            } else {
              dsum += a.dotw[o] * v0;
              if (two) dsum += a.dotw[o + 1] * v1;
            }
          }
        }
      }
      if (want_dot) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        if (t == 0) dot_sm[wn][rl] = dsum;
      }
    }
  }
  if (want_dot) {
    __syncthreads();
    for (int r = tid; r < BM; r += THREADS) {
      const int n = n0 + r;
      if (n < N) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS_N; ++w) s += dot_sm[w][r];
        const int nc = a.rowsC ? a.rowsC[n] : n;
        a.dot_out[(size_t)nc * a.dot_ld + blockIdx.x] = s;
      }
    }
  }
  pdl_trigger();    // late: the successor's CTAs are only parked for the moment the last tiles drain
}

static bool no_ragged_band();

template <class Cfg>
static int launch_gemm(const vgm_gemm_args& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    VGM_CUDA_OK(cudaFuncSetAttribute(dgemm_nt_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int off = no_ragged_band() ? 1 : 0;
    VGM_CUDA_OK(cudaMemcpyToSymbol(g_no_ragged_dev, &off, sizeof(int)));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.M, Cfg::BN), ceil_div(a.N, Cfg::BM));
  VGM_CUDA_OK(launch_pdl(dgemm_nt_kernel<Cfg>, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, a));
  return VGM_OK;
}

// Tile configurations of the large-N GEMM (vgm_dgemm_set_variant / VGM_DGEMM_VARIANT).  Measured on the banded
// residual GEMM of the bench (N = 50 000, M = K = 1000, band 146, B200): 128x128 x1 CTA/SM 15.8 TFLOP/s, 128x64 x2
// 22.4, 64x64 (128 threads) x3 23.4, x4 26.0, 32x64 x7 27.7, 64x64 (256 threads, warp tile 16x32) x4 27.6 -- the
// short banded k loop leaves a large prologue/epilogue share, which more resident CTAs overlap.
using CfgBig = GemmCfg<128, 128, 2, 4>;            // variant 0: 256 threads, warp tile 64x32, 1 CTA/SM
using CfgV2 = GemmCfg<128, 64, 4, 2, 16, 3, 2>;    // variant 2: 256 threads, warp tile 32x32, 2 CTAs/SM
using CfgV12 = GemmCfg<64, 64, 4, 2, 16, 2, 4>;    // variant 12: 256 threads, warp tile 16x32, 4 CTAs/SM
// Narrow column tiles for banded operators: all warps of a CTA work on the SAME 32 columns, so with the ragged band
// they skip the same k steps (with 64-column tiles the two warp columns idle at opposite ends of the k loop) and
// the CTA's own k range is BN + 2 band = 284 instead of 316 at band 126.
using CfgV13 = GemmCfg<128, 32, 8, 1, 16, 2, 4>;   // variant 13: 256 threads, warp tile 16x32, 4 CTAs/SM
using CfgV14 = GemmCfg<64, 32, 4, 1, 16, 2, 8>;    // variant 14 (default): 128 threads, warp tile 16x32, 8 CTAs/SM
static int g_variant = 14;
using CfgMid = GemmCfg<64, 64, 2, 2>;     // 128 threads, warp tile 32x32
using CfgThin = GemmCfg<16, 64, 1, 4>;    // 128 threads, warp tile 16x16 (few rows: halo / tail solves)

int gemm_dot_tiles(int M, int N) {
  // number of column tiles (= partial dot products per row) the chosen configuration produces
  if (N <= 16) return ceil_div(M, CfgThin::BN);
  if ((long long)ceil_div(N, 128) * ceil_div(M, 128) < 120) return ceil_div(M, CfgMid::BN);
  return ceil_div(M, g_variant == 0 ? CfgBig::BN : (g_variant == 2 || g_variant == 12) ? 64 : 32);
}

static int dgemm_dispatch(const vgm_gemm_args& a, cudaStream_t s);

// Executed contraction length summed over column tiles of width bn (the band skips the rest).
static double banded_k_sum(int M, int K, int band, int bn) {
  if (band <= 0) return (double)ceil_div(M, bn) * K;
  double sum = 0.0;
  for (int m0 = 0; m0 < M; m0 += bn) sum += (double)(min(K, m0 + bn + band) - max(0, m0 - band));
  return sum;
}

// The same with the ragged band of the kernel: every 8-column MMA tile runs its own range, in 8-wide k steps.
static double ragged_k_sum(int M, int K, int band) {
  if (band <= 0) return (double)ceil_div(M, 8) * K;
  double sum = 0.0;
  const int k_end = ceil_div(K, 8) * 8;
  for (int m0 = 0; m0 < M; m0 += 8) {
    const int lo = max(0, ((m0 - band) >> 3) * 8), hi = min(k_end, ((m0 + 8 + band + 7) >> 3) * 8);
    sum += (double)(hi - lo);
  }
  return sum;
}

static bool no_ragged_band() {
  static const bool off = getenv("VGM_NO_RAGGED_BAND") != nullptr;
  return off;
}

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s) {
  // executed flops (a banded operator skips the k range that is numerically zero) and compulsory bytes
  const int bn = 64;
  const double flops = no_ragged_band() ? 2.0 * a.N * bn * banded_k_sum(a.M, a.K, a.band, bn)
                                        : 2.0 * a.N * 8 * ragged_k_sum(a.M, a.K, a.band);
  const double bytes = 8.0 * ((double)a.N * a.K + (double)a.N * a.M);
  ProfScope prof(PROF_DGEMM, flops, bytes, s);
  return dgemm_dispatch(a, s);
}

static int dgemm_dispatch(const vgm_gemm_args& a, cudaStream_t s) {
  VGM_REQUIRE(a.A && a.B && a.C, "null operand");
  VGM_REQUIRE(a.N >= 0 && a.M > 0 && a.K > 0, "bad shape");
  VGM_REQUIRE((a.lda % 2) == 0 && (a.ldb % 2) == 0 && (a.ldc % 2) == 0, "leading dimensions must be even (16-byte rows)");
  VGM_REQUIRE(((uintptr_t)a.A % 16) == 0 && ((uintptr_t)a.B % 16) == 0 && ((uintptr_t)a.C % 16) == 0, "operands must be 16-byte aligned");
  if (a.N == 0) return VGM_OK;
  if (a.N <= 16) return launch_gemm<CfgThin>(a, s);
  if ((long long)ceil_div(a.N, 128) * ceil_div(a.M, 128) < 120) return launch_gemm<CfgMid>(a, s);
  switch (g_variant) {
    case 0: return launch_gemm<CfgBig>(a, s);
    case 2: return launch_gemm<CfgV2>(a, s);
    case 13: return launch_gemm<CfgV13>(a, s);
    case 14: return launch_gemm<CfgV14>(a, s);
    case 12: return launch_gemm<CfgV12>(a, s);
    default: return launch_gemm<CfgV14>(a, s);
  }
}

}  // namespace vgm

extern "C" int vgm_dgemm_nt(const vgm_gemm_args* args, vgm_stream_t stream) {
  if (!args) {
    vgm::set_error("vgm_dgemm_nt: null args");
    return VGM_EINVAL;
  }
  return vgm::dgemm_nt(*args, (cudaStream_t)stream);
}

// Tuning knob (tile configuration of the large-N GEMM); 14 is the default.
extern "C" int vgm_dgemm_set_variant(int v) {
  vgm::g_variant = v;
  return VGM_OK;
}

extern "C" int vgm_dgemm_dot_tiles(int M, int N) { return vgm::gemm_dot_tiles(M, N); }
