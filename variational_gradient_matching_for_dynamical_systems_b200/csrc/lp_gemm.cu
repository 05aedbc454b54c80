// Low-precision application of the shared preconditioner:  Z = R G^T  with G = (M + shift I)^-1.
//
// A preconditioner only has to be a fixed SPD approximation of the inverse, so G is applied in
// BF16 on the tensor cores with FP32 accumulation: 1/4 of the operand bytes and none of the FP64
// DMMA pipe, which stays free for the operator GEMM (the residual, the search directions and
// every recurrence of the CG stay in FP64, so the solution still converges to the FP64
// tolerance; the iteration count grows by ~10 %).
//   Z[n][m] = sum_k Rb[n][k] * Gb[m][k]     Rb, Gb: bf16, K-contiguous;  Z: fp64
// plus the per-row partial dots  r.z  (fp64 r from `dotw`) that the CG needs.
#include <cuda_bf16.h>

#include "vgm_common.cuh"

namespace vgm {

struct LpGemmArgs {
  const __nv_bfloat16* A;  // [rows][lda]  residuals in bf16
  const __nv_bfloat16* B;  // [M][ldb]     preconditioner in bf16
  double* C;               // [rows][ldc]
  int lda, ldb, ldc;
  int N, M, K;
  const int* rows;         // logical row -> physical row of A, C, dotw
  const int* n_rows_dev;
  const double* dotw;
  double* dot_out;
  int dot_ld;
};

constexpr int LP_BM = 128, LP_BN = 128, LP_BK = 32, LP_LDS = LP_BK + 8, LP_STAGES = 4, LP_THREADS = 256;
constexpr int LP_SMEM = LP_STAGES * (LP_BM + LP_BN) * LP_LDS * 2;

__device__ __forceinline__ void mma_bf16_16x8x16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(LP_THREADS, 1) lp_gemm_kernel(const LpGemmArgs a) {
  extern __shared__ __align__(16) unsigned char lp_smem[];
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(lp_smem);    // [STAGES][BM][LDS]
  __nv_bfloat16* Bs = As + LP_STAGES * LP_BM * LP_LDS;              // [STAGES][BN][LDS]
  const int N = a.n_rows_dev ? min(*a.n_rows_dev, a.N) : a.N;
  const int n0 = blockIdx.y * LP_BM;
  if (n0 >= N) return;
  const int m0 = blockIdx.x * LP_BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;      // 2 x 4 warps, warp tile 64 x 32
  const int g = lane >> 2, t = lane & 3;

  // copy plan: 16-byte chunks (8 bf16), 4 per tile row, 2 chunks of A and 2 of B per thread
  constexpr int CPR = LP_BK / 8;
  const __nv_bfloat16* a_src[2];
  const __nv_bfloat16* b_src[2];
  int a_dst[2], b_dst[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int id = tid + i * LP_THREADS, r = id / CPR, c = id % CPR;
    int n = min(n0 + r, N - 1);
    if (a.rows) n = a.rows[n];
    a_src[i] = a.A + (size_t)n * a.lda + 8 * c;
    a_dst[i] = r * LP_LDS + 8 * c;
    const int m = min(m0 + r, a.M - 1);
    b_src[i] = a.B + (size_t)m * a.ldb + 8 * c;
    b_dst[i] = r * LP_LDS + 8 * c;
  }
  const int kc = 8 * (tid % CPR);
  const int KT = (a.K + LP_BK - 1) / LP_BK;
  auto load_stage = [&](int kt) {
    const int st = kt % LP_STAGES, k0 = kt * LP_BK;
    int bytes = (a.K - (k0 + kc)) * 2;
    bytes = max(0, min(16, bytes));
    const int koff = (bytes > 0) ? k0 : 0;
    __nv_bfloat16* as = As + st * LP_BM * LP_LDS;
    __nv_bfloat16* bs = Bs + st * LP_BN * LP_LDS;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      cp_async16(as + a_dst[i], a_src[i] + koff, bytes);
      cp_async16(bs + b_dst[i], b_src[i] + koff, bytes);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.f;

#pragma unroll
  for (int s = 0; s < LP_STAGES - 1; ++s) {
    if (s < KT) load_stage(s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<LP_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + LP_STAGES - 1;
      if (nk < KT) load_stage(nk);
      cp_async_commit();
    }
    const __nv_bfloat16* as = As + (kt % LP_STAGES) * LP_BM * LP_LDS + (wm * 64) * LP_LDS;
    const __nv_bfloat16* bs = Bs + (kt % LP_STAGES) * LP_BN * LP_LDS + (wn * 32) * LP_LDS;
#pragma unroll
    for (int kk = 0; kk < LP_BK; kk += 16) {
      uint32_t af[4][4], bf[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16* p = as + (i * 16 + g) * LP_LDS + kk + 2 * t;
        af[i][0] = *reinterpret_cast<const uint32_t*>(p);
        af[i][1] = *reinterpret_cast<const uint32_t*>(p + 8 * LP_LDS);
        af[i][2] = *reinterpret_cast<const uint32_t*>(p + 8);
        af[i][3] = *reinterpret_cast<const uint32_t*>(p + 8 * LP_LDS + 8);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16* p = bs + (j * 8 + g) * LP_LDS + kk + 2 * t;
        bf[j][0] = *reinterpret_cast<const uint32_t*>(p);
        bf[j][1] = *reinterpret_cast<const uint32_t*>(p + 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16x8x16(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  const bool want_dot = (a.dot_out != nullptr);
  __shared__ double dot_sm[4][LP_BM];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rl = wm * 64 + i * 16 + g + 8 * h;
      const int n = n0 + rl;
      const bool row_ok = n < N;
      int nc = row_ok ? n : N - 1;
      if (a.rows) nc = a.rows[nc];
      const size_t roff = (size_t)nc * a.ldc;
      double dsum = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + wn * 32 + j * 8 + 2 * t;
        if (row_ok && m < a.M) {
          const bool two = (m + 1 < a.M);
          const double v0 = (double)acc[i][j][2 * h], v1 = (double)acc[i][j][2 * h + 1];
          const size_t o = roff + m;
          if (two) {
            *reinterpret_cast<double2*>(a.C + o) = make_double2(v0, v1);
          } else {
            a.C[o] = v0;
          }
          if (want_dot) {
            dsum += a.dotw[o] * v0;
            if (two) dsum += a.dotw[o + 1] * v1;
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
    for (int r = tid; r < LP_BM; r += LP_THREADS) {
      const int n = n0 + r;
      if (n < N) {
        const double s = dot_sm[0][r] + dot_sm[1][r] + dot_sm[2][r] + dot_sm[3][r];
        const int nc = a.rows ? a.rows[n] : n;
        a.dot_out[(size_t)nc * a.dot_ld + blockIdx.x] = s;
      }
    }
  }
}

int lp_gemm_dot_tiles(int M) { return ceil_div(M, LP_BN); }

int lp_gemm(const LpGemmArgs& a, cudaStream_t s) {
  VGM_REQUIRE(a.A && a.B && a.C && a.N >= 0 && a.M > 0 && a.K > 0, "lp_gemm arguments");
  VGM_REQUIRE((a.lda % 8) == 0 && (a.ldb % 8) == 0 && (a.ldc % 2) == 0, "lp_gemm leading dimensions");
  if (a.N == 0) return VGM_OK;
  static bool attr = false;
  if (!attr) {
    VGM_CUDA_OK(cudaFuncSetAttribute(lp_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LP_SMEM));
    attr = true;
  }
  dim3 grid(ceil_div(a.M, LP_BN), ceil_div(a.N, LP_BM));
  lp_gemm_kernel<<<grid, LP_THREADS, LP_SMEM, s>>>(a);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

__global__ void f64_to_bf16_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __double2bfloat16(in[i]);
}

}  // namespace vgm

// G (fp64, [rows][ld]) -> bf16 copy with the same shape; used once per preconditioner.
extern "C" int vgm_convert_bf16(const double* in, void* out_bf16, int rows, int ld, vgm_stream_t stream) {
  VGM_REQUIRE(in && out_bf16 && rows > 0 && ld > 0, "convert arguments");
  const size_t n = (size_t)rows * ld;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  vgm::f64_to_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, n);
  VGM_LAUNCH_OK();
  return VGM_OK;
}
