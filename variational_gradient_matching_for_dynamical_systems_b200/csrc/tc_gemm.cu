// tcgen05 / TMEM / TMA BF16 GEMM for the low-precision side of the solver (sm_100a).
//
//   acc[n][m] = sum_k A[n][k] * B[m][k]        A: [n_rows][lda] bf16 (compact rows, K-contiguous)
//                                              B: [M][ldb]      bf16 (a T x T operator, K-contiguous)
// A preconditioner only has to be a fixed SPD approximation of an inverse, so it is applied in
// BF16 with FP32 accumulation on the 5th-generation tensor cores; the FP64 DMMA pipe is left to
// the operator GEMM.  One persistent CTA per SM, warp-specialised:
//   warp 0     TMA producer   (cp.async.bulk.tensor, 128B swizzle, 4-stage mbarrier ring)
//   warp 1     MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=256, K=16; FP32 in TMEM)
//   warps 2-5  epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Two 256-column TMEM accumulators are double-buffered so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA.
#include <cuda.h>
#include <cuda_bf16.h>

#include "vgm_common.cuh"

namespace vgm {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 4, TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2, TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TcGemmArgs {
  int M, K;                 // output columns, contraction length
  int n_rows;               // host upper bound of active rows
  const int* n_rows_dev;    // optional device-side count
  const int* rows;          // compact row j -> physical row of the [.][ld] epilogue operands (null: identity)
  int ld;                   // leading dimension of the epilogue operands
  int mode;                 // epilogue selector
  double* Z;                // MODE_Z: Z[row][m] = acc
  const double* dotw;       //         dot_out[row*dot_ld + column_tile] = sum_m dotw[row][m] * acc
  double* dot_out;
  int dot_ld;
};
enum { TC_MODE_Z = 0 };

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile in shared memory, 128-byte rows, 128B swizzle (what TMA SWIZZLE_128B writes):
// 8-row atoms of 1024 bytes => stride byte offset 1024, leading byte offset unused, version 1.
__device__ __forceinline__ uint64_t make_sw128_desc(const void* smem_ptr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem_ptr) >> 4) & 0x3FFF);       // start address  [0,14)
  d |= (uint64_t)0 << 16;                                    // LBO            [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                          // SBO            [32,46)
  d |= (uint64_t)1 << 46;                                    // version = 1    [46,48)
  d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B   [61,64)
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcGemmArgs a) {
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // [STAGES][128][128B]
  unsigned char* sB = smem + TC_STAGES * TC_A_BYTES;          // [STAGES][256][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full_bar = bars;                                  // [STAGES]
  uint64_t* empty_bar = bars + TC_STAGES;                     // [STAGES]
  uint64_t* tmem_full = bars + 2 * TC_STAGES;                 // [2]
  uint64_t* tmem_empty = bars + 2 * TC_STAGES + 2;            // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int row_tiles = (n_rows + TC_BM - 1) / TC_BM;
  const int col_tiles = (a.M + TC_BN - 1) / TC_BN;
  const int n_tiles = row_tiles * col_tiles;
  const int k_blocks = (a.K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: 512 columns = two 128 x 256 FP32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rt = tile / col_tiles, ct = tile % col_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
          tma_load_2d(sA + stage * TC_A_BYTES, &mapA, &full_bar[stage], kb * TC_BK, rt * TC_BM);
          tma_load_2d(sB + stage * TC_B_BYTES, &mapB, &full_bar[stage], kb * TC_BK, ct * TC_BN);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);             // TMA landed A/B of this stage
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(sA + stage * TC_A_BYTES);
          const uint64_t bdesc = make_sw128_desc(sB + stage * TC_B_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in 16-byte units
            tc_mma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), TC_IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);                   // frees the smem slot when the MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);                       // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int rt = tile / col_tiles, ct = tile % col_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int j = rt * TC_BM + q * 32 + lane;           // compact row handled by this thread
      const bool row_ok = j < n_rows;
      const int prow = row_ok ? (a.rows ? a.rows[j] : j) : 0;
      const size_t roff = (size_t)prow * a.ld;
      const uint32_t taddr = tmem_base + acc * TC_BN + ((uint32_t)(q * 32) << 16);
      double dsum = 0.0;
#pragma unroll 1
      for (int c = 0; c < TC_BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);                          // warp-collective: executed by all lanes
        const int m0 = ct * TC_BN + c;
        if (row_ok && m0 < a.M) {
          if (m0 + 32 <= a.M) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const double z0 = (double)__uint_as_float(v[i]), z1 = (double)__uint_as_float(v[i + 1]);
              *reinterpret_cast<double2*>(a.Z + roff + m0 + i) = make_double2(z0, z1);
              if (a.dotw) {
                const double2 w = *reinterpret_cast<const double2*>(a.dotw + roff + m0 + i);
                dsum += w.x * z0 + w.y * z1;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (m0 + i < a.M) {
                const double z = (double)__uint_as_float(v[i]);
                a.Z[roff + m0 + i] = z;
                if (a.dotw) dsum += a.dotw[roff + m0 + i] * z;
              }
            }
          }
        }
      }
      if (row_ok && a.dot_out) a.dot_out[(size_t)prow * a.dot_ld + ct] = dsum;
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*cuTensorMapEncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);
static cuTensorMapEncodeTiled_t g_encode = nullptr;

static int make_bf16_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    VGM_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return VGM_EDRIVER;
    }
    g_encode = (cuTensorMapEncodeTiled_t)fn;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld);
    return VGM_EDRIVER;
  }
  return VGM_OK;
}

int tc_gemm_dot_tiles(int M) { return ceil_div(M, TC_BN); }

// A: [a_rows][lda] bf16 (compact), B: [M][ldb] bf16
int tc_gemm(const void* A, int a_rows, int lda, const void* B, int ldb, const TcGemmArgs& args, cudaStream_t s) {
  VGM_REQUIRE(A && B && args.M > 0 && args.K > 0 && args.n_rows >= 0, "tc_gemm arguments");
  VGM_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0,
              "tc_gemm operands must have 16-byte aligned rows");
  if (args.n_rows == 0) return VGM_OK;
  static int n_sm = 0;
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    VGM_CUDA_OK(cudaGetDevice(&dev));
    VGM_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    VGM_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    attr = true;
  }
  CUtensorMap mapA, mapB;
  VGM_TRY(make_bf16_map(&mapA, A, a_rows, args.K, lda, TC_BM));
  VGM_TRY(make_bf16_map(&mapB, B, args.M, args.K, ldb, TC_BN));
  const int tiles = ceil_div(args.n_rows, TC_BM) * ceil_div(args.M, TC_BN);
  const int grid = tiles < n_sm ? tiles : n_sm;
  tc_gemm_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(mapA, mapB, args);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

}  // namespace vgm

// Test/bench entry: Z[rows[j]] = A[j] . B^T with A, B bf16 (device), Z fp64; optional dots.
extern "C" int vgm_tc_gemm_bf16(const void* A_bf16, int a_rows, int lda, const void* B_bf16, int M, int K, int ldb,
                                int n_rows, const int* n_rows_dev, const int* rows, double* Z, int ldz,
                                const double* dotw, double* dot_out, int dot_ld, vgm_stream_t stream) {
  vgm::TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.K = K; a.n_rows = n_rows; a.n_rows_dev = n_rows_dev; a.rows = rows; a.ld = ldz;
  a.mode = vgm::TC_MODE_Z; a.Z = Z; a.dotw = dotw; a.dot_out = dot_out; a.dot_ld = dot_ld;
  VGM_REQUIRE(Z && (ldz % 2) == 0, "Z");
  return vgm::tc_gemm(A_bf16, a_rows, lda, B_bf16, ldb, a, (cudaStream_t)stream);
}
