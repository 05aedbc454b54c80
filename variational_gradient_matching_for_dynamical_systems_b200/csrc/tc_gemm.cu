// tcgen05 / TMEM / TMA BF16 GEMM for the low-precision side of the mixed-precision solver (sm_100a).
//
//   out[n][m] = sum_k A0[n][k] B0[m][k]  (+ A1[n][k] B0[m][k] + A0[n][k] B1[m][k])
//       A*: [n_rows][lda] bf16, compact rows, K-contiguous  (search directions / residuals of the systems)
//       B*: [M][ldb]      bf16, a T x T operator, K-contiguous (M = (cI-D)^T(cI-D) or the preconditioner G)
// With both operands split into bf16 (hi, lo) pairs the three leading products give ~2^-16 relative
// accuracy with FP32 accumulation -- enough for the correction solves of iterative refinement, whose
// residuals stay in FP64 on the DMMA pipe (sweep.cu).  The operators of the hot path are numerically
// banded (negligible entries beyond the band), so a column tile only visits the k-blocks inside
// [m0 - band, m0 + BN + band).
// One persistent CTA per SM, warp-specialised:
//   warp 0     TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier ring; one stage = the A and B
//                               k-blocks of every split part, loaded once and used by all three products)
//   warp 1     MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=128, K=16; FP32 in TMEM)
//   warps 2-5  epilogue       (tcgen05.ld 32x32b -> registers -> swizzled smem chunk -> TMA store, 2-deep ring)
// Two 128-column TMEM accumulators are double-buffered so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"
#include "tc_gemm.cuh"

namespace vgm {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_UMMA_K = 16;
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_TILE_BYTES = 128 * TC_BK * 2;                 // one 128 x 64 bf16 operand tile (A or B part)
constexpr int TC_TMEM_COLS = 2 * TC_BN;
constexpr int TC_OUT_TILE_BYTES = 32 * 32 * 4;                 // one 32-row x 32-column FP32 chunk of the output
constexpr int TC_BRES_KB = 4;                                  // k-blocks of B a CTA can keep resident
// BRES ("B resident"): a CTA is bound to one column tile, loads the k-blocks of the operator that tile needs ONCE
// and streams only the A operand; possible when the band leaves <= TC_BRES_KB k-blocks per tile.  It halves the
// L2 -> SMEM traffic, which is what bounds the streaming variant on the narrow low-precision bands.
// SLICED (split product with the operator resident): the ring holds single 16 KB A tiles -- p_hi and p_lo of a k-block
// travel through separate slots with their own barriers -- instead of (p_hi, p_lo) pairs.  Beside 128 KB of resident
// (M_hi, M_lo) tiles the pair layout had room for two stages, i.e. 64 KB in flight per SM with a refill only after
// all twelve MMAs of a k-block: the kernel ran at the latency of that ring (4.0 us per row tile against a 2.8 us HBM
// floor).  Five slots (80 KB; the TMA-store staging of the epilogue shrinks to one chunk per warp to make room) are
// refilled after the eight MMAs that read p_hi and after the four that read p_lo.  VGM_TC_NO_SLICED=1 keeps the pairs.
template <bool SPLIT, bool BRES, bool SLICED_ = false>
struct TcCfg {
  static constexpr bool SLICED = SLICED_;
  static_assert(!SLICED || (SPLIT && BRES), "the sliced ring is the split product with resident operator tiles");
  static constexpr int PARTS = SPLIT ? 2 : 1;                  // (hi, lo) or hi only, for A and for B
  static constexpr int STAGE_BYTES = SLICED ? TC_TILE_BYTES : (BRES ? 1 : 2) * PARTS * TC_TILE_BYTES;
  static constexpr int STAGES = SLICED ? 5 : BRES ? (SPLIT ? 2 : 6) : (SPLIT ? 3 : 6);
  static constexpr int RES_BYTES = BRES ? TC_BRES_KB * PARTS * TC_TILE_BYTES : 0;
  static constexpr int OUT_BUFS = SLICED ? 1 : 2;              // TMA-store staging chunks per epilogue warp
  static constexpr int OUT_STAGING = TC_EPI_WARPS * OUT_BUFS * TC_OUT_TILE_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + RES_BYTES + OUT_STAGING + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(SMEM <= 232448, "more than 227 KB of shared memory");
};
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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

struct TcMaps {
  CUtensorMap a0, a1, b0, b1, out;
};

// k-block range of column tile ct
__device__ __forceinline__ void tc_k_range(int ct, int K, int band, int& kb_lo, int& kb_hi) {
  kb_lo = 0;
  kb_hi = (K + TC_BK - 1) / TC_BK;
  if (band > 0) {
    const int m0 = ct * TC_BN;
    kb_lo = max(0, m0 - band) / TC_BK;
    kb_hi = min(kb_hi, (min(K, m0 + TC_BN + band) + TC_BK - 1) / TC_BK);
  }
}

template <bool SPLIT, bool BRES, bool SLICED = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ TcMaps maps, const TcGemmArgs a) {
  using Cfg = TcCfg<SPLIT, BRES, SLICED>;
  constexpr int STAGES = Cfg::STAGES, PARTS = Cfg::PARTS;
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage layout: [A0][A1 if split] ([B0][B1 if split] unless BRES), 16 KB each; BRES keeps [kb][B0][B1] after them
  unsigned char* b_res = smem + STAGES * Cfg::STAGE_BYTES;
  unsigned char* out_stage = b_res + Cfg::RES_BYTES;             // [EPI_WARPS][2][32 x 128 B], 128B-swizzled rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + Cfg::OUT_STAGING);
  uint64_t* full_bar = bars;                                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;                        // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;                    // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;               // [2]
  uint64_t* b_full = bars + 2 * STAGES + 4;                   // BRES: the resident operator tiles have landed
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_tiles = (a.M + TC_BN - 1) / TC_BN;
  const int ct_fixed = (int)blockIdx.x % col_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 32 * TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: two 128 x 128 FP32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // BRES: the operator tiles of this CTA's column tile are loaded once.  They are constant (M, G), so the load is
  // issued BEFORE pdl_wait(): it overlaps the tail of the preceding kernel (programmatic dependent launch).
  if (BRES && warp == 0 && lane == 0) {
    int kb_lo, kb_hi;
    tc_k_range(ct_fixed, a.K, a.band, kb_lo, kb_hi);
    mbar_expect_tx(b_full, (uint32_t)(kb_hi - kb_lo) * PARTS * TC_TILE_BYTES);
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      unsigned char* rb = b_res + (kb - kb_lo) * PARTS * TC_TILE_BYTES;
      tma_load_2d(rb, &maps.b0, b_full, kb * TC_BK, ct_fixed * TC_BN);
      if (SPLIT) tma_load_2d(rb + TC_TILE_BYTES, &maps.b1, b_full, kb * TC_BK, ct_fixed * TC_BN);
    }
  }
  pdl_wait();        // everything below reads what earlier kernels wrote (row count, A operand) or writes `out`
  const int n_rows = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int row_tiles = (n_rows + TC_BM - 1) / TC_BM;
  const int n_tiles = row_tiles * col_tiles;
  // work items of this CTA: u = u_first, u_first + u_step, ... < u_end;  streaming: u = tile index (row-tile major);
  // BRES: the column tile is fixed (blockIdx.x % col_tiles) and u runs over the row tiles
  const int u_first = BRES ? (int)blockIdx.x / col_tiles : (int)blockIdx.x;
  const int u_step = BRES ? (int)gridDim.x / col_tiles : (int)gridDim.x;
  const int u_end = BRES ? row_tiles : n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // a CTA without work still owns the operator tiles in flight: they must land before its shared memory goes away
      if (BRES && !(u_first < u_end)) mbar_wait(b_full, 0);
      for (int u = u_first; u < u_end; u += u_step) {
        const int rt = BRES ? u : u / col_tiles, ct = BRES ? ct_fixed : u % col_tiles;
        int kb_lo, kb_hi;
        tc_k_range(ct, a.K, a.band, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          if (SLICED) {                                   // one slot per part: p_hi, then p_lo
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], TC_TILE_BYTES);
              tma_load_2d(smem + stage * TC_TILE_BYTES, part ? &maps.a1 : &maps.a0, &full_bar[stage], kb * TC_BK, rt * TC_BM);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            continue;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
          tma_load_2d(st, &maps.a0, &full_bar[stage], kb * TC_BK, rt * TC_BM);
          if (SPLIT) tma_load_2d(st + TC_TILE_BYTES, &maps.a1, &full_bar[stage], kb * TC_BK, rt * TC_BM);
          if (!BRES) {
            tma_load_2d(st + PARTS * TC_TILE_BYTES, &maps.b0, &full_bar[stage], kb * TC_BK, ct * TC_BN);
            if (SPLIT) tma_load_2d(st + (PARTS + 1) * TC_TILE_BYTES, &maps.b1, &full_bar[stage], kb * TC_BK, ct * TC_BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (BRES && u_first < u_end) {
        mbar_wait(b_full, 0);
        tc_fence_after();
      }
      for (int u = u_first; u < u_end; u += u_step, ++it) {
        const int ct = BRES ? ct_fixed : u % col_tiles;
        int kb_lo, kb_hi;
        tc_k_range(ct, a.K, a.band, kb_lo, kb_hi);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        uint32_t accumulate = 0;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          if (SLICED) {
            unsigned char* bt = b_res + (kb - kb_lo) * PARTS * TC_TILE_BYTES;
            const uint64_t b0 = make_sw128_desc(bt), b1 = make_sw128_desc(bt + TC_TILE_BYTES);
            mbar_wait(&full_bar[stage], phase);           // p_hi landed: hi.hi and hi.lo
            tc_fence_after();
            const uint64_t a0 = make_sw128_desc(smem + stage * TC_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
              tc_mma_bf16(tmem_d, a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), TC_IDESC, accumulate);
              accumulate = 1;
            }
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              tc_mma_bf16(tmem_d, a0 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), TC_IDESC, 1u);
            tc_commit(&empty_bar[stage]);                 // the slot is free once these eight MMAs have retired
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            mbar_wait(&full_bar[stage], phase);           // p_lo landed: lo.hi
            tc_fence_after();
            const uint64_t a1 = make_sw128_desc(smem + stage * TC_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              tc_mma_bf16(tmem_d, a1 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), TC_IDESC, 1u);
            tc_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_wait(&full_bar[stage], phase);             // TMA landed this stage
          tc_fence_after();
          unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
          unsigned char* bt = BRES ? b_res + (kb - kb_lo) * PARTS * TC_TILE_BYTES : st + PARTS * TC_TILE_BYTES;
          const uint64_t a0 = make_sw128_desc(st);
          const uint64_t b0 = make_sw128_desc(bt);
          // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in 16-byte units
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            tc_mma_bf16(tmem_d, a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), TC_IDESC, accumulate);
            accumulate = 1;
          }
          if (SPLIT) {
            const uint64_t a1 = make_sw128_desc(st + TC_TILE_BYTES);
            const uint64_t b1 = make_sw128_desc(bt + TC_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              tc_mma_bf16(tmem_d, a1 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), TC_IDESC, 1u);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              tc_mma_bf16(tmem_d, a0 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), TC_IDESC, 1u);
          }
          tc_commit(&empty_bar[stage]);                   // frees the smem slot when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);                       // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    int it = 0;
    int epi_buf = 0;
    for (int u = u_first; u < u_end; u += u_step, ++it) {
      const int rt = BRES ? u : u / col_tiles, ct = BRES ? ct_fixed : u % col_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // partial dots: the 128 weights of this thread's row are fetched BEFORE the accumulator is waited for, so the
      // latency hides behind the MMAs of the tile.  (Staging them through shared memory by TMA instead of these
      // row-per-thread reads changed nothing: what the dots cost the GEMM, +15 us on 43 us, is the second pass of the
      // weight rows through L2.)
      const int row = rt * TC_BM + q * 32 + lane;
      const bool dot = a.out_bf16 && a.dot_out != nullptr && row < n_rows;
      uint4 wv[16];
      if (dot) {
        const __nv_bfloat16* wrow = reinterpret_cast<const __nv_bfloat16*>(a.dot_w) + (size_t)row * a.dot_ldw + ct * TC_BN;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          wv[i] = (ct * TC_BN + 8 * i < a.M) ? __ldg(reinterpret_cast<const uint4*>(wrow + 8 * i)) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // This thread's TMEM lane is row rt*128 + q*32 + lane.  A warp moves its 32 x 32 chunk through shared
      // memory (the 128B swizzle TMA expects: 16-byte piece i of row r sits at piece i ^ (r & 7), which also
      // makes the row-per-thread writes bank-conflict free) and one lane hands it to a TMA store, so global
      // memory sees whole 128-byte lines; rows >= n_rows and columns >= M are clipped by the tensor map.
      const uint32_t taddr = tmem_base + acc * TC_BN + ((uint32_t)(q * 32) << 16);
      unsigned char* my_stage = out_stage + (warp - 2) * Cfg::OUT_BUFS * TC_OUT_TILE_BYTES;
      if (a.out_bf16) {
        // bf16 output: two 32-column TMEM loads make one 64-column (128-byte) row chunk, same swizzle and TMA store
        float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};   // eight independent chains
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 64 * cc;
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr + c, v0);
          tmem_ld32(taddr + c + 32, v1);
          const int m0 = ct * TC_BN + c;
          if (m0 >= a.M) continue;                          // uniform
          if (lane == 0) bulk_wait_read<Cfg::OUT_BUFS - 1>();
          __syncwarp();
          unsigned char* sb = my_stage + epi_buf * TC_OUT_TILE_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t* src = (i < 4) ? (v0 + 8 * i) : (v1 + 8 * (i - 4));
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(src[2 * e]), __uint_as_float(src[2 * e + 1]));
              w[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            if (dot) {
              const uint4 wq = wv[8 * cc + i];
              const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                pa[e] = fmaf(__uint_as_float(w[e] << 16), __uint_as_float(ww[e] << 16), pa[e]);
                pb[e] = fmaf(__uint_as_float(w[e] & 0xffff0000u), __uint_as_float(ww[e] & 0xffff0000u), pb[e]);
              }
            }
            *reinterpret_cast<uint4*>(sb + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_2d(&maps.out, sb, m0, rt * TC_BM + q * 32);
          epi_buf = (epi_buf + 1) % Cfg::OUT_BUFS;
        }
        if (dot) a.dot_out[(size_t)row * a.dot_ld + ct] = ((pa[0] + pb[0]) + (pa[1] + pb[1])) + ((pa[2] + pb[2]) + (pa[3] + pb[3]));
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < TC_BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);                          // warp-collective: executed by all lanes
        const int m0 = ct * TC_BN + c;
        if (m0 >= a.M) continue;                          // uniform
        // the store that last used this buffer must have read it (with two buffers one newer store may be pending)
        if (lane == 0) bulk_wait_read<Cfg::OUT_BUFS - 1>();
        __syncwarp();
        unsigned char* sb = my_stage + epi_buf * TC_OUT_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(sb + lane * 128 + ((i ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tma_store_2d(&maps.out, sb, m0, rt * TC_BM + q * 32);
        epi_buf = (epi_buf + 1) % Cfg::OUT_BUFS;
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }
  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // output stores done
  tc_fence_before();
  __syncthreads();
  pdl_trigger();     // late: a successor parked beside a long persistent kernel only costs it issue slots
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*cuTensorMapEncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);
static cuTensorMapEncodeTiled_t g_encode = nullptr;

// cuTensorMapEncodeTiled costs ~1 us of host time and a GEMM needs up to five maps; the solver launches the same few
// (pointer, shape) combinations hundreds of times per sweep, so encoded maps are kept in a small direct-mapped cache.
struct MapKey {
  const void* base;
  int dtype, rows, cols, ld, box_cols, box_rows;
  bool operator==(const MapKey& o) const {
    return base == o.base && dtype == o.dtype && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows;
  }
};
constexpr int MAP_CACHE = 64;
static MapKey g_map_keys[MAP_CACHE];
static CUtensorMap g_map_vals[MAP_CACHE];
static bool g_map_used[MAP_CACHE];

static int make_map_uncached(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, int rows,
                             int cols, int ld, int box_cols, int box_rows);

static int make_map(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, int rows, int cols,
                    int ld, int box_cols, int box_rows) {
  const MapKey key{base, (int)dtype, rows, cols, ld, box_cols, box_rows};
  size_t h = (size_t)((uintptr_t)base >> 8) * 1000003u + (size_t)rows * 7919u + (size_t)cols * 31u + (size_t)box_rows;
  const int slot = (int)(h % MAP_CACHE);
  if (g_map_used[slot] && g_map_keys[slot] == key) {
    *map = g_map_vals[slot];
    return VGM_OK;
  }
  VGM_TRY(make_map_uncached(map, dtype, elem_bytes, base, rows, cols, ld, box_cols, box_rows));
  g_map_keys[slot] = key;
  g_map_vals[slot] = *map;
  g_map_used[slot] = true;
  return VGM_OK;
}

static int make_map_uncached(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, int rows,
                             int cols, int ld, int box_cols, int box_rows) {
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
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld);
    return VGM_EDRIVER;
  }
  return VGM_OK;
}
static int make_bf16_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld, TC_BK, box_rows);
}

template <bool SPLIT, bool BRES, bool SLICED = false>
static int tc_launch(const TcMaps& maps, const TcGemmArgs& args, int n_sm, cudaStream_t s) {
  using Cfg = TcCfg<SPLIT, BRES, SLICED>;
  static bool attr = false;
  if (!attr) {
    VGM_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<SPLIT, BRES, SLICED>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int row_tiles = ceil_div(args.n_rows, TC_BM), col_tiles = ceil_div(args.M, TC_BN);
  int grid;
  if (BRES) {
    const int per_col = max(1, min(row_tiles, n_sm / col_tiles));
    grid = per_col * col_tiles;                              // a multiple of col_tiles: CTA b serves column tile b % col_tiles
  } else {
    grid = min(row_tiles * col_tiles, n_sm);
  }
  VGM_CUDA_OK(launch_pdl(tc_gemm_kernel<SPLIT, BRES, SLICED>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM, s, maps, args));
  return VGM_OK;
}

// can every column tile keep its operator k-blocks resident?
static bool tc_bres_ok(const TcGemmArgs& a, int n_sm) {
  static const bool off = getenv("VGM_TC_NO_BRES") != nullptr;
  const int col_tiles = ceil_div(a.M, TC_BN);
  if (off || a.band <= 0 || col_tiles > n_sm || ceil_div(a.n_rows, TC_BM) < 2 * (n_sm / col_tiles)) return false;
  for (int ct = 0; ct < col_tiles; ++ct) {
    const int m0 = ct * TC_BN;
    const int lo = max(0, m0 - a.band) / TC_BK, hi = ceil_div(min(a.K, m0 + TC_BN + a.band), TC_BK);
    if (hi - lo > TC_BRES_KB) return false;
  }
  return true;
}

int tc_gemm(const TcGemmArgs& args, cudaStream_t s) {
  VGM_REQUIRE(args.A0 && args.B0 && args.out && args.M > 0 && args.K > 0 && args.n_rows >= 0, "tc_gemm arguments");
  VGM_REQUIRE((args.A1 != nullptr) == (args.B1 != nullptr), "split needs both A1 and B1");
  VGM_REQUIRE((args.lda % 8) == 0 && (args.ldb % 8) == 0 && (args.ldo % 4) == 0, "tc_gemm leading dimensions");
  VGM_REQUIRE(((uintptr_t)args.A0 % 16) == 0 && ((uintptr_t)args.B0 % 16) == 0 && ((uintptr_t)args.A1 % 16) == 0 &&
                  ((uintptr_t)args.B1 % 16) == 0 && ((uintptr_t)args.out % 16) == 0,
              "tc_gemm operands must be 16-byte aligned");
  if (args.n_rows == 0) return VGM_OK;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    VGM_CUDA_OK(cudaGetDevice(&dev));
    VGM_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool split = args.A1 != nullptr;
  const TcGemmArgs& largs = args;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  VGM_TRY(make_bf16_map(&maps.a0, args.A0, args.n_rows, args.K, args.lda, TC_BM));
  VGM_TRY(make_bf16_map(&maps.b0, args.B0, args.M, args.K, args.ldb, TC_BN));
  if (split) {
    VGM_TRY(make_bf16_map(&maps.a1, args.A1, args.n_rows, args.K, args.lda, TC_BM));
    VGM_TRY(make_bf16_map(&maps.b1, args.B1, args.M, args.K, args.ldb, TC_BN));
  }
  if (args.out_bf16) {
    VGM_REQUIRE((args.ldo % 8) == 0, "bf16 output rows must be 16-byte aligned");
    if (args.dot_out)
      VGM_REQUIRE(args.dot_w && (args.M % 8) == 0 && (args.dot_ldw % 8) == 0 && ((uintptr_t)args.dot_w % 16) == 0 &&
                      args.dot_ld >= (args.M + TC_BN - 1) / TC_BN,
                  "tc_gemm partial dots: M % 8 == 0, 16-byte aligned bf16 weight rows, dot_ld >= ceil(M / 128)");
  } else {
    VGM_REQUIRE(!args.dot_out, "tc_gemm partial dots need the bf16 output");
  }
  if (args.out_bf16) {
    VGM_TRY(make_map(&maps.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, args.out, args.n_rows, args.M, args.ldo, 64, 32));
  } else {
    VGM_TRY(make_map(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, args.out, args.n_rows, args.M, args.ldo, 32, 32));
  }
  // executed MMA flops (band) and compulsory HBM bytes (A parts once, FP32 output once)
  double ksum = 0.0;
  for (int m0 = 0; m0 < args.M; m0 += TC_BN)
    ksum += args.band > 0 ? (double)(min(args.K, m0 + TC_BN + args.band) - max(0, m0 - args.band)) : (double)args.K;
  const double flops = 2.0 * args.n_rows * TC_BN * ksum * (split ? 3.0 : 1.0);
  const double bytes = (double)args.n_rows * args.K * 2.0 * (split ? 2.0 : 1.0) +
                       (double)args.n_rows * args.M * (args.out_bf16 ? 2.0 : 4.0);
  ProfScope prof(PROF_TCGEMM, flops, bytes, s);
  if (tc_bres_ok(largs, n_sm)) {
    static const bool no_sliced = getenv("VGM_TC_NO_SLICED") != nullptr;
    if (split) return no_sliced ? tc_launch<true, true>(maps, largs, n_sm, s) : tc_launch<true, true, true>(maps, largs, n_sm, s);
    return tc_launch<false, true>(maps, largs, n_sm, s);
  }
  return split ? tc_launch<true, false>(maps, largs, n_sm, s) : tc_launch<false, false>(maps, largs, n_sm, s);
}

__global__ void f64_to_bf16_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __double2bfloat16(in[i]);
}
// hi = bf16(x), lo = bf16(x - hi): out[0..n) = hi plane, out[n..2n) = lo plane
__global__ void f64_split_bf16_kernel(const double* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double x = in[i];
    const __nv_bfloat16 hi = __double2bfloat16(x);
    out[i] = hi;
    out[n + i] = __double2bfloat16(x - (double)__bfloat162float(hi));
  }
}

}  // namespace vgm

extern "C" int vgm_tc_gemm_bf16(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1,
                                int M, int K, int ldb, int band, float* out, int ldo, vgm_stream_t stream) {
  vgm::TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.n_rows = n_rows; a.M = M; a.K = K; a.band = band;
  a.A0 = A0; a.A1 = A1; a.lda = lda; a.B0 = B0; a.B1 = B1; a.ldb = ldb; a.out = out; a.ldo = ldo;
  return vgm::tc_gemm(a, (cudaStream_t)stream);
}

// Same product with the accumulators rounded to bf16 on the way out (out: [n_rows][ldo] bf16).
extern "C" int vgm_tc_gemm_bf16_out(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1,
                                    int M, int K, int ldb, int band, void* out_bf16, int ldo, vgm_stream_t stream) {
  vgm::TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.n_rows = n_rows; a.M = M; a.K = K; a.band = band;
  a.A0 = A0; a.A1 = A1; a.lda = lda; a.B0 = B0; a.B1 = B1; a.ldb = ldb; a.out = (float*)out_bf16; a.ldo = ldo;
  a.out_bf16 = 1;
  return vgm::tc_gemm(a, (cudaStream_t)stream);
}

// G (fp64, [rows][ld]) -> bf16 copy with the same shape; used once per preconditioner.
extern "C" int vgm_convert_bf16(const double* in, void* out_bf16, int rows, int ld, vgm_stream_t stream) {
  VGM_REQUIRE(in && out_bf16 && rows > 0 && ld > 0, "convert arguments");
  const size_t n = (size_t)rows * ld;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  vgm::f64_to_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, n);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// M (fp64, [rows][ld]) -> [2][rows][ld] bf16 (hi, lo) planes.
extern "C" int vgm_split_bf16(const double* in, void* out_bf16x2, int rows, int ld, vgm_stream_t stream) {
  VGM_REQUIRE(in && out_bf16x2 && rows > 0 && ld > 0, "split arguments");
  const size_t n = (size_t)rows * ld;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  vgm::f64_split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16x2, n);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// vgm_tc_gemm_bf16_out plus the per-row partial dots of the rounded output with a bf16 weight matrix W [n_rows][ldw]:
// dots[j][ct] = sum_{m in column tile ct} bf16(out[j][m]) W[j][m], ct < ceil(M / 128) <= ld_dots  (M % 8 == 0).
extern "C" int vgm_tc_gemm_bf16_out_dots(const void* A0, const void* A1, int n_rows, int lda, const void* B0, const void* B1,
                                         int M, int K, int ldb, int band, void* out_bf16, int ldo, const void* W, int ldw,
                                         float* dots, int ld_dots, vgm_stream_t stream) {
  vgm::TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.n_rows = n_rows; a.M = M; a.K = K; a.band = band;
  a.A0 = A0; a.A1 = A1; a.lda = lda; a.B0 = B0; a.B1 = B1; a.ldb = ldb; a.out = (float*)out_bf16; a.ldo = ldo;
  a.out_bf16 = 1;
  a.dot_w = W; a.dot_ldw = ldw; a.dot_out = dots; a.dot_ld = ld_dots;
  return vgm::tc_gemm(a, (cudaStream_t)stream);
}
