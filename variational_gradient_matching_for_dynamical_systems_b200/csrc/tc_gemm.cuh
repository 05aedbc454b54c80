// Argument block of the tcgen05 BF16 GEMM (tc_gemm.cu) shared with the solver (sweep.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vgm {

// out[j][m] = sum_k A0[j][k] B0[m][k]  (+ A1 B0^T + A0 B1^T when split), raw FP32 accumulators.
struct TcGemmArgs {
  int n_rows;               // rows of A* and out (compact); host upper bound when n_rows_dev is set
  const int* n_rows_dev;    // optional device-side row count (<= n_rows): lets a solver enqueue passes without host polls
  int M, K;                 // output columns, contraction length
  int band;                 // B*[m][k] == 0 for |m - k| > band (<= 0: dense)
  const void* A0;           // [n_rows][lda] bf16
  const void* A1;           // optional low part of A (split)
  int lda;
  const void* B0;           // [M][ldb] bf16
  const void* B1;           // optional low part of B (split)
  int ldb;
  float* out;               // [n_rows][ldo] fp32 -- or, with out_bf16, [n_rows][ldo] bf16 stored at the same address
  int ldo;                  // leading dimension of out in ELEMENTS of its type
  int out_bf16;             // 1: round the accumulators to bf16 on the way out (half the output bytes)
  // optional, with out_bf16 and M % 8 == 0: per-row partial dots of the ROUNDED output with a bf16 weight row,
  // dot_out[j][ct] = sum over the 128 columns of column tile ct of bf16(out[j][m]) * dot_w[j][m]   (the epilogue thread
  // that owns row j of a tile holds those 128 accumulators anyway; the weight row is an L2 hit when it is the A operand)
  const void* dot_w;        // [n_rows][dot_ldw] bf16
  int dot_ldw;
  float* dot_out;           // [n_rows][dot_ld], dot_ld >= ceil(M / 128)
  int dot_ld;
};

int tc_gemm(const TcGemmArgs& args, cudaStream_t s);

}  // namespace vgm
