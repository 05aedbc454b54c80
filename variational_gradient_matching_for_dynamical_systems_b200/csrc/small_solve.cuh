// Single-CTA solver for tiny dependency levels (small_solve.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/vgm_b200.h"

namespace vgm {

constexpr int SMALL_SOLVE_MAX_SYSTEMS = 8;

struct SmallSolveArgs {
  const double* M;           // [T][ld] shared system matrix (FP64)
  int T, ld, band_M;         // band_M: exact half band width of M (<= 0: dense)
  const double* Lambda;      // [n_hidden][ld]
  const double* rhs;         // [n_hidden][ld]
  double* X;                 // state rows; x lives in row slot_state[slot], its content is the initial guess
  const int* slot_state;
  const int* slot_has_self;  // optional
  const int* rows;           // [n_sys] slots to solve (device)
  float* Lscratch;           // [n_sys][T][64] factor scratch
  double tol;
  int max_pass;
  int *done, *iters;         // [n_hidden]
  int* status;               // [n_sys] 0 = converged, 1 = not converged, 2 = non-positive pivot
};

size_t small_solve_scratch_bytes(int T);
bool small_solve_applicable(const vgm_operators* ops, int ruu_const, int n_rows, int T);
int small_solve(const SmallSolveArgs& args, int n_sys, cudaStream_t s);

}  // namespace vgm
