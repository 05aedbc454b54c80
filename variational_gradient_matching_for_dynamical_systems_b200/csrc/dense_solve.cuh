// Direct per-system solver for short time grids (dense_solve.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/vgm_b200.h"

namespace vgm {

constexpr int DENSE_MAX_T = 160;   // lower block triangle + right-hand side block row in shared memory (119 KB at 160)

struct DenseSolveArgs {
  const double* M;           // [T][ld] shared (c I - D)^T (c I - D) when R_uu is the constant c, else null and:
  const double* D;           // [T][ld]
  const double* DT;          // [T][ld]
  const double* DtD;         // [T][ld]  D^T D
  const double* Ruu;         // [n_hidden][ld] R_uu(t) per slot
  int T, ld;
  const double* Lambda;      // [n_hidden][ld]
  const double* rhs;         // [n_hidden][ld]
  double* X;                 // state rows, row slot_state[slot] receives the solution
  const int* slot_state;
  const int* slot_has_self;  // optional; 0 => system matrix is 2 diag(Lambda) (proxies...py:201,:245)
  const int* rows;           // [n_sys] slots to solve (device)
  int* bad;                  // device flag: 1 + slot of a system whose factorisation met a non-positive pivot
};

bool dense_solve_applicable(const vgm_operators* ops, int ruu_const, int T);
int dense_solve(const DenseSolveArgs& args, int n_sys, cudaStream_t s);

}  // namespace vgm
