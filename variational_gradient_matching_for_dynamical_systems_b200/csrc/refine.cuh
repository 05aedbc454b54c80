// Low-precision side of the mixed-precision (iterative refinement) solver: buffers and vector kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vgm {

// All arrays are COMPACT: row j belongs to the j-th entry of the current active list.
struct IrBuf {
  float *r32, *d32, *acc32, *lam32;       // [n][ld] residual, correction, GEMM output (q' or z'), Lambda
  __nv_bfloat16 *p_hi, *p_lo;             // [n][ld] search direction p = hi + lo (the tensor-core A operand)
  __nv_bfloat16 *q_hi, *q_lo;             // [n][ld] where ir_dir writes the NEXT direction (the caller swaps p <-> q):
                                          // reading and rewriting the same bf16 lines in one kernel cost it 12 %
  __nv_bfloat16 *rsb, *s16;               // [n][ld] bf16(s o r) (A operand of the preconditioner GEMM), scaling s
  double *sigma, *rz;                     // [n] residual norm each row was scaled by; r.z
  float* rz_part;                         // [n][rz_tiles] partial dots (s o r).z' per 128-column tile, written by the
  int rz_tiles;                           // preconditioner GEMM's epilogue (tc_gemm dot_out); rz_tiles = ceil(ld / 128)
};

size_t ir_ws_bytes(int n, int ld);
void ir_carve(char*& p, int n, int ld, IrBuf& b);

// r32 = R[slot]/sigma, rsb = bf16(s r); lam32 = Lambda[slot] and s = sqrt((shift + c)/(Lambda + c)) (1 if c <= 0) are
// only rewritten when the active list changed since the previous correction (n_dev[2], set by the compaction);
// d is not initialised: the first ir_update of a correction overwrites it
// `n` is a host upper bound of the row count; `n_dev` (optional) the exact device-side count.
int ir_inner_init(const IrBuf& b, const double* R, const double* Lambda, const double* rr, const int* act, int n,
                  const int* n_dev, int T, int ld, double shift, double c, cudaStream_t s);
// z = s o acc32; rz' = (s o r).acc32 (the bf16 s o r the GEMM consumed); beta = first ? 0 : rz'/rz; p = z + beta p -> (p_hi, p_lo)
// zb16: acc32 holds the preconditioner product as bf16 rows (tc_gemm out_bf16) -- only when ir_dir_takes_bf16(T)
// zb16 == 2: rz' additionally comes from the GEMM epilogue's partial dots (b.rz_part), so the row is streamed once
bool ir_dir_takes_bf16(int T);
bool ir_dir_takes_dots();
int ir_dir(const IrBuf& b, int n, const int* n_dev, int T, int ld, int first, int zb16, cudaStream_t s);
// q = acc32 + lam o p; alpha = rz / p.q; d += alpha p (d = alpha p when `first`); r -= alpha q; rsb = bf16(s o r)
int ir_update(const IrBuf& b, int n, const int* n_dev, int T, int ld, int last, int first, double shift, double c,
              cudaStream_t s);
// x = P[slot] + sigma d; P[slot] = x; X[state(slot)] = x
int ir_finish(const IrBuf& b, double* P, double* X, const int* slot_state, const int* act, int n, const int* n_dev,
              int T, int ld, cudaStream_t s);

}  // namespace vgm
