// Shared helpers for libvgm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vgm_b200.h"

namespace vgm {

void set_error(const char* fmt, ...);
// The library keeps a few process-wide resources (internal streams, events, cached CUDA graphs, kernel attributes)
// that belong to ONE device: one process drives one GPU.  Returns VGM_EINVAL when the calling thread's current device
// is not the one the library was first used on.
int device_guard();

#define VGM_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      vgm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VGM_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

#define VGM_LAUNCH_OK()                                                                \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      vgm::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return VGM_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

#define VGM_REQUIRE(cond, msg)                                                         \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      vgm::set_error("%s:%d: invalid argument: %s (%s)", __FILE__, __LINE__, msg, #cond); \
      return VGM_EINVAL;                                                               \
    }                                                                                  \
  } while (0)

#define VGM_TRY(expr)                                                                  \
  do {                                                                                 \
    int _r = (expr);                                                                   \
    if (_r != VGM_OK) return _r;                                                       \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread.  `red` must hold >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = (lane < nw) ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA (DMMA), D = A(16x8,row) * B(8x8,col) + C
__device__ __forceinline__ void dmma_16x8x8(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// D = A(8x4,row) * B(4x8,col) + C
__device__ __forceinline__ void dmma_8x8x4(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

}  // namespace vgm
