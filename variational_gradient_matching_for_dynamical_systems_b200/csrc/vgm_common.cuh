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

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The solver is a long chain of dependent kernels, many of them a few microseconds long on small row counts (tail
// passes, per-GPU blocks of a multi-GPU run).  A kernel launched through launch_pdl() may be SCHEDULED while its
// predecessor in the stream still runs: its CTAs become resident as slots free up and park in pdl_wait(), which
// returns once the predecessor grid has completed and its writes are visible; pdl_trigger() lets the successor of
// THIS kernel be scheduled.  It is called LATE, when a CTA has nothing left to do: measured on B200, successors
// parked beside a long kernel (trigger at the top) cost it 7 % at K = 100 000 (37.4 vs 34.8 ms per step), while
// the late trigger still hides the launch latency of the successor.  Rules kept by every kernel launched this way: all
// threads call pdl_wait() before the first access to global memory any earlier kernel wrote or reads (kernel
// parameters, tensor maps and the constant operators are fine before it), so completion is transitive down the
// chain.  VGM_NO_PDL=1 falls back to ordinary launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------
// no-ops when the kernel was launched without the PDL attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
