// Optional per-launch CUDA-event timing of the kernel families (bench.py's roofline figures).
#pragma once
#include <cuda_runtime.h>

namespace vgm {

enum { PROF_DGEMM = 0, PROF_TCGEMM = 1, PROF_VEC = 2, PROF_ASSEMBLE = 3, PROF_IR_UPDATE = 4, PROF_IR_DIR = 5,
       PROF_FAMILIES = 6 };

// Records an event pair around the launches issued during its lifetime when profiling is enabled
// (vgm_profile_enable); a no-op otherwise.  `flops` / `bytes` are the algorithmic work of the launches.
struct ProfScope {
  ProfScope(int family, double flops, double bytes, cudaStream_t s);
  ~ProfScope();
  int family;
  cudaStream_t stream;
  cudaEvent_t stop;
  bool active;
};

bool profile_enabled();

}  // namespace vgm
