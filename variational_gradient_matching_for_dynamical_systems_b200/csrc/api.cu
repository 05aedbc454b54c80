// Error reporting + the whole-iteration entry points of libvgm_b200.
#include <stdarg.h>

#include <vector>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"

namespace vgm {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);

// ---- per-family CUDA-event profile ------------------------------------------------------
struct ProfFamily {
  std::vector<cudaEvent_t> ev;   // pairs (start, stop), reused across collections
  size_t used = 0;
  double flops = 0.0, bytes = 0.0;
  long long launches = 0;
};
static bool g_prof_on = false;
static ProfFamily g_prof[PROF_FAMILIES];

ProfScope::ProfScope(int fam, double flops, double bytes, cudaStream_t s) : family(fam), stream(s), stop(nullptr), active(false) {
  if (!g_prof_on || fam < 0 || fam >= PROF_FAMILIES) return;
  ProfFamily& f = g_prof[fam];
  while (f.used + 2 > f.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    f.ev.push_back(e);
  }
  if (cudaEventRecord(f.ev[f.used], s) != cudaSuccess) return;
  stop = f.ev[f.used + 1];
  f.used += 2;
  f.flops += flops;
  f.bytes += bytes;
  f.launches += 1;
  active = true;
}
ProfScope::~ProfScope() {
  if (active) cudaEventRecord(stop, stream);
}

}  // namespace vgm

using namespace vgm;

extern "C" const char* vgm_last_error(void) { return g_err; }
extern "C" int vgm_version(void) { return 101; }

extern "C" int vgm_profile_enable(int on) {
  g_prof_on = (on != 0);
  for (int f = 0; f < PROF_FAMILIES; ++f) {
    g_prof[f].used = 0;
    g_prof[f].flops = g_prof[f].bytes = 0.0;
    g_prof[f].launches = 0;
  }
  return VGM_OK;
}

// Sum of the CUDA-event durations of every launch of `family` since the last enable/collect
// (synchronises the device), with the algorithmic flops / bytes of those launches and their count.
extern "C" int vgm_profile_collect(int family, double* flops, double* bytes, double* ms, long long* launches) {
  VGM_REQUIRE(family >= 0 && family < PROF_FAMILIES, "profile family");
  VGM_CUDA_OK(cudaDeviceSynchronize());
  ProfFamily& f = g_prof[family];
  double total = 0.0;
  for (size_t i = 0; i + 1 < f.used; i += 2) {
    float t = 0.f;
    VGM_CUDA_OK(cudaEventElapsedTime(&t, f.ev[i], f.ev[i + 1]));
    total += t;
  }
  if (flops) *flops = f.flops;
  if (bytes) *bytes = f.bytes;
  if (ms) *ms = total;
  if (launches) *launches = f.launches;
  f.used = 0;
  f.flops = f.bytes = 0.0;
  f.launches = 0;
  return VGM_OK;
}

// ---------------------------------------------------------------------------------------
// One coordinate-ascent iteration = theta update (proxies...py:41-139, analytical) followed by
// the state sweep (proxies...py:159-375, analytical), everything resident on the device.
// ---------------------------------------------------------------------------------------
extern "C" size_t vgm_iteration_ws_bytes(const vgm_model_tables* model, const vgm_sweep_plan* plan) {
  if (!model || !plan) return 0;
  return vgm_param_stats_ws_bytes(model->n_odes, plan->T, model->n_param) + 256 /*theta_star*/ +
         vgm_state_sweep_ws_bytes(plan) + 512;
}

extern "C" int vgm_iteration_device(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                                    const vgm_operators* ops, const vgm_solver_opts* opts, double* X, double* DX,
                                    double* theta, double* stats, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                                    vgm_stream_t stream) {
  VGM_REQUIRE(prog && model && plan && ops && X && DX && theta && stats && ws, "null argument");
  VGM_REQUIRE(ws_bytes >= vgm_iteration_ws_bytes(model, plan), "iteration workspace too small");
  VGM_REQUIRE(((uintptr_t)ws % 256) == 0, "workspace must be 256-byte aligned");
  VGM_REQUIRE(model->n_param == vgm_program_n_param(prog) && model->n_param <= 8, "parameter count");
  cudaStream_t s = (cudaStream_t)stream;
  char* p = (char*)ws;
  const size_t stats_ws = vgm_param_stats_ws_bytes(model->n_odes, plan->T, model->n_param);
  void* ws_stats = p; p += stats_ws;
  double* theta_star = (double*)p; p += 256;
  void* ws_sweep = p;
  const size_t ws_sweep_bytes = ws_bytes - (size_t)(p - (char*)ws);

  // D x_k for every state: one DMMA GEMM                                     (:76, :225)
  vgm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = X; g.B = ops->D; g.C = DX;
  g.lda = g.ldb = g.ldc = plan->ld;
  g.N = model->n_states; g.M = plan->T; g.K = plan->T;
  g.alpha = 1.0;
  g.band = ops->band_D;
  VGM_TRY(dgemm_nt(g, s));
  VGM_TRY(vgm_param_stats(prog, X, DX, plan->ld, plan->T, model->n_odes, model->ode_class, model->ode_idx,
                          model->ode_state, stats, ws_stats, stats_ws, stream));
  VGM_TRY(vgm_param_solve(stats, model->n_param, theta, stream));
  if (model->theta_mode == 0) {
    VGM_CUDA_OK(cudaMemcpyAsync(theta_star, model->theta_literal, model->n_param * sizeof(double),
                                cudaMemcpyHostToDevice, s));
  } else {
    VGM_CUDA_OK(cudaMemcpyAsync(theta_star, theta, model->n_param * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  int rc = vgm_state_sweep(prog, plan, ops, opts, X, DX, theta_star, ws_sweep, ws_sweep_bytes, info, stream);
  if (info) {
    info->gemm_calls += 1;
    info->kernel_launches += 4;
  }
  return rc;
}

extern "C" int vgm_iteration_host(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                                  const vgm_operators* ops, const vgm_solver_opts* opts, const double* X_host_in,
                                  double* X_host_out, double* theta_host, double* X_dev, double* DX_dev, double* theta_dev,
                                  double* stats_dev, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                                  vgm_stream_t stream) {
  VGM_REQUIRE(model && plan && X_host_in && X_host_out && theta_host && X_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)model->n_states * plan->ld * sizeof(double);
  VGM_CUDA_OK(cudaMemcpyAsync(X_dev, X_host_in, bytes, cudaMemcpyHostToDevice, s));
  int rc = vgm_iteration_device(prog, model, plan, ops, opts, X_dev, DX_dev, theta_dev, stats_dev, ws, ws_bytes, info,
                                stream);
  if (rc != VGM_OK && rc != VGM_ENOTCONV) return rc;
  if (model->out_n_rows > 0 && model->out_row_stride > 0 &&
      model->out_row0 + (size_t)(model->out_n_rows - 1) * model->out_row_stride < (size_t)model->n_states) {
    // only the rows the sweep updated: one strided (2-D) copy
    const size_t row_bytes = (size_t)plan->ld * sizeof(double), pitch = row_bytes * model->out_row_stride;
    const size_t off = (size_t)model->out_row0 * plan->ld;
    VGM_CUDA_OK(cudaMemcpy2DAsync(X_host_out + off, pitch, X_dev + off, pitch, row_bytes, model->out_n_rows,
                                  cudaMemcpyDeviceToHost, s));
  } else {
    VGM_CUDA_OK(cudaMemcpyAsync(X_host_out, X_dev, bytes, cudaMemcpyDeviceToHost, s));
  }
  VGM_CUDA_OK(cudaMemcpyAsync(theta_host, theta_dev, model->n_param * sizeof(double), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  return rc;
}
