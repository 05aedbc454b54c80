// Error reporting + the whole-iteration entry points of libvgm_b200.
#include <stdarg.h>
#include <stdlib.h>

#include <chrono>

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

bool pdl_enabled() {
  static const bool on = getenv("VGM_NO_PDL") == nullptr;
  return on;
}

static int g_device = -1;
int device_guard() {
  int dev = -1;
  VGM_CUDA_OK(cudaGetDevice(&dev));
  if (g_device < 0) g_device = dev;
  if (dev != g_device) {
    set_error("libvgm_b200 was first used on CUDA device %d and is now called on device %d: one process drives one GPU "
              "(internal streams, CUDA graphs and kernel attributes are per process)", g_device, dev);
    return VGM_EINVAL;
  }
  return VGM_OK;
}
void ir_graph_cache_clear();

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);
int param_stats_range(const vgm_program* prog, const double* X, const double* DX, int ld, int T, int j0, int j1,
                      const int* ode_class, const int* ode_idx, const int* ode_state, double* partial, cudaStream_t s);
int param_stats_reduce(const vgm_program* prog, const double* partial, int n_odes, double* stats, cudaStream_t s);
int state_sweep_prepare_range(const vgm_program* prog, const vgm_sweep_plan* pl, const vgm_operators* ops,
                              const double* X, const double* DX0, const double* theta_star, int s0, int s1, void* ws,
                              size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s);
int state_sweep_level_rows(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                           int lev, int r0, int r1, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s);
int state_sweep_row_list(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                         const int* rows_dev, int n_rows, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                         cudaStream_t s);
int state_sweep_level_full(const vgm_sweep_plan* pl, const vgm_operators* ops, const vgm_solver_opts* opts, double* X,
                           int lev, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s);

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
bool profile_enabled() { return g_prof_on; }
ProfScope::~ProfScope() {
  if (active) cudaEventRecord(stop, stream);
}

}  // namespace vgm

using namespace vgm;

extern "C" const char* vgm_last_error(void) { return g_err; }
extern "C" int vgm_version(void) { return 102; }
// Drops the cached CUDA graphs of the correction sequence (they hold raw pointers into workspaces that the caller
// may since have freed) and forgets the device binding.  Safe to call at any time; everything is rebuilt on demand.
extern "C" int vgm_shutdown(void) {
  ir_graph_cache_clear();
  g_device = -1;
  return VGM_OK;
}

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
  VGM_TRY(device_guard());
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

// ---------------------------------------------------------------------------------------
// Host-buffer iteration with the copies pipelined against the compute (vgm_model_tables.pipe_*).
//   copy-in stream :  X rows of chunk 0 (+ the cyclic wrap rows), chunk 1, chunk 2, ...
//   compute stream :  D.X(c+1), theta partials(c+1), assemble(c+1), rhs GEMM(c+1)  THEN  solve(c)
//                     -- assembling one chunk ahead keeps every coefficient on the sweep-start snapshot
//                     (proxies...py:191): solve(c) overwrites X rows that assemble(c+1) reads
//   copy-out stream:  updated rows of chunk c as soon as solve(c) is done
// ---------------------------------------------------------------------------------------
namespace vgm {
constexpr int PIPE_MAX_CHUNKS = 32;
struct PipeRes {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[PIPE_MAX_CHUNKS + 1] = {}, ev_done[PIPE_MAX_CHUNKS + 1] = {};
};
static PipeRes g_pipe;
static int pipe_init() {
  if (g_pipe.s_in) return VGM_OK;
  VGM_CUDA_OK(cudaStreamCreateWithFlags(&g_pipe.s_in, cudaStreamNonBlocking));
  VGM_CUDA_OK(cudaStreamCreateWithFlags(&g_pipe.s_out, cudaStreamNonBlocking));
  for (int i = 0; i <= PIPE_MAX_CHUNKS; ++i) {
    VGM_CUDA_OK(cudaEventCreateWithFlags(&g_pipe.ev_in[i], cudaEventDisableTiming));
    VGM_CUDA_OK(cudaEventCreateWithFlags(&g_pipe.ev_done[i], cudaEventDisableTiming));
  }
  return VGM_OK;
}

static int copy_out_rows(const vgm_model_tables* m, const vgm_sweep_plan* plan, const double* X_dev, double* X_host,
                         int k0, int k1, int s0, int s1, cudaStream_t s) {
  const size_t row_bytes = (size_t)plan->ld * sizeof(double);
  if (m->out_n_rows > 0 && m->out_row_stride > 0) {        // hidden slot i is row out_row0 + i * stride
    if (s1 <= s0) return VGM_OK;
    const size_t off = ((size_t)m->out_row0 + (size_t)s0 * m->out_row_stride) * plan->ld;
    VGM_CUDA_OK(cudaMemcpy2DAsync(X_host + off, row_bytes * m->out_row_stride, X_dev + off,
                                  row_bytes * m->out_row_stride, row_bytes, s1 - s0, cudaMemcpyDeviceToHost, s));
  } else if (k1 > k0) {
    VGM_CUDA_OK(cudaMemcpyAsync(X_host + (size_t)k0 * plan->ld, X_dev + (size_t)k0 * plan->ld, row_bytes * (k1 - k0),
                                cudaMemcpyDeviceToHost, s));
  }
  return VGM_OK;
}

static int iteration_host_pipelined(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                                    const vgm_operators* ops, const vgm_solver_opts* opts, const double* X_host_in,
                                    double* X_host_out, double* theta_host, double* X, double* DX, double* theta,
                                    double* stats, void* ws, size_t ws_bytes, vgm_sweep_info* info, cudaStream_t s) {
  const int C = model->pipe_chunks, K = model->n_states, hx = model->pipe_halo_x, hd = model->pipe_halo_dx;
  const int T = plan->T, ld = plan->ld;
  const int* kp = model->pipe_state_ptr_host;
  const int* sp = model->pipe_slot_ptr_host;
  const int* rp = model->pipe_row_ptr_host;          // SOLVE ranges (level-0 rows), shifted left by the halo
  const int* op = model->pipe_out_slot_ptr_host;     // the hidden slots of those solve ranges
  VGM_REQUIRE(C >= 2 && C <= PIPE_MAX_CHUNKS && kp && sp && rp && op && hx >= hd && hd >= 0, "pipeline description");
  VGM_REQUIRE(model->theta_mode == 0 && model->n_odes == K && kp[0] == 0 && kp[C] == K, "pipeline preconditions");
  for (int c = 0; c < C; ++c)
    VGM_REQUIRE(kp[c + 1] - kp[c] >= 2 * hx + 8 && (kp[c] % 8) == 0, "pipeline chunks too small or unaligned");
  VGM_TRY(pipe_init());
  VGM_REQUIRE(ws_bytes >= vgm_iteration_ws_bytes(model, plan) && ((uintptr_t)ws % 256) == 0, "iteration workspace");
  char* p = (char*)ws;
  const size_t stats_ws = vgm_param_stats_ws_bytes(model->n_odes, T, model->n_param);
  double* partial = (double*)p; p += stats_ws;
  double* theta_star = (double*)p; p += 256;
  void* ws_sweep = p;
  const size_t ws_sweep_bytes = ws_bytes - (size_t)(p - (char*)ws);
  const size_t row = (size_t)ld;
  if (info) memset(info, 0, sizeof(*info));

  PipeRes& R = g_pipe;
  // the copy streams start after whatever the caller queued on s
  VGM_CUDA_OK(cudaEventRecord(R.ev_done[C], s));
  VGM_CUDA_OK(cudaStreamWaitEvent(R.s_in, R.ev_done[C], 0));
  VGM_CUDA_OK(cudaStreamWaitEvent(R.s_out, R.ev_done[C], 0));
  // ---- all host->device copies, in the order the compute needs them ----
  // window c = rows [kp[c] + hx, kp[c+1] + hx) (shifted by the halo), preceded by rows [0, hx) and the wrap rows
  VGM_CUDA_OK(cudaMemcpyAsync(X, X_host_in, row * hx * sizeof(double), cudaMemcpyHostToDevice, R.s_in));
  VGM_CUDA_OK(cudaMemcpyAsync(X + row * (K - hx), X_host_in + row * (K - hx), row * hx * sizeof(double),
                              cudaMemcpyHostToDevice, R.s_in));
  for (int c = 0; c < C; ++c) {
    const int a = min(K, kp[c] + hx), b = (c == C - 1) ? K : min(K, kp[c + 1] + hx);
    if (b > a)
      VGM_CUDA_OK(cudaMemcpyAsync(X + row * a, X_host_in + row * a, row * (b - a) * sizeof(double),
                                  cudaMemcpyHostToDevice, R.s_in));
    VGM_CUDA_OK(cudaEventRecord(R.ev_in[c], R.s_in));
  }
  VGM_CUDA_OK(cudaMemcpyAsync(theta_star, model->theta_literal, model->n_param * sizeof(double), cudaMemcpyHostToDevice, s));

  auto dx_rows = [&](int a, int b) -> int {               // D.x_k for rows [a, b)
    if (b <= a) return VGM_OK;
    vgm_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = X + row * a; g.B = ops->D; g.C = DX + row * a;
    g.lda = g.ldb = g.ldc = ld;
    g.N = b - a; g.M = T; g.K = T;
    g.alpha = 1.0;
    g.band = ops->band_D;
    if (info) { info->gemm_calls += 1; info->kernel_launches += 1; }
    return dgemm_nt(g, s);
  };
  // stage(c): everything of chunk c that has to read the sweep-start snapshot
  auto stage = [&](int c) -> int {
    VGM_CUDA_OK(cudaStreamWaitEvent(s, R.ev_in[c], 0));
    if (c == 0) VGM_TRY(dx_rows(K - hd, K));             // wrap rows read by the first chunk
    const int a = (c == 0) ? 0 : min(K, kp[c] + hd), b = (c == C - 1) ? K - hd : min(K, kp[c + 1] + hd);
    VGM_TRY(dx_rows(a, b));
    // theta statistics of the ODE rows whose inputs are complete: rows [kp[c], kp[c+1]) need X up to +hx and D.X of
    // their own state (both resident now)
    VGM_TRY(param_stats_range(prog, X, DX, ld, T, kp[c], kp[c + 1], model->ode_class, model->ode_idx, model->ode_state,
                              partial, s));
    VGM_TRY(state_sweep_prepare_range(prog, plan, ops, X, DX, theta_star, sp[c], sp[c + 1], ws_sweep, ws_sweep_bytes,
                                      info, s));
    if (info) info->kernel_launches += 1;
    return VGM_OK;
  };

  // Solve range c = the level-0 systems whose state lies in [kp[c] - hx, kp[c+1] - hx): shifted left by the halo, so
  // that (i) it only needs coefficients staged up to chunk c and (ii) everything it overwrites is below what
  // stage(c+1) still has to read from the sweep-start snapshot.  The "seam" (states < hx, read cyclically by the
  // LAST stage) is solved at the very end; the last range runs to the end of the level.
  const int seam_rows = rp[0], seam_slots = op[0];
  int rc = VGM_OK;
  static const bool dbg = getenv("VGM_DEBUG") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  // with pipe_last_rows (the last range's systems followed by the seam's) the seam joins the last solve
  const bool seam_joined = model->pipe_last_rows != nullptr && seam_rows > 0;
  for (int c = 0; c < C; ++c) {
    VGM_TRY(stage(c));
    int r = (c == C - 1 && seam_joined)
                ? state_sweep_row_list(plan, ops, opts, X, model->pipe_last_rows, rp[C] - rp[C - 1] + seam_rows, ws_sweep,
                                       ws_sweep_bytes, info, s)
                : state_sweep_level_rows(plan, ops, opts, X, 0, rp[c], rp[c + 1], ws_sweep, ws_sweep_bytes, info, s);
    if (dbg) fprintf(stderr, "[vgm] pipeline: range %d (%d systems) solved at %.2f ms\n", c, rp[c + 1] - rp[c], ms_now());
    if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
    if (c + 1 < C) {                                      // the last range goes out with the seam and the later levels
      VGM_CUDA_OK(cudaEventRecord(R.ev_done[c], s));
      VGM_CUDA_OK(cudaStreamWaitEvent(R.s_out, R.ev_done[c], 0));
      VGM_TRY(copy_out_rows(model, plan, X, X_host_out, c == 0 ? min(hx, kp[1]) : kp[c] - hx, kp[c + 1] - hx, op[c],
                            op[c + 1], R.s_out));
    }
  }
  if (!seam_joined) {
    int r = state_sweep_level_rows(plan, ops, opts, X, 0, 0, seam_rows, ws_sweep, ws_sweep_bytes, info, s);
    if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
  }
  VGM_TRY(param_stats_reduce(prog, partial, model->n_odes, stats, s));
  VGM_TRY(vgm_param_solve(stats, model->n_param, theta, (vgm_stream_t)s));
  for (int lev = 1; lev < plan->n_levels; ++lev) {
    int r = state_sweep_level_full(plan, ops, opts, X, lev, ws_sweep, ws_sweep_bytes, info, s);
    if (r == VGM_ENOTCONV) rc = r; else if (r != VGM_OK) return r;
  }
  VGM_CUDA_OK(cudaEventRecord(R.ev_done[C - 1], s));
  VGM_CUDA_OK(cudaStreamWaitEvent(R.s_out, R.ev_done[C - 1], 0));
  VGM_TRY(copy_out_rows(model, plan, X, X_host_out, kp[C - 1] - hx, kp[C], op[C - 1], op[C], R.s_out));
  VGM_TRY(copy_out_rows(model, plan, X, X_host_out, 0, min(hx, kp[1]), 0, seam_slots, R.s_out));
  VGM_CUDA_OK(cudaMemcpyAsync(theta_host, theta, model->n_param * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (dbg) fprintf(stderr, "[vgm] pipeline: everything enqueued at %.2f ms\n", ms_now());
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  if (dbg) fprintf(stderr, "[vgm] pipeline: compute stream idle at %.2f ms\n", ms_now());
  VGM_CUDA_OK(cudaStreamSynchronize(R.s_out));
  if (dbg) fprintf(stderr, "[vgm] pipeline: copy-out stream idle at %.2f ms\n", ms_now());
  if (info) info->kernel_launches += 3;
  return rc;
}
}  // namespace vgm

extern "C" int vgm_iteration_host(const vgm_program* prog, const vgm_model_tables* model, const vgm_sweep_plan* plan,
                                  const vgm_operators* ops, const vgm_solver_opts* opts, const double* X_host_in,
                                  double* X_host_out, double* theta_host, double* X_dev, double* DX_dev, double* theta_dev,
                                  double* stats_dev, void* ws, size_t ws_bytes, vgm_sweep_info* info,
                                  vgm_stream_t stream) {
  VGM_REQUIRE(model && plan && X_host_in && X_host_out && theta_host && X_dev, "null argument");
  VGM_TRY(device_guard());
  cudaStream_t s = (cudaStream_t)stream;
  if (model->pipe_chunks >= 2 && plan->n_hidden > 0) {
    VGM_REQUIRE(prog && ops && DX_dev && theta_dev && stats_dev && ws, "null argument");
    VGM_REQUIRE(model->n_param == vgm_program_n_param(prog) && model->n_param <= 8, "parameter count");
    return iteration_host_pipelined(prog, model, plan, ops, opts, X_host_in, X_host_out, theta_host, X_dev, DX_dev,
                                    theta_dev, stats_dev, ws, ws_bytes, info, s);
  }
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
