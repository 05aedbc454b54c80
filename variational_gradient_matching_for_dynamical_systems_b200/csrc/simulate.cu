// Synthetic data generation on the device: fourth-order Runge-Kutta integration of the Lorenz96
// system, f_k = (x_{k+1} - x_{k-2}) x_{k-1} - x_k + alpha (Lorenz96_ODEs.txt:1-10, indices on a ring), sampled on a
// time grid.  The reference integrates with scipy odeint (simulate_state_dynamics.py:79-92, third-party LSODA) and
// is not a parity target; this kernel replaces the ~T*sub*20 eager tensor launches of the round-1 generator by ONE
// launch per grid interval.
//
// One launch advances the whole ring by `n_sub` RK4 steps.  A CTA owns SIM_TILE consecutive states and works on them
// plus a halo of 8 * n_sub states per side in shared memory: every right-hand-side evaluation reaches 2 states to the
// left and 1 to the right, so after 4 stages x n_sub steps the owned part is still exact (the halo part is garbage
// and is never stored).  Neighbouring CTAs recompute each other's halos instead of synchronising: 1.6 % redundant
// work at n_sub = 5.
#include "vgm_common.cuh"

namespace vgm {

constexpr int SIM_TILE = 1024, SIM_THREADS = 256, SIM_MAX_SUB = 8, SIM_MAX_W = SIM_TILE + 16 * SIM_MAX_SUB;

__global__ void __launch_bounds__(SIM_THREADS)
lorenz96_rk4_kernel(const double* __restrict__ x_in, double* __restrict__ x_out, int K, double alpha, double h,
                    int n_sub, double* __restrict__ X, int ld, int col) {
  __shared__ double sx[SIM_MAX_W], st[SIM_MAX_W], sa[SIM_MAX_W];
  const int halo = 8 * n_sub, w = SIM_TILE + 2 * halo;
  const long long k0 = (long long)blockIdx.x * SIM_TILE - halo;
  for (int i = threadIdx.x; i < w; i += SIM_THREADS) {
    long long k = (k0 + i) % K;
    if (k < 0) k += K;
    const double v = x_in[k];
    sx[i] = v;
    st[i] = v;
    // the sample of this grid point is the state BEFORE the steps of this launch
    if (X && col >= 0 && i >= halo && i < halo + SIM_TILE && k0 + i < K) X[(size_t)(k0 + i) * ld + col] = v;
  }
  __syncthreads();
  const double c_in[4] = {0.5, 0.5, 1.0, 0.0};       // x + c h k feeds the next stage
  const double c_acc[4] = {1.0, 2.0, 2.0, 1.0};      // weights of k1..k4
  for (int step = 0; step < n_sub; ++step) {
    for (int stage = 0; stage < 4; ++stage) {
      double kv[(SIM_MAX_W + SIM_THREADS - 1) / SIM_THREADS];
      int q = 0;
      for (int i = threadIdx.x; i < w; i += SIM_THREADS, ++q) {
        double f = 0.0;
        if (i >= 2 && i < w - 1) f = (st[i + 1] - st[i - 2]) * st[i - 1] - st[i] + alpha;
        kv[q] = f;
      }
      __syncthreads();
      q = 0;
      for (int i = threadIdx.x; i < w; i += SIM_THREADS, ++q) {
        const double f = kv[q];
        sa[i] = (stage == 0 ? 0.0 : sa[i]) + c_acc[stage] * f;
        if (stage < 3) st[i] = sx[i] + c_in[stage] * h * f;
        else { const double v = sx[i] + (h / 6.0) * sa[i]; sx[i] = v; st[i] = v; }
      }
      __syncthreads();
    }
  }
  for (int i = halo + threadIdx.x; i < halo + SIM_TILE; i += SIM_THREADS)
    if (k0 + i < K) x_out[k0 + i] = sx[i];
}

}  // namespace vgm

// x0 [K] (device) is the state at t = 0 of the integration; `n_spinup` RK4 steps of size h are discarded, then
// X[k][i] = x_k(t_i), t_i = i * n_sub * h, i < n_grid.  `scratch` holds 2 K doubles.  K >= 4.
extern "C" int vgm_lorenz96_rk4(const double* x0, int K, double alpha, double h, int n_sub, int n_spinup, int n_grid,
                                double* X, int ld, double* scratch, vgm_stream_t stream) {
  using namespace vgm;
  VGM_REQUIRE(x0 && X && scratch && K >= 4 && n_sub >= 1 && n_sub <= SIM_MAX_SUB && n_spinup >= 0 && n_grid >= 1 &&
                  ld >= n_grid && h > 0.0, "lorenz96_rk4 arguments");
  cudaStream_t s = (cudaStream_t)stream;
  double* buf[2] = {scratch, scratch + K};
  VGM_CUDA_OK(cudaMemcpyAsync(buf[0], x0, (size_t)K * sizeof(double), cudaMemcpyDeviceToDevice, s));
  const int grid = ceil_div(K, SIM_TILE);
  int cur = 0;
  for (int done = 0; done < n_spinup;) {
    const int n = min(n_sub, n_spinup - done);
    lorenz96_rk4_kernel<<<grid, SIM_THREADS, 0, s>>>(buf[cur], buf[cur ^ 1], K, alpha, h, n, nullptr, ld, -1);
    cur ^= 1;
    done += n;
  }
  for (int i = 0; i < n_grid; ++i) {
    lorenz96_rk4_kernel<<<grid, SIM_THREADS, 0, s>>>(buf[cur], buf[cur ^ 1], K, alpha, h, n_sub, X, ld, i);
    cur ^= 1;
  }
  VGM_LAUNCH_OK();
  return VGM_OK;
}
