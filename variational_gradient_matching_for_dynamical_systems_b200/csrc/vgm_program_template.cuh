// Model-specific coefficient kernels ("program").  This file is compiled twice:
//   * by nvcc into libvgm_b200.so with the hand-written Lorenz96 coefficient functions, and
//   * by NVRTC at run time, prefixed with device functions that the Python code generator
//     emitted from the SymPy rewrite of an arbitrary *_ODEs.txt model
//     (rewrite_odes_as_local_linear_combinations.py:28-58, :76-124).
// It must therefore stay self-contained (no #include) and use only CUDA built-ins.
//
// Expected from the includer:
//   #define VGM_PROG(name)  <prefix>##name        kernel name prefix
//   #define VGM_NPARAM <p>                        number of ODE parameters
//   #define VGM_ARITY  <a>                        max #states referenced by one ODE
//   __device__ void vgm_theta_coeffs(int cls, const double (&s)[VGM_ARITY], double (&B)[VGM_NPARAM], double& b);
//   __device__ void vgm_state_coeffs(int code, const double* th, const double (&s)[VGM_ARITY], double& R, double& r);
//   __device__ void vgm_ode_eval(int cls, const double* th, const double (&s)[VGM_ARITY], double& f, double (&df)[VGM_NPARAM]);
//
// Layout: X, DX, Lambda, rhs, ... are [rows][ld] FP64 with the time index contiguous, so a
// warp reads/writes 256 contiguous bytes per row (coalesced, 8-byte lanes).

#define VGM_NSTAT (VGM_NPARAM + VGM_NPARAM * (VGM_NPARAM + 1) / 2)
#define VGM_STAT_ROWS 8      // ODE rows per CTA in the statistics kernel
#define VGM_PROG_THREADS 256   // max threads per CTA (launch may use fewer, multiple of 32, >= 64)

__device__ __forceinline__ double vgm_prog_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// P2: per-CTA partial sufficient statistics (proxies_for_ode_parameters_and_states.py:64-79)
//   partial[cta][0..p)            = sum B_i (D x_k - b)
//   partial[cta][p + tri(i,l)]    = sum B_i B_l   (i <= l)
extern "C" __global__ void __launch_bounds__(VGM_PROG_THREADS)
VGM_PROG(param_stats)(const double* __restrict__ X, const double* __restrict__ DX, int ld, int T, int n_odes,
                      const int* __restrict__ ode_class, const int* __restrict__ ode_idx,
                      const int* __restrict__ ode_state, double* __restrict__ partial) {
  double acc[VGM_NSTAT];
#pragma unroll
  for (int i = 0; i < VGM_NSTAT; ++i) acc[i] = 0.0;
  const int j0 = blockIdx.x * VGM_STAT_ROWS;
  for (int jj = 0; jj < VGM_STAT_ROWS; ++jj) {
    const int j = j0 + jj;
    if (j >= n_odes) break;
    const int cls = ode_class[j];
    const double* dxrow = DX + (size_t)ode_state[j] * ld;
    const double* xr[VGM_ARITY];
#pragma unroll
    for (int a = 0; a < VGM_ARITY; ++a) xr[a] = X + (size_t)ode_idx[(size_t)j * VGM_ARITY + a] * ld;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      double s[VGM_ARITY], B[VGM_NPARAM], b;
#pragma unroll
      for (int a = 0; a < VGM_ARITY; ++a) s[a] = xr[a][t];
      vgm_theta_coeffs(cls, s, B, b);
      const double e = dxrow[t] - b;
      int q = VGM_NPARAM;
#pragma unroll
      for (int i = 0; i < VGM_NPARAM; ++i) {
        acc[i] += B[i] * e;
#pragma unroll
        for (int l = i; l < VGM_NPARAM; ++l) acc[q++] += B[i] * B[l];
      }
    }
  }
  __shared__ double red[VGM_PROG_THREADS / 32][VGM_NSTAT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < VGM_NSTAT; ++i) {
    const double v = vgm_prog_warp_sum(acc[i]);
    if (lane == 0) red[w][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < VGM_NSTAT; i += blockDim.x) {
    double v = 0.0;
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) v += red[ww][i];
    partial[(size_t)blockIdx.x * VGM_NSTAT + i] = v;
  }
}

// 'minimizer' parameter update (opt-in; proxies_for_ode_parameters_and_states.py:114-128 minimises the loss of
// :379-392 with SciPy): per-CTA partials of
//   partial[cta][0]     = sum_{k,t} (f_k(x(t), theta) - (D x_k)(t))^2
//   partial[cta][1 + i] = 2 sum_{k,t} (f_k - D x_k) d f_k / d theta_i
// f_k and its parameter gradient come from the generated vgm_ode_eval, so models that are NOT linear in a parameter
// (Protein Transduction's K_m) are covered.
extern "C" __global__ void __launch_bounds__(VGM_PROG_THREADS)
VGM_PROG(param_loss)(const double* __restrict__ X, const double* __restrict__ DX, int ld, int T, int n_odes,
                     const int* __restrict__ ode_class, const int* __restrict__ ode_idx,
                     const int* __restrict__ ode_state, const double* __restrict__ theta,
                     double* __restrict__ partial) {
  double th[VGM_NPARAM], acc[VGM_NPARAM + 1];
#pragma unroll
  for (int i = 0; i < VGM_NPARAM; ++i) th[i] = theta[i];
#pragma unroll
  for (int i = 0; i <= VGM_NPARAM; ++i) acc[i] = 0.0;
  const int j0 = blockIdx.x * VGM_STAT_ROWS;
  for (int jj = 0; jj < VGM_STAT_ROWS; ++jj) {
    const int j = j0 + jj;
    if (j >= n_odes) break;
    const int cls = ode_class[j];
    const double* dxrow = DX + (size_t)ode_state[j] * ld;
    const double* xr[VGM_ARITY];
#pragma unroll
    for (int a = 0; a < VGM_ARITY; ++a) xr[a] = X + (size_t)ode_idx[(size_t)j * VGM_ARITY + a] * ld;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      double s[VGM_ARITY], f, df[VGM_NPARAM];
#pragma unroll
      for (int a = 0; a < VGM_ARITY; ++a) s[a] = xr[a][t];
      vgm_ode_eval(cls, th, s, f, df);
      const double e = f - dxrow[t];
      acc[0] += e * e;
#pragma unroll
      for (int i = 0; i < VGM_NPARAM; ++i) acc[1 + i] += 2.0 * e * df[i];
    }
  }
  __shared__ double red[VGM_PROG_THREADS / 32][VGM_NPARAM + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i <= VGM_NPARAM; ++i) {
    const double v = vgm_prog_warp_sum(acc[i]);
    if (lane == 0) red[w][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= VGM_NPARAM; i += blockDim.x) {
    double v = 0.0;
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) v += red[ww][i];
    partial[(size_t)blockIdx.x * (VGM_NPARAM + 1) + i] = v;
  }
}

// P2 covariance helper: the rows B_{k,i}(t) of the parameter linearisation (proxies...py:66-71) written out for
// the opt-in parameter covariance  sum_k B_k^T A B_k  (:79).  ODE row j -> rows j*p .. j*p+p-1 of Bout; the padding
// columns t >= T are zeroed.
extern "C" __global__ void __launch_bounds__(VGM_PROG_THREADS)
VGM_PROG(param_B)(const double* __restrict__ X, int ld, int T, int n_rows, const int* __restrict__ ode_class,
                  const int* __restrict__ ode_idx, double* __restrict__ Bout) {
  const int j = blockIdx.x;
  if (j >= n_rows) return;
  const int cls = ode_class[j];
  const double* xr[VGM_ARITY];
#pragma unroll
  for (int a = 0; a < VGM_ARITY; ++a) xr[a] = X + (size_t)ode_idx[(size_t)j * VGM_ARITY + a] * ld;
  for (int t = threadIdx.x; t < ld; t += blockDim.x) {
    double s[VGM_ARITY], B[VGM_NPARAM], b;
#pragma unroll
    for (int i = 0; i < VGM_NPARAM; ++i) B[i] = 0.0;
    if (t < T) {
#pragma unroll
      for (int a = 0; a < VGM_ARITY; ++a) s[a] = xr[a][t];
      vgm_theta_coeffs(cls, s, B, b);
    }
#pragma unroll
    for (int i = 0; i < VGM_NPARAM; ++i) Bout[((size_t)j * VGM_NPARAM + i) * ld + t] = B[i];
  }
}

// P4 (snapshot part): per hidden slot i (state u) and time t   (proxies...py:208-232)
//   Lambda = sum_{k != u} R_uk^2
//   rhs    = - sum_{k != u} R_uk (r_uk - D x_k)    with D x_k from the sweep-start snapshot;
//            for "live" pairs only -R_uk r_uk is added here and R_uk is saved in Rlive so that
//            + R_uk (D x_k^{new}) can be added once state k has been updated (:225)
//   ruu, Ruu = r_uu, R_uu of the self pair (k == u), zero if u is not in c(u)
extern "C" __global__ void __launch_bounds__(VGM_PROG_THREADS)
VGM_PROG(state_assemble)(const double* __restrict__ X, const double* __restrict__ DX0, int ld, int T, int n_hidden,
                         const int* __restrict__ pair_ptr, const int* __restrict__ pair_ode,
                         const int* __restrict__ pair_code, const int* __restrict__ pair_live,
                         const int* __restrict__ pair_self, const int* __restrict__ ode_idx,
                         const int* __restrict__ ode_state, const double* __restrict__ theta,
                         double* __restrict__ Lambda, double* __restrict__ rhs, double* __restrict__ ruu,
                         double* __restrict__ Ruu, double* __restrict__ Rlive) {
  const int i = blockIdx.x;
  if (i >= n_hidden) return;
  double th[VGM_NPARAM];
#pragma unroll
  for (int q = 0; q < VGM_NPARAM; ++q) th[q] = theta[q];
  const int e0 = pair_ptr[i], e1 = pair_ptr[i + 1];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double lam = 0.0, acc = 0.0, r_self = 0.0, R_self = 0.0;
    for (int e = e0; e < e1; ++e) {
      const int j = pair_ode[e];
      double s[VGM_ARITY];
#pragma unroll
      for (int a = 0; a < VGM_ARITY; ++a) s[a] = X[(size_t)ode_idx[(size_t)j * VGM_ARITY + a] * ld + t];
      double R, r;
      vgm_state_coeffs(pair_code[e], th, s, R, r);
      if (pair_self[e]) {
        R_self = R;
        r_self = r;
      } else {
        const int q = pair_live[e];
        if (q < 0) {
          acc -= R * (r - DX0[(size_t)ode_state[j] * ld + t]);
        } else {
          acc -= R * r;
          Rlive[(size_t)q * ld + t] = R;
        }
        lam += R * R;
      }
    }
    const size_t o = (size_t)i * ld + t;
    Lambda[o] = lam;
    rhs[o] = acc;
    ruu[o] = r_self;
    if (Ruu) Ruu[o] = R_self;
  }
}

// P5 helper: the individual R_uk vectors of one hidden slot (proxies...py:211-214), one row per
// coupled ODE in c(u) order; the self pair's row holds R_uu.
extern "C" __global__ void __launch_bounds__(VGM_PROG_THREADS)
VGM_PROG(state_pair_R)(const double* __restrict__ X, int ld, int T, int slot, const int* __restrict__ pair_ptr,
                       const int* __restrict__ pair_ode, const int* __restrict__ pair_code,
                       const int* __restrict__ ode_idx, const double* __restrict__ theta, double* __restrict__ Rout) {
  double th[VGM_NPARAM];
#pragma unroll
  for (int q = 0; q < VGM_NPARAM; ++q) th[q] = theta[q];
  const int e0 = pair_ptr[slot], e1 = pair_ptr[slot + 1];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    for (int e = e0; e < e1; ++e) {
      const int j = pair_ode[e];
      double s[VGM_ARITY];
#pragma unroll
      for (int a = 0; a < VGM_ARITY; ++a) s[a] = X[(size_t)ode_idx[(size_t)j * VGM_ARITY + a] * ld + t];
      double R, r;
      vgm_state_coeffs(pair_code[e], th, s, R, r);
      Rout[(size_t)(e - e0) * ld + t] = R;
    }
  }
}
