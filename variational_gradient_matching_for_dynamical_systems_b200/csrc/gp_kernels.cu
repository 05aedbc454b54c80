// G1 + G2: GP prior covariance grids and the gradient-matching operators.
//   reference: GP_regression.py:160-260 (grids), :306 (D = dC C^-1), :317 (A = ddC - D dC^T)
#include "vgm_common.cuh"

namespace vgm {

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);

// ------------------------------------------------------------------------------------
// K1 kernel_build: one thread per (i, column pair), 16-byte stores.  HBM-write bound.
// ------------------------------------------------------------------------------------
struct KernelParams {
  int type, n;
  double p[6];      // kernel_param[0..5] as given (missing entries 0)
  double sig2;      // kernel_param[-1]^2                               GP_regression.py:232
};

// Second-order jet in (t0, t1): value, d/dt0, d/dt1, d2/dt0dt1.  The kernel types beyond rbf/RQ are written as the
// reference writes them (GP_regression.py:181-231) and differentiated by propagation, which is what its
// kernel.diff(t0).diff(t1) (:236-237) does symbolically.  sqrt((t0-t1)^2) has no derivative at t0 == t1: the
// reference's lambdified expressions give 0/0 = NaN there and so does the jet (inf * 0).
struct Jet {
  double v, a, b, c;
};
__device__ __forceinline__ Jet jconst(double x) { return {x, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ Jet operator+(Jet u, Jet w) { return {u.v + w.v, u.a + w.a, u.b + w.b, u.c + w.c}; }
__device__ __forceinline__ Jet operator-(Jet u, Jet w) { return {u.v - w.v, u.a - w.a, u.b - w.b, u.c - w.c}; }
__device__ __forceinline__ Jet operator+(Jet u, double x) { return {u.v + x, u.a, u.b, u.c}; }
__device__ __forceinline__ Jet operator*(Jet u, double x) { return {u.v * x, u.a * x, u.b * x, u.c * x}; }
__device__ __forceinline__ Jet operator*(Jet u, Jet w) {
  return {u.v * w.v, u.a * w.v + u.v * w.a, u.b * w.v + u.v * w.b, u.c * w.v + u.a * w.b + u.b * w.a + u.v * w.c};
}
// f(u) with f, f', f'' evaluated at u.v
__device__ __forceinline__ Jet jfn(Jet u, double f, double f1, double f2) {
  return {f, f1 * u.a, f1 * u.b, f2 * u.a * u.b + f1 * u.c};
}
__device__ __forceinline__ Jet jexp(Jet u) { const double e = exp(u.v); return jfn(u, e, e, e); }
__device__ __forceinline__ Jet jsin(Jet u) { const double sn = sin(u.v); return jfn(u, sn, cos(u.v), -sn); }
__device__ __forceinline__ Jet jtanh(Jet u) {
  const double th = tanh(u.v), f1 = 1.0 - th * th;
  return jfn(u, th, f1, -2.0 * th * f1);
}
__device__ __forceinline__ Jet jasin(Jet u) {
  const double w = 1.0 - u.v * u.v, f1 = 1.0 / sqrt(w);
  return jfn(u, asin(u.v), f1, u.v * f1 / w);
}
__device__ __forceinline__ Jet jsqrt(Jet u) {
  const double r = sqrt(u.v);
  return jfn(u, r, 0.5 / r, -0.25 / (r * u.v));
}
__device__ __forceinline__ Jet jrecip(Jet u) {
  const double f = 1.0 / u.v;
  return jfn(u, f, -f * f, 2.0 * f * f * f);
}

__device__ __forceinline__ Jet eval_kernel_jet(const KernelParams& kp, double t0v, double t1v) {
  const double PI = 3.141592653589793;
  const Jet t0 = {t0v, 1.0, 0.0, 0.0}, t1 = {t1v, 0.0, 1.0, 0.0};
  const Jet d = t0 - t1;
  const double* p = kp.p;
  Jet k;
  switch (kp.type) {
    case VGM_KERNEL_PERIODIC: {            // :181-184  exp(-2 sin(pi sqrt(dist^2) / p1)^2 / p0^2), dist = d / p1
      const Jet dist = d * (1.0 / p[1]);
      const Jet sn = jsin(jsqrt(dist * dist) * (PI / p[1]));
      k = jexp(sn * sn * (-2.0 / (p[0] * p[0])));
    } break;
    case VGM_KERNEL_LOCALLY_PERIODIC: {    // :185-188  exp(-d^2 / 2 * p1^2) exp(-2 sin(pi sqrt(d^2) / p2)^2 / p1^2)
      const Jet sn = jsin(jsqrt(d * d) * (PI / p[2]));
      k = jexp(d * d * (-0.5 * p[1] * p[1])) * jexp(sn * sn * (-2.0 / (p[1] * p[1])));
    } break;
    case VGM_KERNEL_RBF_LIN: {             // :189-192  exp(-d^2 / p1^2) + p2 + p3 (t0 - p4)(t1 - p4)
      k = jexp(d * d * (-1.0 / (p[1] * p[1]))) + (t0 + (-p[4])) * (t1 + (-p[4])) * p[3] + p[2];
    } break;
    case VGM_KERNEL_SIGMOID: {             // :193-195  asin((p1 + p2 t0 t1) / sqrt((p1 + p2 t0^2 + 1)(p1 + p2 t1^2 + 1)))
      const Jet num = t0 * t1 * p[2] + p[1];
      const Jet den = (t0 * t0 * p[2] + (p[1] + 1.0)) * (t1 * t1 * p[2] + (p[1] + 1.0));
      k = jasin(num * jrecip(jsqrt(den)));
    } break;
    case VGM_KERNEL_SIGMOID2: {            // :196-198  tanh(p0 t0 t1 + p1)
      k = jtanh(t0 * t1 * p[0] + p[1]);
    } break;
    case VGM_KERNEL_EXP_SIN_SQUARED: {     // :199-204  exp(-2 (sin(pi sqrt(d^2) / p1) / p0)^2)
      const Jet sn = jsin(jsqrt(d * d) * (PI / p[1])) * (1.0 / p[0]);
      k = jexp(sn * sn * (-2.0));
    } break;
    default: {                             // VGM_KERNEL_MATERN :210-223, dist = (d / p0)^2 (squared, as written there)
      const Jet dist = d * d * (1.0 / (p[0] * p[0]));
      if (p[1] == 0.5) {
        k = jexp(dist * (-1.0));
      } else if (p[1] == 1.5) {
        const Jet kk = dist * sqrt(3.0);
        k = (kk + 1.0) * jexp(kk * (-1.0));
      } else {
        const Jet kk = dist * sqrt(5.0);
        k = (kk + kk * kk * (1.0 / 3.0) + 1.0) * jexp(kk * (-1.0));
      }
    } break;
  }
  return k * kp.sig2;
}

__device__ __forceinline__ void eval_kernel(const KernelParams& kp, double t0, double t1, double& k, double& dk,
                                            double& ddk) {
  const double d = t0 - t1;
  if (kp.type == VGM_KERNEL_RBF) {
    // k = s^2 exp(-d^2 / (2 l^2))                                  GP_regression.py:176-179
    const double il2 = 1.0 / (kp.p[0] * kp.p[0]);
    k = kp.sig2 * exp(-0.5 * d * d * il2);
    dk = -d * il2 * k;
    ddk = (il2 - d * d * il2 * il2) * k;
  } else if (kp.type == VGM_KERNEL_RQ) {
    // k = s^2 (1 + d^2 / (2 a l^2))^-a                             GP_regression.py:205-209
    const double al = kp.p[0], il2 = 1.0 / (kp.p[1] * kp.p[1]);
    const double base = 1.0 + d * d * il2 / (2.0 * al);
    const double lb = log(base);
    const double pa1 = exp(-(al + 1.0) * lb);   // base^-(a+1)
    k = kp.sig2 * pa1 * base;
    dk = -kp.sig2 * d * il2 * pa1;
    ddk = kp.sig2 * il2 * pa1 - kp.sig2 * d * d * il2 * il2 * (al + 1.0) / al * pa1 / base;
  } else {
    const Jet j = eval_kernel_jet(kp, t0, t1);
    k = j.v;
    dk = j.a;
    ddk = j.c;
  }
}

__global__ void kernel_build_kernel(KernelParams kp, const double* __restrict__ ta, int Ta,
                                    const double* __restrict__ tb, int Tb, int ld,
                                    double* __restrict__ C, double* __restrict__ dC, double* __restrict__ ddC) {
  const int j = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int i = blockIdx.y;
  if (j >= Tb) return;
  const double ti = ta[i];
  double k0, dk0, ddk0, k1 = 0, dk1 = 0, ddk1 = 0;
  eval_kernel(kp, ti, tb[j], k0, dk0, ddk0);
  const bool two = (j + 1 < Tb);
  if (two) eval_kernel(kp, ti, tb[j + 1], k1, dk1, ddk1);
  const size_t o = (size_t)i * ld + j;
  if (two || (j + 1 < ld)) {   // ld is even, so the pad slot exists: keep it zero
    if (C) *reinterpret_cast<double2*>(C + o) = make_double2(k0, k1);
    if (dC) *reinterpret_cast<double2*>(dC + o) = make_double2(dk0, dk1);
    if (ddC) *reinterpret_cast<double2*>(ddC + o) = make_double2(ddk0, ddk1);
  } else {
    if (C) C[o] = k0;
    if (dC) dC[o] = dk0;
    if (ddC) ddC[o] = ddk0;
  }
}

static int kernel_build_one(const KernelParams& kp, const double* ta, int Ta, const double* tb, int Tb, int ld,
                            double* C, double* dC, double* ddC, cudaStream_t s) {
  if (Ta <= 0 || Tb <= 0 || (!C && !dC && !ddC)) return VGM_OK;
  dim3 block(128), grid(ceil_div(ceil_div(Tb, 2), 128), Ta);
  kernel_build_kernel<<<grid, block, 0, s>>>(kp, ta, Ta, tb, Tb, ld, C, dC, ddC);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// ------------------------------------------------------------------------------------
// Transpose
// ------------------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ in, int ld_in, double* __restrict__ out, int ld_out,
                                 int rows, int cols) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rr = r0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (rr < rows && cc < cols) ? in[(size_t)rr * ld_in + cc] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int orow = c0 + r, ocol = r0 + threadIdx.x;
    if (orow < cols && ocol < rows) out[(size_t)orow * ld_out + ocol] = tile[threadIdx.x][r];
  }
}

int transpose(const double* in, int ld_in, double* out, int ld_out, int rows, int cols, cudaStream_t s) {
  dim3 block(32, 8), grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, block, 0, s>>>(in, ld_in, out, ld_out, rows, cols);
  VGM_LAUNCH_OK();
  return VGM_OK;
}

// ------------------------------------------------------------------------------------
// Blocked Cholesky + right-hand triangular solves built from the DMMA GEMM.
// ------------------------------------------------------------------------------------
constexpr int NB = 32;

// Factor the nb x nb diagonal block at (j0, j0) in shared memory (one CTA) and write L11
// back in place (strict upper part of the block zeroed).
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ S, int ld, int j0, int nb,
                                                         int* __restrict__ bad_pivot) {
  __shared__ double L[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    L[r][c] = (r < nb && c < nb && c <= r) ? S[(size_t)(j0 + r) * ld + j0 + c] : 0.0;
  }
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    const double piv = L[c][c];
    __syncthreads();
    if (!(piv > 0.0) && tid == 0) atomicCAS(bad_pivot, 0, j0 + c + 1);
    const double d = sqrt(piv);
    if (tid == 0) L[c][c] = d;
    for (int r = c + 1 + tid; r < nb; r += blockDim.x) L[r][c] /= d;
    __syncthreads();
    // trailing update of the lower triangle: a[r][cc] -= l[r][c] * l[cc][c],  c < cc <= r
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int r = e / NB, cc = e % NB;
      if (r < nb && cc > c && cc <= r) L[r][cc] -= L[r][c] * L[cc][c];
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    if (r < nb && c < nb) S[(size_t)(j0 + r) * ld + j0 + c] = (c <= r) ? L[r][c] : 0.0;
  }
}

// Row-wise small triangular solves against the diagonal block of L at (j0, j0):
//   mode 0:  x L_jj^T = a  (forward)     mode 1:  x L_jj = a  (backward)
// applied in place to R[row][j0 .. j0+nb) for every row < n_rows.
__global__ void __launch_bounds__(128) trsm_diag_rows_kernel(const double* __restrict__ Lm, int ldl, int j0, int nb,
                                                             double* __restrict__ R, int ldr, int n_rows, int mode) {
  __shared__ double L[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    L[r][c] = (r < nb && c < nb && c <= r) ? Lm[(size_t)(j0 + r) * ldl + j0 + c] : 0.0;
  }
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + tid;
  if (row >= n_rows) return;
  double x[NB];
  double* p = R + (size_t)row * ldr + j0;
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? p[c] : 0.0;
  if (mode == 0) {
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (c < nb) {
        double v = x[c];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < c) v -= x[k] * L[c][k];
        x[c] = v / L[c][c];
      }
    }
  } else {
#pragma unroll
    for (int c = NB - 1; c >= 0; --c) {
      if (c < nb) {
        double v = x[c];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k > c && k < nb) v -= x[k] * L[k][c];
        x[c] = v / L[c][c];
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
    if (c < nb) p[c] = x[c];
}

__global__ void set_identity_kernel(double* __restrict__ I, int ld, int T) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < ld) I[(size_t)i * ld + j] = (i == j && j < T) ? 1.0 : 0.0;
}

// In-place lower Cholesky S = L L^T (S overwritten by L, strict upper part zeroed).
static int potrf(double* S, int ld, int T, int* bad_pivot, cudaStream_t s) {
  for (int j0 = 0; j0 < T; j0 += NB) {
    const int nb = min(NB, T - j0);
    const int below = T - j0 - nb;
    potrf_diag_kernel<<<1, 256, 0, s>>>(S, ld, j0, nb, bad_pivot);
    VGM_LAUNCH_OK();
    if (below > 0) {
      // L21 = A21 L11^-T, one thread per row
      trsm_diag_rows_kernel<<<ceil_div(below, 128), 128, 0, s>>>(S, ld, j0, nb, S + (size_t)(j0 + nb) * ld, ld, below, 0);
      VGM_LAUNCH_OK();
    }
    if (below > 0) {
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      double* L21 = S + (size_t)(j0 + nb) * ld + j0;
      double* A22 = S + (size_t)(j0 + nb) * ld + (j0 + nb);
      g.A = L21; g.lda = ld;
      g.B = L21; g.ldb = ld;
      g.C = A22; g.ldc = ld; g.C0 = A22;
      g.N = below; g.M = below; g.K = nb;
      g.alpha = -1.0; g.beta = 1.0;
      VGM_TRY(dgemm_nt(g, s));
    }
  }
  return VGM_OK;
}

// R <- R F1^-T F2^-1, R is [n_rows x T] row-major, in place; F1, F2 lower triangular, F2T = F2^T.
//   Cholesky S = L L^T:        F1 = F2 = L        ->  R S^-1
//   LU       P S = Lu U:       F1 = U^T, F2 = Lu  ->  R U^-1 Lu^-1  (the column permutation is applied by the caller)
static int tri_solve_right(const double* F1, const double* F2, const double* F2T, int ld, int T, double* R, int ldr,
                           int n_rows, cudaStream_t s) {
  const double *L = F1, *LT = F2T;
  // W L^T = R : column blocks ascending;  W[:,j] = (R[:,j] - sum_{k<j} W[:,k] L[j,k]^T) L_jj^-T
  for (int j0 = 0; j0 < T; j0 += NB) {
    const int nb = min(NB, T - j0);
    if (j0 > 0) {
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = R; g.lda = ldr;
      g.B = L + (size_t)j0 * ld; g.ldb = ld;
      g.C = R + j0; g.ldc = ldr; g.C0 = g.C;
      g.N = n_rows; g.M = nb; g.K = j0;
      g.alpha = -1.0; g.beta = 1.0;
      VGM_TRY(dgemm_nt(g, s));
    }
    trsm_diag_rows_kernel<<<ceil_div(n_rows, 128), 128, 0, s>>>(L, ld, j0, nb, R, ldr, n_rows, 0);
    VGM_LAUNCH_OK();
  }
  // Y L = W : column blocks descending;  Y[:,j] = (W[:,j] - sum_{k>j} Y[:,k] L[k,j]) L_jj^-1
  const int nblk = ceil_div(T, NB);
  for (int jb = nblk - 1; jb >= 0; --jb) {
    const int j0 = jb * NB, nb = min(NB, T - j0), j1 = j0 + nb;
    if (j1 < T) {
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = R + j1; g.lda = ldr;
      g.B = LT + (size_t)j0 * ld + j1; g.ldb = ld;
      g.C = R + j0; g.ldc = ldr; g.C0 = g.C;
      g.N = n_rows; g.M = nb; g.K = T - j1;
      g.alpha = -1.0; g.beta = 1.0;
      VGM_TRY(dgemm_nt(g, s));
    }
    trsm_diag_rows_kernel<<<ceil_div(n_rows, 128), 128, 0, s>>>(F2, ld, j0, nb, R, ldr, n_rows, 1);
    VGM_LAUNCH_OK();
  }
  return VGM_OK;
}

static int chol_solve_right(const double* L, const double* LT, int ld, int T, double* R, int ldr, int n_rows,
                            cudaStream_t s) {
  return tri_solve_right(L, L, LT, ld, T, R, ldr, n_rows, s);
}

// ------------------------------------------------------------------------------------
// LU with partial pivoting (what the reference's np.linalg.solve / LAPACK gesv does, GP_regression.py:306) for
// covariance matrices that are not positive definite: the tanh "sigmoid2" kernel and the reference's Matern
// 1.5 / 2.5 written on the SQUARED distance (:212-223) are indefinite.  One CTA, right-looking, rank-1 updates
// out of L2: O(T^3 / 3) on one SM is tens of milliseconds at T = 1000, a one-time set-up cost for rare kernels.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) lu_single_cta_kernel(double* __restrict__ Am, int ld, int T,
                                                             int* __restrict__ perm, int* __restrict__ flag) {
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = tid; i < T; i += blockDim.x) perm[i] = i;
  __syncthreads();
  for (int k = 0; k < T; ++k) {
    // pivot: first row of maximal |a[r][k]|, r >= k (LAPACK idamax)
    double bv = -1.0;
    int bi = k;
    for (int r = k + tid; r < T; r += blockDim.x) {
      const double v = fabs(Am[(size_t)r * ld + k]);
      if (v > bv) { bv = v; bi = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < nw ? red_v[lane] : -1.0;
      bi = lane < nw ? red_i[lane] : T;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        s_piv = bi;
        if (!(bv > 0.0)) atomicCAS(flag, 0, k + 1);        // exactly singular (or NaN) column
        const int t = perm[k]; perm[k] = perm[bi]; perm[bi] = t;
      }
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k)
      for (int c = tid; c < T; c += blockDim.x) {
        const double a = Am[(size_t)k * ld + c], b = Am[(size_t)p * ld + c];
        Am[(size_t)k * ld + c] = b;
        Am[(size_t)p * ld + c] = a;
      }
    __syncthreads();
    const double inv = 1.0 / Am[(size_t)k * ld + k];
    const double* rowk = Am + (size_t)k * ld;
    for (int r = k + 1 + warp; r < T; r += nw) {
      double* rowr = Am + (size_t)r * ld;
      const double l = rowr[k] * inv;
      for (int c = k + 1 + lane; c < T; c += 32) rowr[c] -= l * rowk[c];
      __syncwarp();
      if (lane == 0) rowr[k] = l;
    }
    __syncthreads();
  }
}

// LUm (packed) -> Lu (unit lower, explicit ones), UT = U^T (lower)
__global__ void lu_unpack_kernel(const double* __restrict__ LUm, int ld, int T, double* __restrict__ Lu,
                                 double* __restrict__ UT) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= ld) return;
  const bool in = c < T;
  Lu[(size_t)r * ld + c] = !in ? 0.0 : (c < r ? LUm[(size_t)r * ld + c] : (c == r ? 1.0 : 0.0));
  UT[(size_t)r * ld + c] = (in && c <= r) ? LUm[(size_t)c * ld + r] : 0.0;
}

// D[row][perm[i]] = V[row][i]
__global__ void permute_cols_kernel(const double* __restrict__ V, int ldv, double* __restrict__ Dm, int ldd, int T,
                                    const int* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
  if (i < T) Dm[(size_t)row * ldd + perm[i]] = V[(size_t)row * ldv + i];
}

struct CholWs {
  double* L;
  double* LT;
  int* flag;
  double* extra;   // 3 more matrices (LU route)
  int* perm;
};

// 2 matrices for the Cholesky route; the LU route (vgm_gm_operators_lu) needs 5 plus the permutation
static size_t chol_ws_bytes(int T, int ld) {
  return 5 * align_up((size_t)T * ld * sizeof(double), 256) + align_up((size_t)T * sizeof(int), 256) + 256;
}

static int chol_ws_carve(void* ws, size_t ws_bytes, int T, int ld, CholWs& w) {
  VGM_REQUIRE(ws && ws_bytes >= chol_ws_bytes(T, ld), "workspace too small");
  VGM_REQUIRE(((uintptr_t)ws % 16) == 0, "workspace must be 16-byte aligned");
  char* p = (char*)ws;
  const size_t mat = align_up((size_t)T * ld * sizeof(double), 256);
  w.L = (double*)p;
  w.LT = (double*)(p + mat);
  w.flag = (int*)(p + 5 * mat);
  w.extra = (double*)(p + 2 * mat);
  w.perm = (int*)(p + 5 * mat + 256);
  return VGM_OK;
}

static int factor_into_ws(const double* S, int T, int ld, CholWs& w, cudaStream_t s) {
  VGM_CUDA_OK(cudaMemcpyAsync(w.L, S, (size_t)T * ld * sizeof(double), cudaMemcpyDeviceToDevice, s));
  VGM_CUDA_OK(cudaMemsetAsync(w.flag, 0, sizeof(int), s));
  VGM_TRY(potrf(w.L, ld, T, w.flag, s));
  VGM_TRY(transpose(w.L, ld, w.LT, ld, T, T, s));
  return VGM_OK;
}

static int check_pivot(CholWs& w, cudaStream_t s) {
  int flag = 0;
  VGM_CUDA_OK(cudaMemcpyAsync(&flag, w.flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  if (flag != 0) {
    set_error("Cholesky: pivot %d is not positive (matrix not numerically SPD)", flag - 1);
    return VGM_ENOTSPD;
  }
  return VGM_OK;
}

}  // namespace vgm

using namespace vgm;

extern "C" int vgm_kernel_build(int kernel_type, const double* param_host, int n_param, const double* t, int T,
                                const double* t2, int T2, int ld, int ld2, double* C, double* dC, double* ddC,
                                double* Cobs, double* Cso, vgm_stream_t stream) {
  VGM_REQUIRE(param_host && n_param >= 2, "kernel parameters");
  VGM_REQUIRE(kernel_type >= VGM_KERNEL_RBF && kernel_type <= VGM_KERNEL_MATERN, "kernel_type");
  static const int need[] = {2, 3, 2, 3, 5, 3, 2, 2, 2};   // entries of kernel_param each type reads (see vgm_b200.h)
  if (n_param < need[kernel_type]) {
    set_error("kernel type %d reads %d entries of kernel_param, %d given", kernel_type, need[kernel_type], n_param);
    return VGM_EINVAL;
  }
  if (kernel_type == VGM_KERNEL_MATERN && param_host[1] != 0.5 && param_host[1] != 1.5 && param_host[1] != 2.5) {
    set_error("matern: nu = %g; only 0.5, 1.5, 2.5 evaluate in the reference (GP_regression.py:215-223)", param_host[1]);
    return VGM_EINVAL;
  }
  VGM_REQUIRE(T >= 0 && T2 >= 0 && (ld % 2) == 0 && (ld2 % 2) == 0 && ld >= T && ld2 >= T2, "shape");
  KernelParams kp;
  kp.type = kernel_type;
  kp.n = n_param;
  for (int i = 0; i < 6; ++i) kp.p[i] = i < n_param ? param_host[i] : 0.0;
  kp.sig2 = param_host[n_param - 1] * param_host[n_param - 1];
  cudaStream_t s = (cudaStream_t)stream;
  if (T > 0) VGM_TRY(kernel_build_one(kp, t, T, t, T, ld, C, dC, ddC, s));
  if (T2 > 0 && Cobs) VGM_TRY(kernel_build_one(kp, t2, T2, t2, T2, ld2, Cobs, nullptr, nullptr, s));
  if (T > 0 && T2 > 0 && Cso) VGM_TRY(kernel_build_one(kp, t, T, t2, T2, ld2, Cso, nullptr, nullptr, s));
  return VGM_OK;
}

extern "C" int vgm_transpose(const double* in, int ld_in, double* out, int ld_out, int rows, int cols,
                             vgm_stream_t stream) {
  VGM_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "transpose arguments");
  return transpose(in, ld_in, out, ld_out, rows, cols, (cudaStream_t)stream);
}

extern "C" size_t vgm_gm_operators_ws_bytes(int T, int ld) { return chol_ws_bytes(T, ld); }

extern "C" int vgm_gm_operators(const double* C, const double* dC, const double* ddC, int T, int ld, int batch,
                                double* D, double* A, double* Cinv, void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(C && dC && D && T > 0 && (ld % 2) == 0 && ld >= T && batch >= 1, "gm_operators arguments");
  VGM_REQUIRE(!A || ddC, "A requested without ddC");
  cudaStream_t s = (cudaStream_t)stream;
  CholWs w;
  VGM_TRY(chol_ws_carve(ws, ws_bytes, T, ld, w));
  const size_t mat = (size_t)T * ld;
  for (int b = 0; b < batch; ++b) {
    const double* Cb = C + b * mat;
    const double* dCb = dC + b * mat;
    double* Db = D + b * mat;
    VGM_TRY(factor_into_ws(Cb, T, ld, w, s));
    // D = dC C^-1                                                   GP_regression.py:306
    VGM_CUDA_OK(cudaMemcpyAsync(Db, dCb, mat * sizeof(double), cudaMemcpyDeviceToDevice, s));
    VGM_TRY(chol_solve_right(w.L, w.LT, ld, T, Db, ld, T, s));
    if (A) {
      // A = ddC - D dC^T                                            GP_regression.py:317
      vgm_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.A = Db; g.lda = ld;
      g.B = dCb; g.ldb = ld;
      g.C = A + b * mat; g.ldc = ld; g.C0 = ddC + b * mat;
      g.N = T; g.M = T; g.K = T;
      g.alpha = -1.0; g.beta = 1.0;
      VGM_TRY(dgemm_nt(g, s));
    }
    if (Cinv) {
      double* Ib = Cinv + b * mat;
      set_identity_kernel<<<dim3(ceil_div(ld, 128), T), 128, 0, s>>>(Ib, ld, T);
      VGM_LAUNCH_OK();
      VGM_TRY(chol_solve_right(w.L, w.LT, ld, T, Ib, ld, T, s));
    }
    VGM_TRY(check_pivot(w, s));
  }
  return VGM_OK;
}

extern "C" int vgm_gm_operators_lu(const double* C, const double* dC, const double* ddC, int T, int ld, double* D,
                                   double* A, void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(C && dC && D && T > 0 && (ld % 2) == 0 && ld >= T, "gm_operators_lu arguments");
  VGM_REQUIRE(!A || ddC, "A requested without ddC");
  cudaStream_t s = (cudaStream_t)stream;
  CholWs w;
  VGM_TRY(chol_ws_carve(ws, ws_bytes, T, ld, w));
  const size_t mat = (size_t)T * ld, mat_al = align_up(mat * sizeof(double), 256) / sizeof(double);
  double *LUm = w.extra, *UT = w.extra + mat_al, *V = w.extra + 2 * mat_al;
  VGM_CUDA_OK(cudaMemcpyAsync(LUm, C, mat * sizeof(double), cudaMemcpyDeviceToDevice, s));
  VGM_CUDA_OK(cudaMemsetAsync(w.flag, 0, sizeof(int), s));
  lu_single_cta_kernel<<<1, 1024, 0, s>>>(LUm, ld, T, w.perm, w.flag);
  VGM_LAUNCH_OK();
  lu_unpack_kernel<<<dim3(ceil_div(ld, 128), T), 128, 0, s>>>(LUm, ld, T, w.L, UT);
  VGM_LAUNCH_OK();
  VGM_TRY(transpose(w.L, ld, w.LT, ld, T, T, s));
  // P C = Lu U  =>  D = dC C^-1 = (dC U^-1 Lu^-1) P                  GP_regression.py:306
  VGM_CUDA_OK(cudaMemcpyAsync(V, dC, mat * sizeof(double), cudaMemcpyDeviceToDevice, s));
  VGM_TRY(tri_solve_right(UT, w.L, w.LT, ld, T, V, ld, T, s));
  VGM_CUDA_OK(cudaMemsetAsync(D, 0, mat * sizeof(double), s));
  permute_cols_kernel<<<dim3(ceil_div(T, 128), T), 128, 0, s>>>(V, ld, D, ld, T, w.perm);
  VGM_LAUNCH_OK();
  if (A) {
    vgm_gemm_args g;                                                 // A = ddC - D dC^T, :317
    memset(&g, 0, sizeof(g));
    g.A = D; g.lda = ld;
    g.B = dC; g.ldb = ld;
    g.C = A; g.ldc = ld; g.C0 = ddC;
    g.N = T; g.M = T; g.K = T;
    g.alpha = -1.0; g.beta = 1.0;
    VGM_TRY(dgemm_nt(g, s));
  }
  int flag = 0;
  VGM_CUDA_OK(cudaMemcpyAsync(&flag, w.flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  if (flag != 0) {
    set_error("LU: column %d has no non-zero pivot (matrix singular)", flag - 1);
    return VGM_ENOTSPD;
  }
  return VGM_OK;
}

// ------------------------------------------------------------------------------------
// GP_regression.fitting_state_observations, GP_regression.py:33-51: the GP fit of the observed states.
// ------------------------------------------------------------------------------------
namespace vgm {
__global__ void add_to_diagonal_kernel(double* __restrict__ S, int ld, int T, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) S[(size_t)i * ld + i] += v;
}
static size_t gp_fit_ws_bytes(int T, int T2, int ld2) {
  return chol_ws_bytes(T2, ld2) + align_up((size_t)T2 * ld2 * sizeof(double), 256) /*S*/ +
         align_up((size_t)ld2 * sizeof(double), 256) /*alpha*/ + align_up((size_t)T * ld2 * sizeof(double), 256) /*R*/;
}
}  // namespace vgm

extern "C" size_t vgm_gp_fit_ws_bytes(int T, int T2, int ld2) { return gp_fit_ws_bytes(T, T2, ld2); }

extern "C" int vgm_gp_fit(const double* Cobs, const double* Cso, const double* Cmat, int T, int T2, int ld, int ld2,
                          const double* Y, const double* obs_var_host, int n_obs, double* mean, double* pred_cov,
                          void* ws, size_t ws_bytes, vgm_stream_t stream) {
  VGM_REQUIRE(Cobs && Cso && T > 0 && T2 > 0 && ld >= T && ld2 >= T2 && (ld % 2) == 0 && (ld2 % 2) == 0, "gp_fit shapes");
  VGM_REQUIRE(n_obs >= 0 && (n_obs == 0 || (Y && obs_var_host && mean)), "gp_fit observations");
  VGM_REQUIRE(!pred_cov || Cmat, "the predictive covariance needs cov");
  VGM_REQUIRE(ws && ws_bytes >= gp_fit_ws_bytes(T, T2, ld2) && ((uintptr_t)ws % 256) == 0, "gp_fit workspace");
  cudaStream_t s = (cudaStream_t)stream;
  char* p = (char*)ws;
  CholWs w;
  VGM_TRY(chol_ws_carve(p, chol_ws_bytes(T2, ld2), T2, ld2, w));
  p += align_up(chol_ws_bytes(T2, ld2), 256);
  double* S = (double*)p; p += align_up((size_t)T2 * ld2 * sizeof(double), 256);
  double* alpha = (double*)p; p += align_up((size_t)ld2 * sizeof(double), 256);
  double* R = (double*)p;
  VGM_CUDA_OK(cudaMemcpyAsync(S, Cobs, (size_t)T2 * ld2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  bool factored = false;
  for (int i = 0; i < n_obs; ++i) {
    // cov_obs is REASSIGNED inside the reference's loop (:42): observed state i sees sum_{j<=i} sigma_j^2
    add_to_diagonal_kernel<<<ceil_div(T2, 128), 128, 0, s>>>(S, ld2, T2, obs_var_host[i]);
    VGM_LAUNCH_OK();
    VGM_TRY(factor_into_ws(S, T2, ld2, w, s));
    VGM_TRY(check_pivot(w, s));
    factored = true;
    // alpha = y_i S^-1 (S symmetric), mean_i = Cso alpha                                            (:43)
    VGM_CUDA_OK(cudaMemcpyAsync(alpha, Y + (size_t)i * ld2, (size_t)ld2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    VGM_TRY(chol_solve_right(w.L, w.LT, ld2, T2, alpha, ld2, 1, s));
    vgm_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = alpha; g.lda = ld2; g.B = Cso; g.ldb = ld2; g.C = mean + (size_t)i * ld; g.ldc = ld;
    g.N = 1; g.M = T; g.K = T2; g.alpha = 1.0;
    VGM_TRY(dgemm_nt(g, s));
  }
  if (pred_cov) {
    // state_pred_cov = cov - Cso solve(cov_obs, Cso^T) with the FINAL accumulated cov_obs            (:45)
    if (!factored) {
      VGM_TRY(factor_into_ws(S, T2, ld2, w, s));
      VGM_TRY(check_pivot(w, s));
    }
    VGM_CUDA_OK(cudaMemcpyAsync(R, Cso, (size_t)T * ld2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    VGM_TRY(chol_solve_right(w.L, w.LT, ld2, T2, R, ld2, T, s));
    vgm_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = R; g.lda = ld2; g.B = Cso; g.ldb = ld2; g.C = pred_cov; g.ldc = ld; g.C0 = Cmat;
    g.N = T; g.M = T; g.K = T2; g.alpha = -1.0; g.beta = 1.0;
    VGM_TRY(dgemm_nt(g, s));
  }
  return VGM_OK;
}

extern "C" int vgm_spd_inverse(const double* S, int T, int ld, double* Sinv, void* ws, size_t ws_bytes,
                               vgm_stream_t stream) {
  VGM_REQUIRE(S && Sinv && T > 0 && (ld % 2) == 0 && ld >= T, "spd_inverse arguments");
  cudaStream_t s = (cudaStream_t)stream;
  CholWs w;
  VGM_TRY(chol_ws_carve(ws, ws_bytes, T, ld, w));
  VGM_TRY(factor_into_ws(S, T, ld, w, s));
  set_identity_kernel<<<dim3(ceil_div(ld, 128), T), 128, 0, s>>>(Sinv, ld, T);
  VGM_LAUNCH_OK();
  VGM_TRY(chol_solve_right(w.L, w.LT, ld, T, Sinv, ld, T, s));
  return check_pivot(w, s);
}
