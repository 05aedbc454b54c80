// G1 + G2: GP prior covariance grids and the gradient-matching operators.
//   reference: GP_regression.py:160-260 (grids), :306 (D = dC C^-1), :317 (A = ddC - D dC^T)
#include "vgm_common.cuh"

namespace vgm {

int dgemm_nt(const vgm_gemm_args& a, cudaStream_t s);

// ------------------------------------------------------------------------------------
// K1 kernel_build: one thread per (i, column pair), 16-byte stores.  HBM-write bound.
// ------------------------------------------------------------------------------------
struct KernelParams {
  int type;
  double p0, p1, sig2;
};

__device__ __forceinline__ void eval_kernel(const KernelParams& kp, double d, double& k, double& dk, double& ddk) {
  if (kp.type == VGM_KERNEL_RBF) {
    // k = s^2 exp(-d^2 / (2 l^2))                                  GP_regression.py:176-179
    const double il2 = 1.0 / (kp.p0 * kp.p0);
    k = kp.sig2 * exp(-0.5 * d * d * il2);
    dk = -d * il2 * k;
    ddk = (il2 - d * d * il2 * il2) * k;
  } else {
    // k = s^2 (1 + d^2 / (2 a l^2))^-a                             GP_regression.py:205-209
    const double al = kp.p0, il2 = 1.0 / (kp.p1 * kp.p1);
    const double base = 1.0 + d * d * il2 / (2.0 * al);
    const double lb = log(base);
    const double pa1 = exp(-(al + 1.0) * lb);   // base^-(a+1)
    k = kp.sig2 * pa1 * base;
    dk = -kp.sig2 * d * il2 * pa1;
    ddk = kp.sig2 * il2 * pa1 - kp.sig2 * d * d * il2 * il2 * (al + 1.0) / al * pa1 / base;
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
  eval_kernel(kp, ti - tb[j], k0, dk0, ddk0);
  const bool two = (j + 1 < Tb);
  if (two) eval_kernel(kp, ti - tb[j + 1], k1, dk1, ddk1);
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

// R <- R L^-T L^-1  (= R S^-1), R is [n_rows x T] row-major, in place.  LT = L^T.
static int chol_solve_right(const double* L, const double* LT, int ld, int T, double* R, int ldr, int n_rows,
                            cudaStream_t s) {
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
    trsm_diag_rows_kernel<<<ceil_div(n_rows, 128), 128, 0, s>>>(L, ld, j0, nb, R, ldr, n_rows, 1);
    VGM_LAUNCH_OK();
  }
  return VGM_OK;
}

struct CholWs {
  double* L;
  double* LT;
  int* flag;
};

static size_t chol_ws_bytes(int T, int ld) { return 2 * align_up((size_t)T * ld * sizeof(double), 256) + 256; }

static int chol_ws_carve(void* ws, size_t ws_bytes, int T, int ld, CholWs& w) {
  VGM_REQUIRE(ws && ws_bytes >= chol_ws_bytes(T, ld), "workspace too small");
  VGM_REQUIRE(((uintptr_t)ws % 16) == 0, "workspace must be 16-byte aligned");
  char* p = (char*)ws;
  const size_t mat = align_up((size_t)T * ld * sizeof(double), 256);
  w.L = (double*)p;
  w.LT = (double*)(p + mat);
  w.flag = (int*)(p + 2 * mat);
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
  VGM_REQUIRE(kernel_type == VGM_KERNEL_RBF || kernel_type == VGM_KERNEL_RQ, "kernel_type");
  VGM_REQUIRE(kernel_type != VGM_KERNEL_RQ || n_param >= 3, "RQ needs [alpha, lengthscale, sigma]");
  VGM_REQUIRE(T >= 0 && T2 >= 0 && (ld % 2) == 0 && (ld2 % 2) == 0 && ld >= T && ld2 >= T2, "shape");
  KernelParams kp;
  kp.type = kernel_type;
  kp.p0 = param_host[0];
  kp.p1 = param_host[1];
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
