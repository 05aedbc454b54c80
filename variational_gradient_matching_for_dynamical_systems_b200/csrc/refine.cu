// Vector kernels of the mixed-precision solver (one CTA per system row; every kernel streams each array
// once, 16-byte vector accesses, row-wide reductions through shared memory) and the band-width probe.
//
// The correction solves A_u d = r_u, A_u = M + diag(Lambda_u) (proxies_for_ode_parameters_and_states.py:232,
// 245,262 solve the same system with LAPACK), run as FP32 preconditioned CG whose two operator products
// per step are tcgen05 GEMMs (tc_gemm.cu) writing raw FP32 accumulators; everything elementwise and every
// dot product of a CG step lives in the two kernels ir_update / ir_dir below.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "vgm_common.cuh"
#include "vgm_profile.cuh"
#include "refine.cuh"

namespace vgm {

constexpr int IR_THREADS = 256;

size_t ir_ws_bytes(int n, int ld) {
  const size_t f32 = align_up((size_t)n * ld * sizeof(float), 256);
  const size_t b16 = align_up((size_t)n * ld * sizeof(__nv_bfloat16), 256);
  const size_t vec = align_up((size_t)n * sizeof(double), 256);
  const size_t part = align_up((size_t)n * ((ld + 127) / 128) * sizeof(float), 256);
  return 4 * f32 + 6 * b16 + 2 * vec + part;
}

void ir_carve(char*& p, int n, int ld, IrBuf& b) {
  const size_t f32 = align_up((size_t)n * ld * sizeof(float), 256);
  const size_t b16 = align_up((size_t)n * ld * sizeof(__nv_bfloat16), 256);
  const size_t vec = align_up((size_t)n * sizeof(double), 256);
  b.r32 = (float*)p; p += f32;
  b.d32 = (float*)p; p += f32;
  b.acc32 = (float*)p; p += f32;
  b.lam32 = (float*)p; p += f32;
  b.p_hi = (__nv_bfloat16*)p; p += b16;
  b.p_lo = (__nv_bfloat16*)p; p += b16;
  b.q_hi = (__nv_bfloat16*)p; p += b16;
  b.q_lo = (__nv_bfloat16*)p; p += b16;
  b.rsb = (__nv_bfloat16*)p; p += b16;
  b.s16 = (__nv_bfloat16*)p; p += b16;
  b.sigma = (double*)p; p += vec;
  b.rz = (double*)p; p += vec;
  b.rz_tiles = (ld + 127) / 128;
  b.rz_part = (float*)p; p += align_up((size_t)n * b.rz_tiles * sizeof(float), 256);
}

// ---- 4-element vector helpers (rows are 16-byte aligned: ld % 8 == 0) --------------------------------
struct bf4 { __nv_bfloat162 a, b; };
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_bf4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void st_bf4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&lo);
  u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float bf_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// diagonal scaling of the preconditioner, s = bf16(sqrt((shift + c) / (Lambda + c))), from the FP32 copy of Lambda:
// ir_inner_init stores it (s16) for ir_dir, ir_update recomputes the identical value instead of reading it
__device__ __forceinline__ float precond_scale(float lam, float num, float c) {
  return (c > 0.f) ? bf_round(sqrtf(num / (lam + c))) : 1.f;
}
// The GEMM writes columns < T only: whatever sits in the row padding of acc32 must not reach a dot product.
__device__ __forceinline__ float4 ld_acc4(const float* p, int i, int T) {
  float4 v = ld_f4(p);
  if (i + 4 > T) {
    if (i + 1 >= T) v.y = 0.f;
    if (i + 2 >= T) v.z = 0.f;
    if (i + 3 >= T) v.w = 0.f;
  }
  return v;
}

// Loop over the 4-element groups of one row; the last group may be partial (T % 4 != 0 is handled by the
// zero padding of every row up to ld, which all kernels preserve).
#define IR_ROW_LOOP(i, T) for (int i = 4 * threadIdx.x; i < (T); i += 4 * IR_THREADS)

__global__ void __launch_bounds__(IR_THREADS)
ir_inner_init_kernel(IrBuf b, const double* __restrict__ R, const double* __restrict__ Lambda,
                     const double* __restrict__ rr, const int* __restrict__ act, const int* __restrict__ n_dev, int T,
                     int ld, double shift, double c) {
  pdl_wait();
  const int j = blockIdx.x;
  if (n_dev && j >= *n_dev) return;
  const int slot = act[j];
  const double sig = sqrt(rr[slot]);
  const double inv = sig > 0.0 ? 1.0 / sig : 0.0;
  const size_t oj = (size_t)j * ld, os = (size_t)slot * ld;
  // same active list as in the previous correction: row j still belongs to the same system, lam32 / s16 are valid
  const bool remap = !n_dev || n_dev[2];
  const bool vec = (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(Lambda)) & 15) == 0;
  IR_ROW_LOOP(i, T) {
    float4 r, lam, sv;
    float* rp = &r.x; float* lp = &lam.x; float* sp = &sv.x;
    double rd[4], ld4[4];
    if (vec) {                                   // 16-byte loads (a group may reach into the row padding, ld >= T rounded up to 4: masked below)
      *reinterpret_cast<double2*>(rd) = *reinterpret_cast<const double2*>(R + os + i);
      *reinterpret_cast<double2*>(rd + 2) = *reinterpret_cast<const double2*>(R + os + i + 2);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) rd[e] = (i + e < T) ? R[os + i + e] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = i + e < T;
      rp[e] = ok ? (float)(rd[e] * inv) : 0.f;
    }
    if (remap) {
      if (vec) {
        *reinterpret_cast<double2*>(ld4) = *reinterpret_cast<const double2*>(Lambda + os + i);
        *reinterpret_cast<double2*>(ld4 + 2) = *reinterpret_cast<const double2*>(Lambda + os + i + 2);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) ld4[e] = (i + e < T) ? Lambda[os + i + e] : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = i + e < T;
        lp[e] = ok ? (float)ld4[e] : 0.f;
        sp[e] = ok ? precond_scale(lp[e], (float)(shift + c), (float)c) : 0.f;
      }
      *reinterpret_cast<float4*>(b.lam32 + oj + i) = lam;
      st_bf4(b.s16 + oj + i, sv);
    } else {
      sv = ld_bf4(b.s16 + oj + i);
    }
    *reinterpret_cast<float4*>(b.r32 + oj + i) = r;
    st_bf4(b.rsb + oj + i, make_float4(sv.x * r.x, sv.y * r.y, sv.z * r.z, sv.w * r.w));
  }
  if (threadIdx.x == 0) b.sigma[j] = sig;
  pdl_trigger();
}

// REG: the row fits one pass (T <= 4 * IR_THREADS), so z / (p, q) stay in registers and the kernel needs no
// dynamic shared memory -- its CTAs then co-reside with the persistent tcgen05 GEMM of the other row chunk.
template <bool REG>
__global__ void __launch_bounds__(IR_THREADS) ir_dir_kernel(IrBuf b, const int* __restrict__ n_dev, int T, int ld,
                                                            int first) {
  extern __shared__ float ir_sm[];       // z of this row (!REG)
  __shared__ double red[40];
  pdl_wait();
  const int j = blockIdx.x;
  if (n_dev && j >= *n_dev) return;
  const size_t o = (size_t)j * ld;
  double part = 0.0;
  float4 zreg = make_float4(0.f, 0.f, 0.f, 0.f);
  IR_ROW_LOOP(i, T) {
    const float4 acc = ld_acc4(b.acc32 + o + i, i, T);
    const float4 sv = ld_bf4(b.s16 + o + i);
    const float4 rs = ld_bf4(b.rsb + o + i);               // r.z = (s o r).acc, with the rounded s o r the GEMM saw
    const float4 z = make_float4(sv.x * acc.x, sv.y * acc.y, sv.z * acc.z, sv.w * acc.w);
    if (REG) zreg = z; else *reinterpret_cast<float4*>(ir_sm + i) = z;
    part += (double)rs.x * acc.x + (double)rs.y * acc.y + (double)rs.z * acc.z + (double)rs.w * acc.w;
  }
  float4 hreg = zreg, lreg = zreg;
  if (REG && !first) {
    const int i = 4 * threadIdx.x;
    if (i < T) { hreg = ld_bf4(b.p_hi + o + i); lreg = ld_bf4(b.p_lo + o + i); }
  }
  const double rz_new = block_sum(part, red);
  const double rz_old = first ? 0.0 : b.rz[j];
  const float beta = (first || !(rz_old > 0.0)) ? 0.f : (float)(rz_new / rz_old);
  IR_ROW_LOOP(i, T) {
    float4 p = REG ? zreg : *reinterpret_cast<const float4*>(ir_sm + i);
    if (!first) {
      const float4 h = REG ? hreg : ld_bf4(b.p_hi + o + i), l = REG ? lreg : ld_bf4(b.p_lo + o + i);
      p.x += beta * (h.x + l.x); p.y += beta * (h.y + l.y); p.z += beta * (h.z + l.z); p.w += beta * (h.w + l.w);
    }
    const float4 hi = make_float4(bf_round(p.x), bf_round(p.y), bf_round(p.z), bf_round(p.w));
    st_bf4(b.q_hi + o + i, hi);
    st_bf4(b.q_lo + o + i, make_float4(p.x - hi.x, p.y - hi.y, p.z - hi.z, p.w - hi.w));
  }
  if (threadIdx.x == 0) b.rz[j] = rz_new;
  pdl_trigger();
}

// One WARP per system (T <= 1024, T % 8 == 0): a lane keeps its 4 groups of 8 elements in registers, every load of
// the row is in flight at once (~14 KB per warp), bf16 arrays move 16 bytes per lane like the FP32 ones, and the row
// dot product needs warp shuffles only -- no __syncthreads.
constexpr int IR_WG = 4;
struct f8 { float4 a, b; };
__device__ __forceinline__ f8 ld_bf8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z));
  const float2 v3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
  f8 r;
  r.a = make_float4(v0.x, v0.y, v1.x, v1.y);
  r.b = make_float4(v2.x, v2.y, v3.x, v3.y);
  return r;
}
__device__ __forceinline__ void st_bf8(__nv_bfloat16* p, float4 a, float4 b) {
  const __nv_bfloat162 q0 = __floats2bfloat162_rn(a.x, a.y), q1 = __floats2bfloat162_rn(a.z, a.w);
  const __nv_bfloat162 q2 = __floats2bfloat162_rn(b.x, b.y), q3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&q0); u.y = *reinterpret_cast<const uint32_t*>(&q1);
  u.z = *reinterpret_cast<const uint32_t*>(&q2); u.w = *reinterpret_cast<const uint32_t*>(&q3);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 axpy4(float s, float4 a, float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}
__device__ __forceinline__ float4 bfr4(float4 a) { return make_float4(bf_round(a.x), bf_round(a.y), bf_round(a.z), bf_round(a.w)); }
// per-lane partial dot products stay in FP32 (a lane adds at most 32 products; FP32 -> FP64 conversions run at a
// fraction of the FP32 rate and showed up as the busiest pipe of these kernels); rows are reduced in FP64
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// MINB = resident CTAs per SM the register allocation is capped for: the kernel is latency-bound (ncu: 3.7 active
// warps per scheduler, 77 % of the stall cycles wait for loads), so more resident warps mean more loads in flight
template <int MINB>
__global__ void __launch_bounds__(IR_THREADS, MINB) ir_dir_warp_kernel(IrBuf b, const int* __restrict__ n_dev, int n, int T,
                                                                       int ld, int first, int zb16) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (IR_THREADS / 32) + (threadIdx.x >> 5);
  const int nn = n_dev ? min(*n_dev, n) : n;
  if (j >= nn) return;
  const size_t o = (size_t)j * ld;
  f8 z[IR_WG], pp[IR_WG];
  float part = 0.f;
#pragma unroll
  for (int g = 0; g < IR_WG; ++g) {
    const int i = 8 * (lane + 32 * g);
    z[g].a = z[g].b = pp[g].a = pp[g].b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < T) {                                              // T % 8 == 0: whole groups only
      float4 a0, a1;                                         // z' = the preconditioner GEMM's output
      if (zb16) {                                            // stored as bf16 rows (same ld) at the start of acc32
        const f8 av = ld_bf8(reinterpret_cast<const __nv_bfloat16*>(b.acc32) + o + i);
        a0 = av.a; a1 = av.b;
      } else {
        a0 = ld_f4(b.acc32 + o + i); a1 = ld_f4(b.acc32 + o + i + 4);
      }
      const f8 sv = ld_bf8(b.s16 + o + i);
      const f8 rs = ld_bf8(b.rsb + o + i);                   // r.z = (s o r).acc, with the rounded s o r the GEMM saw
      if (!first) {
        const f8 h = ld_bf8(b.p_hi + o + i), l = ld_bf8(b.p_lo + o + i);
        pp[g].a = add4(h.a, l.a); pp[g].b = add4(h.b, l.b);
      }
      z[g].a = mul4(sv.a, a0); z[g].b = mul4(sv.b, a1);
      part += dot4(rs.a, a0) + dot4(rs.b, a1);
    }
  }
  const double rz_new = warp_sum((double)part);
  const double rz_old = first ? 0.0 : b.rz[j];
  const float beta = (first || !(rz_old > 0.0)) ? 0.f : (float)(rz_new / rz_old);
#pragma unroll
  for (int g = 0; g < IR_WG; ++g) {
    const int i = 8 * (lane + 32 * g);
    if (i < T) {
      const float4 p0 = axpy4(beta, pp[g].a, z[g].a), p1 = axpy4(beta, pp[g].b, z[g].b);
      const float4 h0 = bfr4(p0), h1 = bfr4(p1);
      st_bf8(b.q_hi + o + i, h0, h1);
      st_bf8(b.q_lo + o + i, sub4(p0, h0), sub4(p1, h1));
    }
  }
  if (lane == 0) b.rz[j] = rz_new;
  pdl_trigger();
}

// The same update when the preconditioner GEMM's epilogue has already left (s o r).z' as per-tile partial dots: beta is
// known after eight loads and a warp sum, so the row is streamed once (z' 2, s 2, p 4 in; p 4 out: 12 B per element
// instead of 14) and nothing has to wait in registers for a row-wide reduction.
template <int MINB>
__global__ void __launch_bounds__(IR_THREADS, MINB) ir_dir_stream_kernel(IrBuf b, const int* __restrict__ n_dev, int n, int T,
                                                                         int ld, int first) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (IR_THREADS / 32) + (threadIdx.x >> 5);
  const int nn = n_dev ? min(*n_dev, n) : n;
  if (j >= nn) return;
  const size_t o = (size_t)j * ld;
  const int tiles = (T + 127) >> 7;
  const double rz_new = warp_sum(lane < tiles ? (double)b.rz_part[(size_t)j * b.rz_tiles + lane] : 0.0);
  const double rz_old = first ? 0.0 : b.rz[j];
  const float beta = (first || !(rz_old > 0.0)) ? 0.f : (float)(rz_new / rz_old);
  const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(b.acc32);
#pragma unroll
  for (int g = 0; g < IR_WG; ++g) {
    const int i = 8 * (lane + 32 * g);
    if (i < T) {                                                // T % 8 == 0: whole groups only
      const f8 av = ld_bf8(zb + o + i);
      const f8 sv = ld_bf8(b.s16 + o + i);
      float4 p0 = mul4(sv.a, av.a), p1 = mul4(sv.b, av.b);
      if (!first) {
        const f8 h = ld_bf8(b.p_hi + o + i), l = ld_bf8(b.p_lo + o + i);
        p0 = axpy4(beta, add4(h.a, l.a), p0);
        p1 = axpy4(beta, add4(h.b, l.b), p1);
      }
      const float4 h0 = bfr4(p0), h1 = bfr4(p1);
      st_bf8(b.q_hi + o + i, h0, h1);
      st_bf8(b.q_lo + o + i, sub4(p0, h0), sub4(p1, h1));
    }
  }
  if (lane == 0) b.rz[j] = rz_new;
  pdl_trigger();
}

template <bool REG>
__global__ void __launch_bounds__(IR_THREADS) ir_update_kernel(IrBuf b, const int* __restrict__ n_dev, int T, int ld,
                                                               int last, int first, float sc_num, float sc_c) {
  extern __shared__ float ir_sm[];       // p and q of this row (!REG)
  __shared__ double red[40];
  float* sp = ir_sm;
  float* sq = ir_sm + ((T + 3) & ~3);
  pdl_wait();
  const int j = blockIdx.x;
  if (n_dev && j >= *n_dev) return;
  const size_t o = (size_t)j * ld;
  float part = 0.f;
  float4 preg = make_float4(0.f, 0.f, 0.f, 0.f), qreg = preg, lreg4 = preg;
  IR_ROW_LOOP(i, T) {
    const float4 h = ld_bf4(b.p_hi + o + i), l = ld_bf4(b.p_lo + o + i);
    const float4 acc = ld_acc4(b.acc32 + o + i, i, T);
    const float4 lam = ld_f4(b.lam32 + o + i);
    if (REG) lreg4 = lam;
    const float4 p = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
    const float4 q = make_float4(fmaf(lam.x, p.x, acc.x), fmaf(lam.y, p.y, acc.y), fmaf(lam.z, p.z, acc.z),
                                 fmaf(lam.w, p.w, acc.w));
    if (REG) { preg = p; qreg = q; } else {
      *reinterpret_cast<float4*>(sp + i) = p;
      *reinterpret_cast<float4*>(sq + i) = q;
    }
    part += fmaf(p.x, q.x, fmaf(p.y, q.y, fmaf(p.z, q.z, p.w * q.w)));
  }
  // REG: issue the loads of the second phase before the row-wide reduction so their latency hides behind it
  float4 dreg = make_float4(0.f, 0.f, 0.f, 0.f), rreg = preg;
  if (REG) {
    const int i = 4 * threadIdx.x;
    if (i < T) {
      if (!first) dreg = ld_f4(b.d32 + o + i);             // first step of a correction: d = alpha p
      if (!last) rreg = ld_f4(b.r32 + o + i);
    }
  }
  const double pq = block_sum((double)part, red);
  const float alpha = (pq > 0.0) ? (float)(b.rz[j] / pq) : 0.f;
  IR_ROW_LOOP(i, T) {
    const float4 p = REG ? preg : *reinterpret_cast<const float4*>(sp + i);
    const float4 q = REG ? qreg : *reinterpret_cast<const float4*>(sq + i);
    float4 d = REG ? dreg : (first ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_f4(b.d32 + o + i));
    d.x = fmaf(alpha, p.x, d.x); d.y = fmaf(alpha, p.y, d.y); d.z = fmaf(alpha, p.z, d.z); d.w = fmaf(alpha, p.w, d.w);
    *reinterpret_cast<float4*>(b.d32 + o + i) = d;
    if (!last) {
      float4 r = REG ? rreg : ld_f4(b.r32 + o + i);
      r.x = fmaf(-alpha, q.x, r.x); r.y = fmaf(-alpha, q.y, r.y); r.z = fmaf(-alpha, q.z, r.z); r.w = fmaf(-alpha, q.w, r.w);
      *reinterpret_cast<float4*>(b.r32 + o + i) = r;
      const float4 lam = REG ? lreg4 : ld_f4(b.lam32 + o + i);
      const float4 sv = make_float4(precond_scale(lam.x, sc_num, sc_c), precond_scale(lam.y, sc_num, sc_c),
                                    precond_scale(lam.z, sc_num, sc_c), precond_scale(lam.w, sc_num, sc_c));
      st_bf4(b.rsb + o + i, make_float4(sv.x * r.x, sv.y * r.y, sv.z * r.z, sv.w * r.w));
    }
  }
  pdl_trigger();
}

__global__ void __launch_bounds__(IR_THREADS)
ir_finish_kernel(IrBuf b, double* __restrict__ P, double* __restrict__ X, const int* __restrict__ slot_state,
                 const int* __restrict__ act, const int* __restrict__ n_dev, int T, int ld) {
  pdl_wait();
  const int j = blockIdx.x;
  if (n_dev && j >= *n_dev) return;
  const int slot = act[j];
  const double sig = b.sigma[j];
  const float* d = b.d32 + (size_t)j * ld;
  double* p = P + (size_t)slot * ld;
  double* x = X + (size_t)slot_state[slot] * ld;
  if ((ld & 3) == 0) {
    // 16-byte accesses for whole groups of four; a partial last group (T % 4 != 0) goes element by element
    IR_ROW_LOOP(i, T) {
      if (i + 4 > T) {
        for (int t = i; t < T; ++t) {
          const double v = p[t] + sig * (double)d[t];
          p[t] = v;
          x[t] = v;
        }
        continue;
      }
      const float4 dv = ld_f4(d + i);
      double2 a = *reinterpret_cast<const double2*>(p + i), c = *reinterpret_cast<const double2*>(p + i + 2);
      a.x += sig * (double)dv.x; a.y += sig * (double)dv.y;
      c.x += sig * (double)dv.z; c.y += sig * (double)dv.w;
      *reinterpret_cast<double2*>(p + i) = a;
      *reinterpret_cast<double2*>(p + i + 2) = c;
      *reinterpret_cast<double2*>(x + i) = a;
      *reinterpret_cast<double2*>(x + i + 2) = c;
    }
  } else {
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const double v = p[t] + sig * (double)d[t];
      p[t] = v;
      x[t] = v;
    }
  }
  pdl_trigger();
}

int ir_inner_init(const IrBuf& b, const double* R, const double* Lambda, const double* rr, const int* act, int n,
                  const int* n_dev, int T, int ld, double shift, double c, cudaStream_t s) {
  if (n <= 0) return VGM_OK;
  ProfScope prof(PROF_VEC, 0.0, (double)n * T * 16.0, s);   // R in, r32 + rsb out, s16 in (the common case: list unchanged)
  VGM_CUDA_OK(launch_pdl(ir_inner_init_kernel, dim3(n), dim3(IR_THREADS), 0, s, b, R, Lambda, rr, act, n_dev, T, ld, shift, c));
  return VGM_OK;
}

static int ir_dir_variant() {
  // default 3: register allocation capped for 3 resident CTAs per SM (85 registers, 24 bytes of spills): measured
  // 5.65 ms per step at 0.86 of the HBM copy bandwidth against 6.09 ms / 0.80 with 2 CTAs (92 registers) and
  // 5.76 ms with 4 (176 bytes of spills); 2 / 4 select those, 1 the CTA-per-row kernel (6.6 ms)
  static const int variant = getenv("VGM_IR_DIR_VARIANT") ? atoi(getenv("VGM_IR_DIR_VARIANT")) : 3;
  return variant;
}

// true when ir_dir runs the warp-per-row kernel, which can take the preconditioner product as bf16
bool ir_dir_takes_bf16(int T) {
  static const bool off = getenv("VGM_Z_FP32") != nullptr;
  return !off && ir_dir_variant() != 1 && T <= 8 * 32 * IR_WG && (T % 8) == 0;
}

// VGM_NO_GEMM_DOTS=1: ir_dir computes (s o r).z' itself (one more 2-byte read per element and a two-phase kernel)
bool ir_dir_takes_dots() {
  static const bool off = getenv("VGM_NO_GEMM_DOTS") != nullptr;
  return !off;
}

int ir_dir(const IrBuf& b, int n, const int* n_dev, int T, int ld, int first, int zb16, cudaStream_t s) {
  if (n <= 0) return VGM_OK;
  VGM_REQUIRE(!zb16 || ir_dir_takes_bf16(T), "ir_dir: bf16 accumulators need the warp-per-row kernel");
  if (zb16 == 2) {
    // reads z' 2, s16 2, (p_hi, p_lo 4); writes p_hi, p_lo 4
    ProfScope prof(PROF_IR_DIR, 0.0, (double)n * T * (first ? 8.0 : 12.0), s);
    const dim3 grid(ceil_div(n, IR_THREADS / 32));
    VGM_CUDA_OK(launch_pdl(ir_dir_stream_kernel<4>, grid, dim3(IR_THREADS), 0, s, b, n_dev, n, T, ld, first));
    return VGM_OK;
  }
  // reads acc 4 (2 as bf16), s16 2, rsb 2, (p_hi, p_lo 4); writes p_hi, p_lo 4
  ProfScope prof(PROF_IR_DIR, 0.0, (double)n * T * ((first ? 12.0 : 16.0) - (zb16 ? 2.0 : 0.0)), s);
  // VGM_IR_DIR_VARIANT=1: one CTA per row with the row in registers (like ir_update) instead of one warp per row
  const int variant = ir_dir_variant();
  if (variant == 1 && T <= 4 * IR_THREADS)
    VGM_CUDA_OK(launch_pdl(ir_dir_kernel<true>, dim3(n), dim3(IR_THREADS), 0, s, b, n_dev, T, ld, first));
  else if (T <= 8 * 32 * IR_WG && (T % 8) == 0) {
    const dim3 grid(ceil_div(n, IR_THREADS / 32));
    if (variant == 2) VGM_CUDA_OK(launch_pdl(ir_dir_warp_kernel<2>, grid, dim3(IR_THREADS), 0, s, b, n_dev, n, T, ld, first, zb16));
    else if (variant == 4) VGM_CUDA_OK(launch_pdl(ir_dir_warp_kernel<4>, grid, dim3(IR_THREADS), 0, s, b, n_dev, n, T, ld, first, zb16));
    else VGM_CUDA_OK(launch_pdl(ir_dir_warp_kernel<3>, grid, dim3(IR_THREADS), 0, s, b, n_dev, n, T, ld, first, zb16));
  }
  else if (T <= 4 * IR_THREADS)
    VGM_CUDA_OK(launch_pdl(ir_dir_kernel<true>, dim3(n), dim3(IR_THREADS), 0, s, b, n_dev, T, ld, first));
  else
    VGM_CUDA_OK(launch_pdl(ir_dir_kernel<false>, dim3(n), dim3(IR_THREADS), ((T + 3) & ~3) * sizeof(float), s, b, n_dev, T, ld, first));
  return VGM_OK;
}

int ir_update(const IrBuf& b, int n, const int* n_dev, int T, int ld, int last, int first, double shift, double c,
              cudaStream_t s) {
  if (n <= 0) return VGM_OK;
  // reads p_hi, p_lo 4, acc32 4, lam32 4, d32 4, (r32 4); writes d32 4, (r32 4, rsb 2)
  ProfScope prof(PROF_IR_UPDATE, 0.0, (double)n * T * ((last ? 20.0 : 30.0) - (first ? 4.0 : 0.0)), s);
  const float sc_num = (float)(shift + c), sc_c = (float)c;
  if (T <= 4 * IR_THREADS)
    VGM_CUDA_OK(launch_pdl(ir_update_kernel<true>, dim3(n), dim3(IR_THREADS), 0, s, b, n_dev, T, ld, last, first, sc_num, sc_c));
  else
    VGM_CUDA_OK(launch_pdl(ir_update_kernel<false>, dim3(n), dim3(IR_THREADS), 2 * ((T + 3) & ~3) * sizeof(float), s, b, n_dev, T, ld,
                           last, first, sc_num, sc_c));
  return VGM_OK;
}

int ir_finish(const IrBuf& b, double* P, double* X, const int* slot_state, const int* act, int n, const int* n_dev,
              int T, int ld, cudaStream_t s) {
  if (n <= 0) return VGM_OK;
  ProfScope prof(PROF_VEC, 0.0, (double)n * T * 28.0, s);
  VGM_CUDA_OK(launch_pdl(ir_finish_kernel, dim3(n), dim3(IR_THREADS), 0, s, b, P, X, slot_state, act, n_dev, T, ld));
  return VGM_OK;
}

// ---- band-width probe ---------------------------------------------------------------------------------
__global__ void band_absmax_kernel(const double* __restrict__ B, int rows, int cols, int ld, double* __restrict__ out) {
  __shared__ double red[40];
  double m = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * cols; i += (size_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs(B[(i / cols) * ld + (i % cols)]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
  }
}
__global__ void band_width_kernel(const double* __restrict__ B, int rows, int cols, int ld, double rel,
                                  const double* __restrict__ absmax, int* __restrict__ band) {
  const double thr = rel * absmax[0];
  int w = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * cols; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const double v = fabs(B[(size_t)r * ld + c]);
    if (v > thr || v != v) w = max(w, abs(r - c));
  }
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if ((threadIdx.x & 31) == 0 && w > 0) atomicMax(band, w);
}

}  // namespace vgm

extern "C" int vgm_band_halfwidth(const double* B, int rows, int cols, int ld, double rel_threshold, int* band_host,
                                  vgm_stream_t stream) {
  VGM_REQUIRE(B && band_host && rows > 0 && cols > 0 && ld >= cols && rel_threshold >= 0.0, "band arguments");
  cudaStream_t s = (cudaStream_t)stream;
  void* scratch = nullptr;
  VGM_CUDA_OK(cudaMallocAsync(&scratch, 16, s));
  VGM_CUDA_OK(cudaMemsetAsync(scratch, 0, 16, s));
  double* absmax = (double*)scratch;
  int* band = (int*)((char*)scratch + 8);
  const int blocks = 296;
  vgm::band_absmax_kernel<<<blocks, 256, 0, s>>>(B, rows, cols, ld, absmax);
  vgm::band_width_kernel<<<blocks, 256, 0, s>>>(B, rows, cols, ld, rel_threshold, absmax, band);
  VGM_LAUNCH_OK();
  VGM_CUDA_OK(cudaMemcpyAsync(band_host, band, sizeof(int), cudaMemcpyDeviceToHost, s));
  VGM_CUDA_OK(cudaFreeAsync(scratch, s));
  VGM_CUDA_OK(cudaStreamSynchronize(s));
  return VGM_OK;
}
