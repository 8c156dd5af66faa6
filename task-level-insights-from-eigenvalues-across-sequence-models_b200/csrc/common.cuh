// common.cuh -- shared device/host helpers for libeigb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/eigb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libeigb200 is written for sm_100a (B200) only"
#endif

namespace eigb200 {

// ---- error plumbing -------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
int  num_sms();

#define EIGB_CHECK_ARG(cond, ...)                                  \
  do { if (!(cond)) { ::eigb200::set_error(__VA_ARGS__); return EIGB200_EINVAL; } } while (0)
#define EIGB_CUDA(call)                                            \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::eigb200::cuda_fail(e__, #call); } while (0)
#define EIGB_LAUNCH_CHECK(name)                                    \
  do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return ::eigb200::cuda_fail(e__, name); } while (0)

// ---- threshold edges ------------------------------------------------------------------------------------------
// nb = nthr + 1 bins.  Bins 0..nb-2 are closed intervals [lo, hi]; bin nb-1 is v > gt.
// For float32 values the edges are chosen on the host so that a float32 compare reproduces either the float64
// compare NumPy >= 2 performs (lo = ceil32(thr), hi = floor32(thr)) or the float32 compare of NumPy 1.24.
struct EdgesF { float lo[7]; float hi[7]; float gt; int nb; };
struct EdgesD { double lo[7]; double hi[7]; double gt; int nb; };
int make_edges_f(const double* thr, int nthr, int compare_mode, EdgesF* e);
int make_edges_d(const double* thr, int nthr, EdgesD* e);

// ---- device helpers -------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Packed fp32 pairs (sm_100: fma / mul / add .f32x2 = FFMA2 / FMUL2 / FADD2, one issue slot for two IEEE fp32 operations, each half rounded exactly as the
// scalar instruction).  A pair lives in an aligned 64-bit register; pk2(a, a) becomes the instruction's scalar-broadcast operand.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// torch.nn.functional.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }
// torch.nn.functional.elu(alpha=1)
__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : expm1f(z); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }
__device__ __forceinline__ float silu_f(float z) { return z / (1.f + expf(-z)); }
// nn.GELU() exact form
__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f)); }

// Branch-free variants for the GEMM epilogues, where 16k activations per tile share the SM's issue slots with the pipeline:
//   erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32 rounding level for 1 + erf), exp by ex2.approx, 1/x by rcp.approx.
// ~16 instructions per GELU instead of ~45 for the erff() call with its divergent branch.
__device__ __forceinline__ float fast_exp_f(float z) {            // e^z, relative error ~2 ulp for |z| < 80
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float fast_rcp_f(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  return r;
}
// GELU(z) = z Phi(z) = max(z, 0) - (|z| / 2) erfc(|z| / sqrt 2), erfc(x) = P(t) t exp(-x^2), t = 1 / (1 + p x)  (Abramowitz-Stegun 7.1.26).  Written on |z|
// with the 1 / sqrt 2 and 1 / 2 factors folded into the constants: 13 instructions (2 MUFU), no compare / select.  S scales the result (S = 1: plain GELU;
// the fused out_proj -> GLU kernel asks for S_a GELU(z) directly).
__device__ __forceinline__ float gelu_fast_scaled_f(float z, float zS, float S) {   // zS = S z (formed by the caller's bias FMA when S != 1)
  const float az = fabsf(z);
  const float t = fast_rcp_f(fmaf(0.3275911f * 0.70710678118654752440f, az, 1.f));
  float p = fmaf(1.061405429f * 0.5f, t, -1.453152027f * 0.5f);                      // P(t) / 2
  p = fmaf(p, t, 1.421413741f * 0.5f);
  p = fmaf(p, t, -0.284496736f * 0.5f);
  p = fmaf(p, t, 0.254829592f * 0.5f);
  const float xs = az * 0.84932180028801904272f;                                    // sqrt(log2(e) / 2): xs^2 = (z^2 / 2) log2 e
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(xs * xs)));
  const float w = (az * t) * e;
  return fmaf(-(p * S), w, fmaxf(zS, 0.f));
}
__device__ __forceinline__ float gelu_fast_f(float z) {
  const float az = fabsf(z);
  const float t = fast_rcp_f(fmaf(0.3275911f * 0.70710678118654752440f, az, 1.f));
  float p = fmaf(1.061405429f * 0.5f, t, -1.453152027f * 0.5f);
  p = fmaf(p, t, 1.421413741f * 0.5f);
  p = fmaf(p, t, -0.284496736f * 0.5f);
  p = fmaf(p, t, 0.254829592f * 0.5f);
  const float xs = az * 0.84932180028801904272f;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(xs * xs)));
  return fmaf(-p, (az * t) * e, fmaxf(z, 0.f));
}
__device__ __forceinline__ float sigmoid_fast_f(float z) { return fast_rcp_f(1.f + fast_exp_f(-z)); }

// Per-thread bin bookkeeping.  c[0..6]: closed bins (unused ones have lo=+inf, hi=-inf so they never fire),
// c[7]: the open last bin (v > gt), c[8]: phase-bin-0 count.  flush_bins() maps them onto the 8 output slots.
#define EIGB_NCNT 9
__device__ __forceinline__ void bin_f32(float v, const EdgesF& e, int (&c)[EIGB_NCNT]) {
#pragma unroll
  for (int j = 0; j < 7; ++j) c[j] += (v >= e.lo[j] && v <= e.hi[j]) ? 1 : 0;
  c[7] += (v > e.gt) ? 1 : 0;
}
__device__ __forceinline__ void bin_f64(double v, const EdgesD& e, int (&c)[EIGB_NCNT]) {
#pragma unroll
  for (int j = 0; j < 7; ++j) c[j] += (v >= e.lo[j] && v <= e.hi[j]) ? 1 : 0;
  c[7] += (v > e.gt) ? 1 : 0;
}
// slot index of per-thread counter j for a histogram with nb bins
__device__ __forceinline__ int slot_of(int j, int nb) { return j < 7 ? j : (j == 7 ? nb - 1 : 7); }

// Butterfly "transpose-reduce": every lane holds N partial sums v[0..N); on return v[0] is the full 32-lane sum of
// partial index (lane >> (5 - log2 N)).  N*(1 - 1/N) + (5 - log2 N) shuffles instead of 5*N.
template <int N>
__device__ __forceinline__ float transpose_reduce(float (&v)[N], int lane) {
  static_assert(N >= 1 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two <= 32");
  int off = 16;
#pragma unroll
  for (int m = N / 2; m >= 1; m >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const float send = up ? v[i] : v[i + m];
      const float keep = up ? v[i + m] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}
template <int N> struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };
#endif  // __CUDACC__

}  // namespace eigb200
