// k5_s4_conv.cu -- S4 layer call in CNN mode: DPLR convolution kernel and the causal convolution it drives.
//
// Reference operators (JAX/Flax, models/s4.py):
//   kernel_DPLR(Lambda, P, Q, B, C, step, L)      :50-69   truncated generating function at the L roots of unity (4 Cauchy sums + Woodbury), ifft, real part
//   causal_convolution(u, K) + D * u               :72-79, :169-173   (FFT based in the reference)
//   S4 = vmap(S4Layer) over the feature axis       :182-188
// Three kernels, no FFT library:
//   s4_at_roots_kernel   one thread per (feature, frequency), float64 complex.  The reference's g = (2/step)(1-W)/(1+W), c = 2/(1+W) are singular at the
//                        Nyquist root W = -1; with r_xy = sum_n v_xy[n] / ((2/step)(1-W) - Lambda_n (1+W)) the same quantity is
//                        2 (r00 - (1+W) r01 r10 / (1 + (1+W) r11)), regular everywhere.
//   s4_idft_real_kernel  K[t] = Re(1/L sum_l a_l e^{+2 pi i l t / L}), direct O(L^2) per feature in float64 with an exact twiddle table in shared memory
//                        (parameter-sized work: H L^2 = 5e8 complex MACs at H = 128, L = 2048), written transposed (L,H) for the convolution.
//   s4_causal_conv_kernel  y[b,t,h] = sum_{s<=t} K[t-s,h] u[b,s,h] + D[h] u[b,t,h], direct form, fp32: CTA = (64 outputs) x (32 features) of one sequence,
//                        64-token source tiles in shared memory, every thread 8 consecutive outputs with a 15-lag register window per 8 source tokens
//                        (64 FMA per 23 shared-memory reads).  B H T^2 / 2 FMA in total -- FMA-pipe bound.
#include "common.cuh"

namespace eigb200 {

struct dcomplex { double re, im; };
__device__ __forceinline__ dcomplex cmul(dcomplex a, dcomplex b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ dcomplex cadd(dcomplex a, dcomplex b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ dcomplex csub(dcomplex a, dcomplex b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ dcomplex cdiv(dcomplex a, dcomplex b) {
  const double d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
__device__ __forceinline__ dcomplex cconj(dcomplex a) { return {a.re, -a.im}; }

// parameters (H,N) complex64 interleaved; at (H,L) complex128 out
__global__ void __launch_bounds__(128) s4_at_roots_kernel(const float2* __restrict__ Lam, const float2* __restrict__ P, const float2* __restrict__ Q,
                                                          const float2* __restrict__ Bv, const float2* __restrict__ Cv, const float* __restrict__ step,
                                                          int H, int N, int L, dcomplex* __restrict__ at) {
  extern __shared__ double sh[];                                   // [5][N] complex: Lambda, C* B, C* P, Q* B, Q* P
  dcomplex* sl = reinterpret_cast<dcomplex*>(sh);
  const int h = blockIdx.y;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float2 l = Lam[(size_t)h * N + n], p = P[(size_t)h * N + n], q = Q[(size_t)h * N + n], b = Bv[(size_t)h * N + n], c = Cv[(size_t)h * N + n];
    const dcomplex cc = cconj({(double)c.x, (double)c.y}), qc = cconj({(double)q.x, (double)q.y});
    const dcomplex bb{(double)b.x, (double)b.y}, pp{(double)p.x, (double)p.y};
    sl[n] = {(double)l.x, (double)l.y};
    sl[N + n] = cmul(cc, bb); sl[2 * N + n] = cmul(cc, pp); sl[3 * N + n] = cmul(qc, bb); sl[4 * N + n] = cmul(qc, pp);
  }
  __syncthreads();
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double sn, cs;
  sincospi(-2.0 * (double)l / (double)L, &sn, &cs);                // W = exp(-2 pi i l / L)
  const dcomplex onem{1.0 - cs, -sn}, onep{1.0 + cs, sn};
  const double ts = 2.0 / (double)step[h];
  dcomplex r00{0, 0}, r01{0, 0}, r10{0, 0}, r11{0, 0};
  for (int n = 0; n < N; ++n) {
    const dcomplex den = csub({ts * onem.re, ts * onem.im}, cmul(sl[n], onep));
    const dcomplex inv = cdiv({1.0, 0.0}, den);
    r00 = cadd(r00, cmul(sl[N + n], inv)); r01 = cadd(r01, cmul(sl[2 * N + n], inv));
    r10 = cadd(r10, cmul(sl[3 * N + n], inv)); r11 = cadd(r11, cmul(sl[4 * N + n], inv));
  }
  const dcomplex corr = cdiv(cmul(cmul(onep, r01), r10), cadd({1.0, 0.0}, cmul(onep, r11)));
  const dcomplex v = csub(r00, corr);
  at[(size_t)h * L + l] = {2.0 * v.re, 2.0 * v.im};
}

// K_t[t,h] = Re(1/L sum_l at[h,l] e^{+2 pi i l t / L}); CTA = 256 outputs t of one feature
__global__ void __launch_bounds__(256) s4_idft_real_kernel(const dcomplex* __restrict__ at, int H, int L, float* __restrict__ Kt) {
  extern __shared__ double sh[];                                   // twiddles (cos, sin)(2 pi j / L), j < L
  double* tc = sh; double* tsn = sh + L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) sincospi(2.0 * (double)j / (double)L, &tsn[j], &tc[j]);
  __syncthreads();
  const int h = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const dcomplex* a = at + (size_t)h * L;
  double acc = 0.0;
  int j = 0;                                                       // (l * t) mod L, advanced incrementally
  for (int l = 0; l < L; ++l) {
    const dcomplex v = a[l];
    acc += v.re * tc[j] - v.im * tsn[j];
    j += t; if (j >= L) j -= L;
  }
  Kt[(size_t)t * H + h] = (float)(acc / (double)L);
}

constexpr int CV_TT = 64, CV_H = 32;

__global__ void __launch_bounds__(256) s4_causal_conv_kernel(const float* __restrict__ u, const float* __restrict__ Kt, const float* __restrict__ Dv,
                                                             float* __restrict__ y, int64_t T, int H) {
  __shared__ float us[CV_TT][CV_H];
  __shared__ float ks[2 * CV_TT][CV_H];                            // lag index j = lag - lagmin, lagmin = T0 - S0 - 63; row 127 unused
  const int hl = threadIdx.x & 31, tq = threadIdx.x >> 5;
  const int h0 = blockIdx.y * CV_H, b = blockIdx.z;
  const int64_t T0 = (int64_t)blockIdx.x * CV_TT;
  const int h = h0 + hl;
  const bool hok = h < H;
  const float* ub = u + (size_t)b * T * H;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int64_t S0 = 0; S0 <= T0; S0 += CV_TT) {
    __syncthreads();
    for (int i = threadIdx.x; i < CV_TT * CV_H; i += 256) {
      const int s = i >> 5, hh = i & 31;
      const int64_t sg = S0 + s;
      us[s][hh] = (sg < T && h0 + hh < H) ? __ldg(ub + sg * H + h0 + hh) : 0.f;
    }
    const int64_t lagmin = T0 - S0 - (CV_TT - 1);
    for (int i = threadIdx.x; i < (2 * CV_TT - 1) * CV_H; i += 256) {
      const int j = i >> 5, hh = i & 31;
      const int64_t lag = lagmin + j;
      ks[j][hh] = (lag >= 0 && lag < T && h0 + hh < H) ? __ldg(Kt + lag * H + h0 + hh) : 0.f;   // negative lags (future tokens) contribute 0
    }
    __syncthreads();
#pragma unroll 1
    for (int s0 = 0; s0 < CV_TT; s0 += 8) {
      // outputs t = tq*8 + i, sources s = s0 + jj: lag index = tq*8 + i - s0 - jj + 63, i - jj in [-7, 7]
      const int base = tq * 8 - s0 + (CV_TT - 1) - 7;
      float win[15], uv[8];
#pragma unroll
      for (int k = 0; k < 15; ++k) win[k] = ks[base + k][hl];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) uv[jj] = us[s0 + jj][hl];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i] = fmaf(win[i - jj + 7], uv[jj], acc[i]);
    }
  }
  if (hok) {
    const float d = Dv ? __ldg(Dv + h) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t t = T0 + tq * 8 + i;
      if (t < T) y[((size_t)b * T + t) * H + h] = fmaf(d, __ldg(ub + t * H + h), acc[i]);
    }
  }
}

}  // namespace eigb200

using namespace eigb200;

extern "C" size_t eigb200_s4_kernel_workspace_bytes(int H, int L) { return (size_t)H * L * sizeof(dcomplex); }

extern "C" int eigb200_s4_kernel(void* stream, const float* d_Lambda, const float* d_P, const float* d_Q, const float* d_B, const float* d_C,
                                 const float* d_step, int H, int N, int L, float* d_Kt, void* d_workspace, size_t workspace_bytes) {
  EIGB_CHECK_ARG(d_Lambda && d_P && d_Q && d_B && d_C && d_step && d_Kt && d_workspace, "s4_kernel: null pointer");
  EIGB_CHECK_ARG(H > 0 && H <= 65535 && N > 0 && N <= 1024 && L > 0 && L <= 8192, "s4_kernel: bad shape H=%d N=%d L=%d (N <= 1024, L <= 8192)", H, N, L);
  EIGB_CHECK_ARG(workspace_bytes >= eigb200_s4_kernel_workspace_bytes(H, L), "s4_kernel: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  dcomplex* at = reinterpret_cast<dcomplex*>(d_workspace);
  {
    dim3 grid((L + 127) / 128, H);
    const size_t smem = (size_t)5 * N * sizeof(dcomplex);
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(s4_at_roots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    s4_at_roots_kernel<<<grid, 128, smem, st>>>(reinterpret_cast<const float2*>(d_Lambda), reinterpret_cast<const float2*>(d_P),
                                                reinterpret_cast<const float2*>(d_Q), reinterpret_cast<const float2*>(d_B),
                                                reinterpret_cast<const float2*>(d_C), d_step, H, N, L, at);
    EIGB_LAUNCH_CHECK("s4_at_roots_kernel");
  }
  {
    dim3 grid((L + 255) / 256, H);
    const size_t smem = (size_t)2 * L * sizeof(double);
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(s4_idft_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    s4_idft_real_kernel<<<grid, 256, smem, st>>>(at, H, L, d_Kt);
    EIGB_LAUNCH_CHECK("s4_idft_real_kernel");
  }
  return EIGB200_OK;
}

extern "C" int eigb200_s4_causal_conv(void* stream, const float* d_u, const float* d_Kt, const float* d_D, float* d_y, int64_t B, int64_t T, int H) {
  EIGB_CHECK_ARG(d_u && d_Kt && d_y, "s4_causal_conv: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0, "s4_causal_conv: bad shape");
  dim3 grid((unsigned)((T + CV_TT - 1) / CV_TT), (H + CV_H - 1) / CV_H, (unsigned)B);
  s4_causal_conv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_u, d_Kt, d_D, d_y, T, H);
  EIGB_LAUNCH_CHECK("s4_causal_conv_kernel");
  return EIGB200_OK;
}
