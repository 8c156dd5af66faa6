// k6_hist.cu -- on-device log-spaced histograms and quantiles of the eigenvalue radii per (layer, head[, state]) (BASELINE north star: "on-device log-spaced
// histograms/quantiles per layer/head/state").  The reference only has the 7 + 6 fixed threshold bins of threshold_analysis (analysis/eval_eig.py:335-391,
// served by K1 / eigb200_ratio_hist); this is the finer view of the same arrays without moving them to the host: a (B,N,inner) array -> per `inner` column a
// histogram over [lo, hi) with nbins log-spaced bins (+ underflow / overflow / NaN slots), int64, accumulated (so several batches or GPUs can be summed:
// integer counts, order independent), and quantiles read off the cumulative counts with log-linear interpolation inside a bin
// (relative resolution (hi / lo)^(1 / nbins) - 1: 4.6 % with 512 bins over 10 decades).
#include "common.cuh"

namespace eigb200 {

constexpr int HI_THREADS = 256;
constexpr int HI_SMEM_INTS = 12 * 1024;                            // 48 KB of block-private counters

// slots per column: [0] v < lo (incl. v <= 0), [1 .. nbins] the log-spaced bins, [nbins + 1] v >= hi, [nbins + 2] NaN
template <typename T>
__global__ void __launch_bounds__(HI_THREADS) log_hist_kernel(const T* __restrict__ v, int64_t rows, int64_t inner, int i0, int itile, float log2_lo, float inv_w,
                                                              int nbins, unsigned long long* __restrict__ hist) {
  extern __shared__ int sh[];
  const int nslot = nbins + 3;
  for (int i = threadIdx.x; i < itile * nslot; i += HI_THREADS) sh[i] = 0;
  __syncthreads();
  const int64_t total = rows * itile;
  for (int64_t e = (int64_t)blockIdx.x * HI_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * HI_THREADS) {
    const int64_t r = e / itile;
    const int c = (int)(e - r * itile);
    const float x = (float)v[r * inner + i0 + c];
    int slot;
    if (x != x) slot = nbins + 2;
    else if (!(x > 0.f)) slot = 0;
    else {
      const float t = (__log2f(x) - log2_lo) * inv_w;              // bin coordinate
      slot = t < 0.f ? 0 : (t >= (float)nbins ? nbins + 1 : 1 + (int)t);
    }
    atomicAdd(&sh[c * nslot + slot], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < itile * nslot; i += HI_THREADS) {
    const int cnt = sh[i];
    if (cnt) atomicAdd(hist + (size_t)(i0 + i / nslot) * nslot + (i % nslot), (unsigned long long)cnt);
  }
}

// one thread per (column, quantile): walk the cumulative counts of the finite values; interpolate log-linearly inside the bin
__global__ void hist_quantiles_kernel(const unsigned long long* __restrict__ hist, int64_t inner, int nbins, double log2_lo, double w, const double* __restrict__ qs, int nq,
                                      double lo, double hi, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= inner * nq) return;
  const int64_t c = idx / nq;
  const int qi = (int)(idx - c * nq);
  const unsigned long long* h = hist + (size_t)c * (nbins + 3);
  unsigned long long n = 0;
  for (int s = 0; s < nbins + 2; ++s) n += h[s];
  if (n == 0) { out[idx] = nan(""); return; }
  const double target = qs[qi] * (double)n;
  unsigned long long cum = 0;
  for (int s = 0; s < nbins + 2; ++s) {
    const unsigned long long cnt = h[s];
    if (cnt && (double)(cum + cnt) >= target) {
      if (s == 0) { out[idx] = lo; return; }                        // below the range: report its edge
      if (s == nbins + 1) { out[idx] = hi; return; }
      const double frac = (target - (double)cum) / (double)cnt;
      out[idx] = exp2(log2_lo + ((double)(s - 1) + fmin(fmax(frac, 0.0), 1.0)) * w);
      return;
    }
    cum += cnt;
  }
  out[idx] = hi;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_log_hist(void* stream, const void* d_values, int dtype, int64_t B, int64_t N, int64_t inner, double lo, double hi, int nbins, int64_t* d_hist) {
  EIGB_CHECK_ARG(d_values && d_hist, "log_hist: null pointer");
  EIGB_CHECK_ARG(dtype == EIGB200_F32 || dtype == EIGB200_F64, "log_hist: dtype must be F32 or F64");
  EIGB_CHECK_ARG(B > 0 && N > 0 && inner > 0, "log_hist: bad shape");
  EIGB_CHECK_ARG(lo > 0.0 && hi > lo && nbins >= 1 && nbins + 3 <= HI_SMEM_INTS, "log_hist: need 0 < lo < hi and 1 <= nbins <= %d", HI_SMEM_INTS - 3);
  const int nslot = nbins + 3;
  const int itile_max = HI_SMEM_INTS / nslot;
  const float log2_lo = (float)log2(lo);
  const float inv_w = (float)((double)nbins / (log2(hi) - log2(lo)));
  const int64_t rows = B * N;
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t i0 = 0; i0 < inner; i0 += itile_max) {
    const int itile = (int)((inner - i0) < itile_max ? (inner - i0) : itile_max);
    const int64_t total = rows * itile;
    int64_t blocks = (total + HI_THREADS * 8 - 1) / (HI_THREADS * 8);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)itile * nslot * sizeof(int);
    if (dtype == EIGB200_F32)
      log_hist_kernel<float><<<(unsigned)blocks, HI_THREADS, smem, st>>>((const float*)d_values, rows, inner, (int)i0, itile, log2_lo, inv_w, nbins, (unsigned long long*)d_hist);
    else
      log_hist_kernel<double><<<(unsigned)blocks, HI_THREADS, smem, st>>>((const double*)d_values, rows, inner, (int)i0, itile, log2_lo, inv_w, nbins, (unsigned long long*)d_hist);
    EIGB_LAUNCH_CHECK("log_hist_kernel");
  }
  return EIGB200_OK;
}

extern "C" int eigb200_hist_quantiles(void* stream, const int64_t* d_hist, int64_t inner, double lo, double hi, int nbins, const double* d_q, int nq, double* d_out) {
  EIGB_CHECK_ARG(d_hist && d_q && d_out, "hist_quantiles: null pointer");
  EIGB_CHECK_ARG(inner > 0 && nq > 0 && lo > 0.0 && hi > lo && nbins >= 1, "hist_quantiles: bad arguments");
  const int64_t n = inner * nq;
  hist_quantiles_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const unsigned long long*)d_hist, inner, nbins, log2(lo),
                                                                                      (log2(hi) - log2(lo)) / nbins, d_q, nq, lo, hi, d_out);
  EIGB_LAUNCH_CHECK("hist_quantiles_kernel");
  return EIGB200_OK;
}
