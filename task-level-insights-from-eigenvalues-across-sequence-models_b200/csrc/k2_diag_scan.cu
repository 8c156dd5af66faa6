// k2_diag_scan.cu -- K2a: diagonal complex linear recurrence  h_t = lam * h_{t-1} + Bu_t   (LRU / S5).
//
// Reference operator: jax.lax.associative_scan(binary_operator_diag, (Lambda_elements, Bu_elements))
//   models/lru.py:14-19, :92-95;  models/s5.py:51-62, :78-85 (reverse=True for the bidirectional half).
// The reference materialises Lambda_elements = repeat(lam, T) (lru.py:92) -- the per-token eigenvalue array; here the
// eigenvalue stays in two registers.
//
// Roofline: HBM.  Algorithmic bytes per state update = 16 (read Bu 8, write h 8); 8 flops.
// Layout: Bu/h (B,T,P) complex64 interleaved, P contiguous.  A thread owns CPT adjacent channels of one sequence and
// walks time; a warp therefore touches 32*CPT*8 contiguous bytes per time step (256 B or 512 B, 128-bit accesses for
// CPT=2).  Time is unrolled by U so that U independent vector loads are in flight per thread while the serial
// complex FMA chain (2 dependent FMAs per step) runs.  When B*P is too small to fill 148 SMs the time axis is split
// into chunks: pass 1 computes per-chunk end states (zero initial state), a tiny carry kernel combines them
// (carry_c = lam^len * carry_{c-1} + end_{c-1}), pass 2 re-runs each chunk from its carry and writes h (24 B/update).
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace eigb200 {


struct cplx { float re, im; };
__device__ __forceinline__ cplx cfma(cplx a, cplx h, cplx b) {      // a*h + b
  cplx r;
  r.re = fmaf(a.re, h.re, fmaf(-a.im, h.im, b.re));
  r.im = fmaf(a.re, h.im, fmaf(a.im, h.re, b.im));
  return r;
}

// MODE 0: single pass (write h).  MODE 1: chunk end-state only.  MODE 2: start from carry, write h.
// U: time steps (independent vector loads) in flight per thread.  8 when B*P fills the machine; 32 for the mid-sized case (BASELINE C3 at 128 sequences per
// GPU: 221 threads per SM) where 8 leave 14 KB per SM in flight -- far below what HBM latency needs -- and the old answer, splitting time into chunks, paid a
// second read of Bu (24 instead of 16 bytes per update: 0.49 of the copy peak).
template <int CPT, int MODE, int DS_U = 8, bool PIPE = false>
__global__ void __launch_bounds__(128) diag_scan_kernel(const float* __restrict__ lam, const float* __restrict__ Bu, float* __restrict__ h,
                                                        float* __restrict__ chunk_state, int64_t Bn, int64_t T, int P, int chunk_len, int nchunks, int reverse) {
  const int pv = P / CPT;                                       // channel groups per sequence
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = gid / pv;
  const int pg = (int)(gid - b * pv);
  const int chunk = blockIdx.y;
  if (b >= Bn) return;
  const int p0 = pg * CPT;
  cplx a[CPT], s[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) { a[c].re = lam[2 * (p0 + c)]; a[c].im = lam[2 * (p0 + c) + 1]; s[c].re = 0.f; s[c].im = 0.f; }
  // chunk time range in processing order
  const int64_t k0 = (int64_t)chunk * chunk_len;
  const int64_t k1 = min(T, k0 + chunk_len);
  if (MODE == 2) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float* cs = chunk_state + (((size_t)b * nchunks + chunk) * P + p0 + c) * 2;
      s[c].re = cs[0]; s[c].im = cs[1];
    }
  }
  const size_t base = ((size_t)b * T) * P + p0;                 // complex index of (b, 0, p0)
  using vec_t = typename std::conditional<CPT == 2, float4, float2>::type;
  const vec_t* src = reinterpret_cast<const vec_t*>(Bu);
  vec_t* dst = reinterpret_cast<vec_t*>(h);
  const size_t vbase = base / CPT;
  const size_t vstride = (size_t)P / CPT;

  auto load = [&](vec_t (&v)[DS_U], int64_t k) {
#pragma unroll
    for (int u = 0; u < DS_U; ++u) {
      const int64_t kk = k + u;
      if (kk < k1) {
        const int64_t t = reverse ? (T - 1 - kk) : kk;
        if constexpr (CPT == 2) v[u] = ldg_stream_f4(src + vbase + (size_t)t * vstride);
        else v[u] = ldg_stream_f2(src + vbase + (size_t)t * vstride);
      }
    }
  };
  // PIPE: the next trip's loads are issued before this trip's recurrence (two register sets): a thread always has DS_U .. 2 DS_U loads in flight instead
  // of alternating between "all in flight" and "none" -- the mid-sized case has too few warps per SM to cover that bubble with other warps
  vec_t vn[PIPE ? DS_U : 1];
  if constexpr (PIPE) load(vn, k0);
  for (int64_t k = k0; k < k1; k += DS_U) {
    vec_t v[DS_U];
    if constexpr (PIPE) {
#pragma unroll
      for (int u = 0; u < DS_U; ++u) v[u] = vn[u];
      if (k + DS_U < k1) load(vn, k + DS_U);
    } else {
      load(v, k);
    }
#pragma unroll
    for (int u = 0; u < DS_U; ++u) {
      const int64_t kk = k + u;
      if (kk < k1) {
        const int64_t t = reverse ? (T - 1 - kk) : kk;
        if constexpr (CPT == 2) {
          s[0] = cfma(a[0], s[0], cplx{v[u].x, v[u].y});
          s[1] = cfma(a[1], s[1], cplx{v[u].z, v[u].w});
          if (MODE != 1) dst[vbase + (size_t)t * vstride] = make_float4(s[0].re, s[0].im, s[1].re, s[1].im);
        } else {
          s[0] = cfma(a[0], s[0], cplx{v[u].x, v[u].y});
          if (MODE != 1) dst[vbase + (size_t)t * vstride] = make_float2(s[0].re, s[0].im);
        }
      }
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float* cs = chunk_state + (((size_t)b * nchunks + chunk) * P + p0 + c) * 2;
      cs[0] = s[c].re; cs[1] = s[c].im;
    }
  }
}

// chunk_state[b][c][p] holds the zero-init end state of chunk c on entry, the carry INTO chunk c on exit.
__global__ void diag_scan_carry_kernel(const float* __restrict__ lam, float* chunk_state, int64_t BP, int P, int nchunks, int64_t T, int chunk_len) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= BP) return;
  const int64_t b = gid / P; const int p = (int)(gid - b * P);
  const cplx a{lam[2 * p], lam[2 * p + 1]};
  cplx apow{1.f, 0.f};                                            // lam^chunk_len by sequential products (same rounding style as the scan)
  for (int i = 0; i < chunk_len; ++i) apow = cfma(a, apow, cplx{0.f, 0.f});
  cplx carry{0.f, 0.f};
  for (int c = 0; c < nchunks; ++c) {
    float* cs = chunk_state + (((size_t)b * nchunks + c) * P + p) * 2;
    const cplx end{cs[0], cs[1]};
    cs[0] = carry.re; cs[1] = carry.im;
    carry = cfma(apow, carry, end);                               // every chunk but the last is full length
  }
}

// Parameter-only eigenvalues of the diagonal SSMs (analysis/eval_eig.py:303-329):
//   kind 0 (LRU):        lam = exp(-exp(nu_log) + i exp(theta_log))                  models/lru.py:88
//   kind 1 (S5, ZOH):    lam = exp((Lambda_re + i Lambda_im) * exp(log_step))         models/s5.py:34-46
//   kind 2 (S5, bilinear): lam = (1 + d/2 L) / (1 - d/2 L)                            models/s5.py:16-31
__global__ void ssm_lambda_kernel(int kind, const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2, int P, float* __restrict__ lam) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float re, im;
  if (kind == 0) {
    const float mag = expf(-expf(p0[i])), th = expf(p1[i]);
    float sn, cs; sincosf(th, &sn, &cs);
    re = mag * cs; im = mag * sn;
  } else {
    const float step = expf(p2[i]);
    const float lr = p0[i], li = p1[i];
    if (kind == 1) {
      const float mag = expf(lr * step);
      float sn, cs; sincosf(li * step, &sn, &cs);
      re = mag * cs; im = mag * sn;
    } else {
      const float hr = 0.5f * step * lr, hi = 0.5f * step * li;        // (1 + h) / (1 - h)
      const float nr = 1.f + hr, ni = hi, dr = 1.f - hr, di = -hi;
      const float den = dr * dr + di * di;
      re = (nr * dr + ni * di) / den; im = (ni * dr - nr * di) / den;
    }
  }
  lam[2 * i] = re; lam[2 * i + 1] = im;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_ssm_lambda(void* stream, int kind, const float* d_p0, const float* d_p1, const float* d_p2, int P, float* d_lam) {
  EIGB_CHECK_ARG(kind >= 0 && kind <= 2, "ssm_lambda: kind %d not in {0 LRU, 1 S5-ZOH, 2 S5-bilinear}", kind);
  EIGB_CHECK_ARG(d_p0 && d_p1 && d_lam && (kind == 0 || d_p2) && P > 0, "ssm_lambda: bad arguments");
  ssm_lambda_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kind, d_p0, d_p1, d_p2, P, d_lam);
  EIGB_LAUNCH_CHECK("ssm_lambda_kernel");
  return EIGB200_OK;
}

// Workspace for the chunked variant is allocated lazily per call from a small per-thread cache (stream-ordered).
extern "C" int eigb200_diag_scan(void* stream, const float* d_lam, const float* d_Bu, float* d_h, int64_t B, int64_t T, int P, int reverse) {
  EIGB_CHECK_ARG(d_lam && d_Bu && d_h, "diag_scan: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0 && P > 0, "diag_scan: bad shape B=%lld T=%lld P=%d", (long long)B, (long long)T, P);
  EIGB_CHECK_ARG(((uintptr_t)d_Bu & 15) == 0 && ((uintptr_t)d_h & 15) == 0, "diag_scan: Bu and h must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = num_sms();
  const int64_t channels = B * P;
  const bool cpt2 = (P % 2 == 0) && (channels / 2 >= (int64_t)sms * 1024);
  const int cpt = cpt2 ? 2 : 1;
  const int64_t threads = channels / cpt;
  // mid-sized batches: one pass with 32 loads in flight per thread (>= 2 warps per SM)
  if (threads >= (int64_t)sms * 64 && threads < (int64_t)sms * 512 && getenv("EIGB200_DIAG_DEEP") == nullptr) {
    const unsigned gx1 = (unsigned)((threads + 127) / 128);
    static const bool nopipe = getenv("EIGB200_DIAG_NOPIPE") != nullptr;
    if (cpt == 2) diag_scan_kernel<2, 0, 16><<<dim3(gx1, 1), 128, 0, st>>>(d_lam, d_Bu, d_h, nullptr, B, T, P, (int)T, 1, reverse);
    else if (nopipe) diag_scan_kernel<1, 0, 32><<<dim3(gx1, 1), 128, 0, st>>>(d_lam, d_Bu, d_h, nullptr, B, T, P, (int)T, 1, reverse);
    else diag_scan_kernel<1, 0, 32, true><<<dim3(gx1, 1), 128, 0, st>>>(d_lam, d_Bu, d_h, nullptr, B, T, P, (int)T, 1, reverse);
    EIGB_LAUNCH_CHECK("diag_scan_kernel");
    return EIGB200_OK;
  }
  // split time only when the channel parallelism alone leaves most of the machine idle
  constexpr int DS_U = 8;
  int nchunks = 1;
  if (threads < (int64_t)sms * 256) {
    nchunks = (int)(((int64_t)sms * 512 + threads - 1) / threads);
    const int64_t maxc = (T + 4 * DS_U - 1) / (4 * DS_U);
    if (nchunks > maxc) nchunks = (int)maxc;
    if (nchunks > 64) nchunks = 64;
    if (nchunks < 1) nchunks = 1;
  }
  int chunk_len = (int)((T + nchunks - 1) / nchunks);
  nchunks = (int)((T + chunk_len - 1) / chunk_len);
  const unsigned gx = (unsigned)((threads + 127) / 128);
  if (nchunks == 1) {
    dim3 grid(gx, 1);
    if (cpt == 2) diag_scan_kernel<2, 0><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, nullptr, B, T, P, chunk_len, 1, reverse);
    else diag_scan_kernel<1, 0><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, nullptr, B, T, P, chunk_len, 1, reverse);
    EIGB_LAUNCH_CHECK("diag_scan_kernel");
    return EIGB200_OK;
  }
  float* ws = nullptr;
  {
    // the chunk states come from the device's stream-ordered pool; keep freed blocks cached across synchronisation points (the default
    // release threshold of 0 hands them back to the driver at every sync, which turns the next call's allocation into a cudaMalloc)
    static thread_local int pool_ready_dev = -1;
    int dev = 0;
    EIGB_CUDA(cudaGetDevice(&dev));
    if (pool_ready_dev != dev) {
      cudaMemPool_t pool;
      EIGB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = 256ull << 20;
      EIGB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pool_ready_dev = dev;
    }
  }
  EIGB_CUDA(cudaMallocAsync((void**)&ws, (size_t)B * nchunks * P * 2 * sizeof(float), st));
  dim3 grid(gx, nchunks);
  if (cpt == 2) diag_scan_kernel<2, 1><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, ws, B, T, P, chunk_len, nchunks, reverse);
  else diag_scan_kernel<1, 1><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, ws, B, T, P, chunk_len, nchunks, reverse);
  diag_scan_carry_kernel<<<(unsigned)((channels + 127) / 128), 128, 0, st>>>(d_lam, ws, channels, P, nchunks, T, chunk_len);
  if (cpt == 2) diag_scan_kernel<2, 2><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, ws, B, T, P, chunk_len, nchunks, reverse);
  else diag_scan_kernel<1, 2><<<grid, 128, 0, st>>>(d_lam, d_Bu, d_h, ws, B, T, P, chunk_len, nchunks, reverse);
  EIGB_LAUNCH_CHECK("diag_scan chunked");
  EIGB_CUDA(cudaFreeAsync(ws, st));
  return EIGB200_OK;
}
