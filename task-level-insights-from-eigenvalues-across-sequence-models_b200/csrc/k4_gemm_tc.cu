// k4_gemm_tc.cu -- K4: nn.Linear on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Reference operators: every nn.Linear of the analysis path -- SSD in_proj / out_proj (models/mamba.py:64, :109, :118, :153),
// GLU (models/common.py:53-57), Wqkv / Wvqkn / out_proj / MLP (models/attention.py:120-132; norm_attention.py:201-215; common.py:37-46).
// The reference runs them as cuBLAS SGEMM in full fp32 (torch default allow_tf32=False), so the tensor-core path keeps fp32-level
// accuracy with the 3xTF32 split:  a = a_hi + a_lo,  w = w_hi + w_lo  (hi = round-to-nearest tf32, lo = tf32(x - hi)),
//   A W^T  ~=  A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T      (dropped term a_lo*w_lo ~ 2^-22 relative), fp32 accumulation in TMEM.
//
// Shape regime: M = B*T is huge (2.1e6 rows at BASELINE C2), K = d_model is small (<= 256), N <= a few hundred: a MEMORY-bound GEMM
// (read K*4 + write N*4 bytes per row).  Design:
//   * persistent CTAs, one per SM (cta_group::1, UMMA M = 128 rows per tile, N = BN <= 128 columns per CTA);
//   * the whole weight slice of the CTA (BN x K, hi and lo, K-major SWIZZLE_128B) is loaded ONCE by TMA and stays in shared memory;
//   * A streams through a ring of 32-column (128-byte) K-chunks: TMA (SWIZZLE_128B box 32 x 128) -> converter warps split the
//     chunk into hi (in place) and lo -> fence.proxy.async -> one thread issues 12 tcgen05.mma (4 k-steps x 3 terms) per chunk;
//   * accumulators are double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue warps pull the accumulator with tcgen05.ld (32 lanes x 32 columns per warp instruction), apply bias / exact-erf GELU /
//     GLU gate / residual in registers and store 128-bit vectors;
//   * N wider than one CTA's slice is split across `nsplit` CTAs that walk the same row tiles together (the second read of the A tile
//     is an L2 hit).  For the GLU epilogue a CTA's slice is [BG value columns | their BG gate columns].
// Warp roles (448 threads): warps 0-7 epilogue (TMEM lane quarter = warp % 4, even / odd 32-column groups = warp / 4), warps 8-11 converters,
// warp 12 TMA producer, warp 13 MMA issuer.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace eigb200 {

constexpr int TC_BM = 128;                 // rows per tile (UMMA M)
constexpr int TC_KC = 32;                  // fp32 columns per K-chunk = one 128-byte swizzle row
constexpr int TC_CHUNK_BYTES = TC_BM * TC_KC * 4;          // 16 KB
constexpr int TC_EPI_WARPS = 16;           // epilogue warps: TMEM lane quarter = warp % 4, 32-column group = warp / 4 (+ TC_EPI_GROUPS per trip)
constexpr int TC_EPI_GROUPS = TC_EPI_WARPS / 4;
#ifdef EIGB_ROLE_ORDER                                       // experiment: producers in the lowest warp ids (converters 0-3, MMA 4, TMA 5, 6-7 idle, epilogue 8-23)
constexpr int TC_CONV_WARP0 = 0;
constexpr int TC_MMA_WARP = 4;
constexpr int TC_TMA_WARP = 5;
constexpr int TC_EPI_WARP0 = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 8) * 32;
#else
constexpr int TC_EPI_WARP0 = 0;
constexpr int TC_CONV_WARP0 = TC_EPI_WARPS;                  // 4 converter warps
constexpr int TC_TMA_WARP = TC_EPI_WARPS + 4;
constexpr int TC_MMA_WARP = TC_EPI_WARPS + 5;
constexpr int TC_THREADS = (TC_EPI_WARPS + 6) * 32;
#endif
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_SMEM_LIMIT = 227 * 1024;

struct TcParams {
  const float* bias; float* C; int64_t ldc; const float* R; int64_t ldr;
  const float2* ln_stats;                                                   // LayerNorm (mean, rstd) per row applied by the converter (nullable)
  int64_t M; int N, K, epilogue;
  int bn;            // columns per CTA (UMMA N), multiple of 32, <= 128
  int bg;            // GLU: value columns per CTA (bn = 2*bg); otherwise bn
  int nsplit, kchunks, nstages, nterms, workers;
  const float* eig_w; float* eig_part;   // GLU epilogue only (nullable): per-row partial gate dot products and moments of the OUTPUT rows, see tc_epilogue
  int r_v8;          // residual rows are 32-byte aligned: add them in the accumulator layout with 256-bit loads
  int epi_stage;     // bytes of per-warp staging tiles behind the raw ring (16 warps x 2 KB): the plain / GELU / residual epilogues transpose through
                     // shared memory (4 STS + 4 LDS per 16 columns) instead of 192 shuffle / select instructions per 32 columns
  int c_v8;          // output rows are 32-byte aligned (and any residual is r_v8): every thread stores its own row segment with 256-bit stores
  int zero;          // always 0; a run-time value the compilers cannot fold (mbar_arrive_after)
  int64_t ntiles;
  // fp16-split operands (kind::f16, see "fp16 split" below): weight chunks of 64 K-values, activation pre-scale S_a (a power of two; folded into the
  // LayerNorm constants when there is one), 1 / (S_a S_w) for the epilogue (device scalar written by the weight preparation), sticky overflow flag
  int kch_w; float a_scale; const float* out_scale; int* ovf_flag;
  int na_stages;     // streamed kernel with on-SM conversion: raw A boxes (16 KB) in the ring in front of the W stages
};

// ---------------------------------------------------------------------------------------------------------------------------
// weight preparation: split W (N,K) into tf32 hi / lo in the per-split row order the CTAs consume, zero padded
// ---------------------------------------------------------------------------------------------------------------------------
// With a fused LayerNorm the affine part is folded into the weights: LN(a) W^T = ((a - mu) rstd) (W diag(gamma))^T + W beta, so the
// converter only applies the per-row (mu, rstd) and the bias vector becomes bias + W beta (ln_bias_kernel).
__global__ void split_weights_kernel(const float* __restrict__ W, float* __restrict__ hi, float* __restrict__ lo,
                                     int N, int K, int kpad, int bn, int bg, int nsplit, int glu, const float* __restrict__ gamma) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nsplit * bn * kpad;
  if (idx >= total) return;
  const int k = idx % kpad, row = idx / kpad;
  const int split = row / bn, local = row - split * bn;
  int n;
  if (glu) {                                             // CTA slice: value / gate columns interleaved in groups of 16 (glu_weight_row)
    n = glu_weight_row(local, split, bg, N / 2);
  } else {
    n = split * bn + local;
  }
  float w = 0.f;
  if (n >= 0 && n < N && k < K) w = W[(size_t)n * K + k] * (gamma ? gamma[k] : 1.f);
  const float h = to_tf32(w);
  hi[idx] = h;
  lo[idx] = to_tf32(w - h);
}

// fp16 split of the weights.  Pass 1: max |w gamma| (positive floats order like their bit patterns -> atomicMax on the bits).
__global__ void absmax_weights_kernel(const float* __restrict__ W, const float* __restrict__ gamma, int N, int K, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * K; idx += gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(W[idx] * (gamma ? gamma[idx % K] : 1.f)));
#pragma unroll
  for (int ofs = 16; ofs >= 1; ofs >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, ofs));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}
// S_w = 2^floor(log2(2^14 / max|w|)): the largest weight lands in [2^13, 2^14], far from fp16's overflow and with 25 octaves above its subnormals
__device__ __forceinline__ float weight_scale_f16(float absmax) {
  if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
  int e;
  frexpf(16384.f / absmax, &e);                                    // x = m 2^e, m in [0.5, 1)  =>  floor(log2 x) = e - 1
  e = min(max(e - 1, -100), 100);
  return ldexpf(1.f, e);
}
// Pass 2: hi = fp16(w gamma S_w), lo = fp16(w gamma S_w - hi) in the per-split row order the CTAs consume, rows of kp64 halfs (zero padded);
// scal[0] = 1 / (S_a S_w) for the epilogue, scal[2] = S_w, scal[3] = S_a.
__global__ void split_weights_f16_kernel(const float* __restrict__ W, __half* __restrict__ hi, __half* __restrict__ lo,
                                         int N, int K, int kp64, int bn, int bg, int nsplit, int glu, const float* __restrict__ gamma,
                                         float* __restrict__ scal, float a_scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nsplit * bn * kp64;
  const float sw = weight_scale_f16(__uint_as_float(reinterpret_cast<const unsigned*>(scal)[1]));
  if (idx == 0) { scal[0] = 1.f / (a_scale * sw); scal[2] = sw; scal[3] = a_scale; }
  if (idx >= total) return;
  const int k = idx % kp64, row = idx / kp64;
  const int split = row / bn, local = row - split * bn;
  int n;
  if (glu) n = glu_weight_row(local, split, bg, N / 2);
  else n = split * bn + local;
  float w = 0.f;
  if (n >= 0 && n < N && k < K) w = W[(size_t)n * K + k] * (gamma ? gamma[k] : 1.f) * sw;
  const __half h = __float2half_rn(w);
  hi[idx] = h;
  lo[idx] = __float2half_rn(w - __half2float(h));
}

// bias2[n] = bias[n] + sum_k W[n,k] beta[k]: one warp per output column
__global__ void ln_bias_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ beta,
                               float* __restrict__ bias2, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(W[(size_t)n * K + k], beta[k], acc);
#pragma unroll
  for (int ofs = 16; ofs >= 1; ofs >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, ofs);
  if (lane == 0) bias2[n] = acc + (bias ? bias[n] : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------------
// epilogue (8 warps): TMEM accumulator -> registers -> bias / GELU / GLU gate -> shuffle transpose -> (+ residual) -> 128-byte coalesced stores
// ---------------------------------------------------------------------------------------------------------------------------
// 4x4 transpose of float4 items inside each group of 4 lanes: lane 4g+i holds the four quads of its row 4g+i; on return it holds quad i of
// rows 4g+j, j = 0..3 (4 lanes cover 64 contiguous bytes of one row).
__device__ __forceinline__ void transpose4x4_f4(float (&v)[16], int lane) {
#pragma unroll
  for (int s = 2; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if ((q & s) == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float lo = v[4 * q + e], hi = v[4 * (q | s) + e];
          const float recv = __shfl_xor_sync(0xffffffffu, up ? lo : hi, s);
          v[4 * q + e] = up ? recv : lo;
          v[4 * (q | s) + e] = up ? hi : recv;
        }
      }
    }
  }
}

// SC: the accumulator carries the operand scales S_a S_w of the fp16 split; 1 / (S_a S_w) rides in the FMA that adds the bias
template <int EPI, bool SC = false>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint32_t tmem_base, int bn, uint32_t bar_dfull0, uint32_t bar_dempty0,
                                            const float* bias_s, int worker, int split, int warp, int lane, float* stage_s = nullptr) {
  auto bar_dfull = [&](int j) { return bar_dfull0 + 8u * j; };
  auto bar_dempty = [&](int j) { return bar_dempty0 + 8u * j; };
  constexpr bool GLU = EPI == EIGB200_EPI_GLU_RESIDUAL;
  const float osc = SC ? __ldg(p.out_scale) : 1.f;
  const float osc_gate = -1.4426950408889634f * osc;
  const int nout = GLU ? p.N / 2 : p.N;
  const int cols_out = GLU ? p.bg : bn;                              // output columns produced by this CTA
  const int n_cta0 = split * cols_out;
  const int quarter = warp & 3, grp = warp >> 2;                     // TMEM lane quarter; first 32-column accumulator group of this warp
  const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
  const bool use_r = p.R && (GLU || EPI == EIGB200_EPI_RESIDUAL);
  const bool r_own = use_r && p.r_v8;                                // residual added in the accumulator layout with 256-bit loads
  const float* eigw_s = bias_s + 192;                                // [cols_out] gate weights of this CTA's output columns (768 bytes after the bias)
  int j = 0; uint32_t dph = 0;
  for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
    const int64_t own_row = tile * TC_BM + quarter * 32 + lane;      // accumulator layout: lane = row
    const uint32_t d_tmem = tmem_base + (uint32_t)(j * bn) + lane_sel;
    bool waited = false, arrived = false;
    for (int cg = 32 * grp; cg < bn; cg += 32 * TC_EPI_GROUPS) {
      const bool last = cg + 32 * TC_EPI_GROUPS >= bn;               // this warp's last TMEM read of the accumulator: release it early
      if (GLU) {
        const int oc = n_cta0 + (cg >> 1);                           // first of the 16 output columns of this group
        float rr[16];
#ifdef EIGB_ABL_EPI                                                  // ablation build (tools/ablate_gemm.sh): wait / tcgen05.ld / release only
        if (false) {
#else
        if (r_own) {                                                 // residual prefetch: its DRAM latency hides behind the accumulator wait
#endif
          const float* rptr = p.R + own_row * p.ldr + oc;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (own_row < p.M && oc + 8 * q + 8 <= nout) ldg_stream_v8(rptr + 8 * q, rr + 8 * q);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) rr[8 * q + e] = (own_row < p.M && oc + 8 * q + e < nout) ? __ldg(rptr + 8 * q + e) : 0.f;
            }
          }
        }
        if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); waited = true; }
        float a[32];
        tmem_ld_32x32(d_tmem + cg, a);
        if (last) { tc_fence_before(); mbar_arrive(bar_dempty(j)); arrived = true; }
#ifdef EIGB_ABL_EPI
        if (a[0] == 12345.678f && a[31] == -9.f) p.C[0] = a[5];
        continue;
#endif
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {                               // gate bias pre-scaled by -log2(e): sigmoid(g + b) = 1 / (1 + 2^(g * -log2e + b'))
          float e2;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(a[16 + i], SC ? osc_gate : -1.4426950408889634f, bias_s[cg + 16 + i])));
          v[i] = (SC ? fmaf(a[i], osc, bias_s[cg + i]) : (a[i] + bias_s[cg + i])) * fast_rcp_f(1.f + e2);
        }
        if (r_own) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += rr[i];
        }
        if (p.eig_part && r_own) {
          // Extractor fusion: the finished output row segment (16 columns of x_out = GLU + skip) contributes its part of the eigenvalue gate
          // x_out . W_dt and of the LayerNorm moments of the NEXT block, so that no kernel has to re-read x_out: per (row, 16-column group)
          // (dot, mean, M2) go to part[(g16 * 3 + c) * M + row] (coalesced over rows), combined in a fixed order by eigb200_mamba2_eig_partials.
          const float* we = eigw_s + (cg >> 1);
          float dot = 0.f, sum = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) { dot = fmaf(v[i], we[i], dot); sum += v[i]; }
          const float mean = sum * 0.0625f;
          float m2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; m2 = fmaf(d, d, m2); }
          if (own_row < p.M) {
            float* pp = p.eig_part + (size_t)(oc >> 4) * 3 * p.M + own_row;
            pp[0] = dot; pp[p.M] = mean; pp[2 * p.M] = m2;
          }
        }
        if (p.c_v8) {                                                // own row, 2 full sectors: no transpose
          if (own_row < p.M) {
            float* cptr = p.C + own_row * p.ldc + oc;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (oc + 8 * q + 8 <= nout) stg_v8(cptr + 8 * q, v + 8 * q);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (oc + 8 * q + e < nout) cptr[8 * q + e] = v[8 * q + e];
              }
            }
          }
          continue;
        }
        transpose4x4_f4(v, lane);
        const int n = oc + 4 * (lane & 3);
        const int64_t row0 = tile * TC_BM + quarter * 32 + (lane & ~3);
        if (n < nout) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int64_t mrow = row0 + jj;
            if (mrow < p.M) {
              float* cptr = p.C + mrow * p.ldc + n;
              if (n + 3 < nout) {
                float4 o = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                if (use_r && !p.r_v8) {
                  const float4 r4 = ldg_stream_f4(reinterpret_cast<const float4*>(p.R + mrow * p.ldr + n));
                  o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                }
                *reinterpret_cast<float4*>(cptr) = o;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < nout) cptr[e] = v[4 * jj + e] + ((use_r && !p.r_v8) ? p.R[mrow * p.ldr + n + e] : 0.f);
              }
            }
          }
        }
      } else {
        const int gi = lane & 7, gg = lane >> 3;                     // after the transpose: column quad gi of rows 8*gg + jj
        const int64_t row0 = tile * TC_BM + quarter * 32 + 8 * gg;
        const int n = n_cta0 + cg + 4 * gi;
        const bool col_ok = n < nout;
        const bool full = n + 3 < nout;
        const float* rptr = p.R + own_row * p.ldr + n_cta0 + cg;
        float rr[32];
#ifdef EIGB_ABL_EPI
        if (false) {
#else
        if (r_own) {                                                 // prefetch behind the accumulator wait
#endif
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (own_row < p.M && n_cta0 + cg + 8 * q + 8 <= nout) ldg_stream_v8(rptr + 8 * q, rr + 8 * q);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) rr[8 * q + e] = (own_row < p.M && n_cta0 + cg + 8 * q + e < nout) ? __ldg(rptr + 8 * q + e) : 0.f;
            }
          }
        }
        if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); waited = true; }
        float v[32];
        tmem_ld_32x32(d_tmem + cg, v);
        if (last) { tc_fence_before(); mbar_arrive(bar_dempty(j)); arrived = true; }
#ifdef EIGB_ABL_EPI
        if (v[0] == 12345.678f && v[31] == -9.f) p.C[0] = v[5];
        continue;
#endif
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float vv = SC ? fmaf(v[i], osc, bias_s[cg + i]) : v[i] + bias_s[cg + i];
          if (EPI == EIGB200_EPI_GELU) vv = gelu_fast_f(vv);
          v[i] = vv;
        }
        if (r_own) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += rr[i];
        }
        if (stage_s != nullptr) {
          // transpose through this warp's 2 KB staging tile, 16 columns at a time: thread = row writes 4 quads (quad position XOR-swizzled by the
          // row pair: conflict-free), then 4 lanes cover 64 contiguous bytes of a row and a store instruction writes 8 rows x 64 B
          float* st = stage_s + warp * 512;
          const int srow = lane >> 2, sq = lane & 3;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(st + lane * 16 + ((q ^ ((lane >> 1) & 3)) << 2)) =
                  make_float4(v[16 * half + 4 * q], v[16 * half + 4 * q + 1], v[16 * half + 4 * q + 2], v[16 * half + 4 * q + 3]);
            __syncwarp();
            const int ncol = n_cta0 + cg + 16 * half + 4 * sq;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr_ = 8 * it + srow;
              const float4 o = *reinterpret_cast<const float4*>(st + rr_ * 16 + ((sq ^ ((rr_ >> 1) & 3)) << 2));
              const int64_t mrow = tile * TC_BM + quarter * 32 + rr_;
              if (mrow < p.M && ncol < nout) {
                float* cptr = p.C + mrow * p.ldc + ncol;
                if (ncol + 3 < nout) *reinterpret_cast<float4*>(cptr) = o;
                else {
                  const float oe[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) if (ncol + e < nout) cptr[e] = oe[e];
                }
              }
            }
            __syncwarp();
          }
          continue;
        }
        if (p.c_v8) {                                                // own row, 4 full sectors: no transpose
          if (own_row < p.M) {
            float* cptr = p.C + own_row * p.ldc + n_cta0 + cg;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (n_cta0 + cg + 8 * q + 8 <= nout) stg_v8(cptr + 8 * q, v + 8 * q);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (n_cta0 + cg + 8 * q + e < nout) cptr[8 * q + e] = v[8 * q + e];
              }
            }
          }
          continue;
        }
        transpose8x8_f4(v, lane);
        if (col_ok) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int64_t mrow = row0 + jj;
            if (mrow < p.M) {
              float* cptr = p.C + mrow * p.ldc + n;
              if (full) {
                float4 o = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                if (use_r && !p.r_v8) {
                  const float4 r4 = ldg_stream_f4(reinterpret_cast<const float4*>(p.R + mrow * p.ldr + n));
                  o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                }
                *reinterpret_cast<float4*>(cptr) = o;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < nout) cptr[e] = v[4 * jj + e] + ((use_r && !p.r_v8) ? p.R[mrow * p.ldr + n + e] : 0.f);
              }
            }
          }
        }
      }
    }
    if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); }
    if (!arrived) { tc_fence_before(); mbar_arrive(bar_dempty(j)); }
    if (++j == 2) { j = 0; dph ^= 1; }
  }
}

// Epilogue of the streamed-operand kernel (a separate copy: the resident-weight kernels are sensitive to any change of their epilogue's code
// generation -- folding the two behind a template flag cost them 10 %): the CTA walks (row tile, N tile) pairs blockIdx.x, + gridDim.x, ... with the N tile fastest; bias_s points
// at the whole bias vector in accumulator-column order in GLOBAL memory (p.nsplit * bn entries).
template <int EPI, bool STREAM = true, bool SC = false>
__device__ __forceinline__ void tc_epilogue_stream(const TcParams& p, uint32_t tmem_base, int bn, uint32_t bar_dfull0, uint32_t bar_dempty0,
                                            const float* bias_s, int worker, int split, int warp, int lane) {
  auto bar_dfull = [&](int j) { return bar_dfull0 + 8u * j; };
  auto bar_dempty = [&](int j) { return bar_dempty0 + 8u * j; };
  constexpr bool GLU = EPI == EIGB200_EPI_GLU_RESIDUAL;
  const float osc = SC ? __ldg(p.out_scale) : 1.f;                   // fp16 split: the accumulator carries S_a S_w
  const int nout = GLU ? p.N / 2 : p.N;
  const int cols_out = GLU ? p.bg : bn;                              // output columns produced by this CTA
  int n_cta0 = split * cols_out;
  const int quarter = warp & 3, grp = warp >> 2;                     // TMEM lane quarter; first 32-column accumulator group of this warp
  const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
  const bool use_r = p.R && (GLU || EPI == EIGB200_EPI_RESIDUAL);
  const bool r_own = use_r && p.r_v8;                                // residual added in the accumulator layout with 256-bit loads
  int j = 0; uint32_t dph = 0;
  const int64_t it_end = STREAM ? p.ntiles * p.nsplit : p.ntiles;
  const int64_t it_step = STREAM ? (int64_t)gridDim.x : (int64_t)p.workers;
  const float* bias0 = bias_s;
  for (int64_t it = STREAM ? (int64_t)blockIdx.x : (int64_t)worker; it < it_end; it += it_step) {
    const int64_t tile = STREAM ? it / p.nsplit : it;
    if (STREAM) {
      const int sp = (int)(it - tile * p.nsplit);
      n_cta0 = sp * cols_out;
      bias_s = bias0 + (size_t)sp * bn;
    }
    const int64_t own_row = tile * TC_BM + quarter * 32 + lane;      // accumulator layout: lane = row
    const uint32_t d_tmem = tmem_base + (uint32_t)(j * bn) + lane_sel;
    bool waited = false, arrived = false;
    for (int cg = 32 * grp; cg < bn; cg += 32 * TC_EPI_GROUPS) {
      const bool last = cg + 32 * TC_EPI_GROUPS >= bn;               // this warp's last TMEM read of the accumulator: release it early
      if (GLU) {
        const int oc = n_cta0 + (cg >> 1);                           // first of the 16 output columns of this group
        float rr[16];
        if (r_own) {                                                 // residual prefetch: its DRAM latency hides behind the accumulator wait
          const float* rptr = p.R + own_row * p.ldr + oc;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (own_row < p.M && oc + 8 * q + 8 <= nout) ldg_stream_v8(rptr + 8 * q, rr + 8 * q);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) rr[8 * q + e] = (own_row < p.M && oc + 8 * q + e < nout) ? __ldg(rptr + 8 * q + e) : 0.f;
            }
          }
        }
        if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); waited = true; }
        float a[32];
        tmem_ld_32x32(d_tmem + cg, a);
        if (last) { tc_fence_before(); mbar_arrive(bar_dempty(j)); arrived = true; }
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] = SC ? fmaf(a[i], osc, bias_s[cg + i]) * sigmoid_fast_f(fmaf(a[16 + i], osc, bias_s[cg + 16 + i]))
                    : (a[i] + bias_s[cg + i]) * sigmoid_fast_f(a[16 + i] + bias_s[cg + 16 + i]);
        if (r_own) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += rr[i];
        }
        transpose4x4_f4(v, lane);
        const int n = oc + 4 * (lane & 3);
        const int64_t row0 = tile * TC_BM + quarter * 32 + (lane & ~3);
        if (n < nout) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int64_t mrow = row0 + jj;
            if (mrow < p.M) {
              float* cptr = p.C + mrow * p.ldc + n;
              if (n + 3 < nout) {
                float4 o = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                if (use_r && !p.r_v8) {
                  const float4 r4 = ldg_stream_f4(reinterpret_cast<const float4*>(p.R + mrow * p.ldr + n));
                  o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                }
                *reinterpret_cast<float4*>(cptr) = o;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < nout) cptr[e] = v[4 * jj + e] + ((use_r && !p.r_v8) ? p.R[mrow * p.ldr + n + e] : 0.f);
              }
            }
          }
        }
      } else {
        const int gi = lane & 7, gg = lane >> 3;                     // after the transpose: column quad gi of rows 8*gg + jj
        const int64_t row0 = tile * TC_BM + quarter * 32 + 8 * gg;
        const int n = n_cta0 + cg + 4 * gi;
        const bool col_ok = n < nout;
        const bool full = n + 3 < nout;
        const float* rptr = p.R + own_row * p.ldr + n_cta0 + cg;
        float rr[32];
        if (r_own) {                                                 // prefetch behind the accumulator wait
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (own_row < p.M && n_cta0 + cg + 8 * q + 8 <= nout) ldg_stream_v8(rptr + 8 * q, rr + 8 * q);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) rr[8 * q + e] = (own_row < p.M && n_cta0 + cg + 8 * q + e < nout) ? __ldg(rptr + 8 * q + e) : 0.f;
            }
          }
        }
        if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); waited = true; }
        float v[32];
        tmem_ld_32x32(d_tmem + cg, v);
        if (last) { tc_fence_before(); mbar_arrive(bar_dempty(j)); arrived = true; }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float vv = SC ? fmaf(v[i], osc, bias_s[cg + i]) : v[i] + bias_s[cg + i];
          if (EPI == EIGB200_EPI_GELU) vv = gelu_fast_f(vv);
          v[i] = vv;
        }
        if (r_own) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += rr[i];
        }
        transpose8x8_f4(v, lane);
        if (col_ok) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int64_t mrow = row0 + jj;
            if (mrow < p.M) {
              float* cptr = p.C + mrow * p.ldc + n;
              if (full) {
                float4 o = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                if (use_r && !p.r_v8) {
                  const float4 r4 = ldg_stream_f4(reinterpret_cast<const float4*>(p.R + mrow * p.ldr + n));
                  o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                }
                *reinterpret_cast<float4*>(cptr) = o;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < nout) cptr[e] = v[4 * jj + e] + ((use_r && !p.r_v8) ? p.R[mrow * p.ldr + n + e] : 0.f);
              }
            }
          }
        }
      }
    }
    if (!waited) { mbar_wait(bar_dfull(j), dph); tc_fence_after(); }
    if (!arrived) { tc_fence_before(); mbar_arrive(bar_dempty(j)); }
    if (++j == 2) { j = 0; dph ^= 1; }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// the GEMM kernel
// ---------------------------------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapWhi,
               const __grid_constant__ CUtensorMap tmapWlo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned base (SWIZZLE_128B atoms)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int bn = p.bn, kch = p.kchunks, nst = p.nstages;
  const uint32_t w_chunk_bytes = (uint32_t)bn * 128u;
  const uint32_t whi = base;
  const uint32_t wlo = whi + kch * w_chunk_bytes;
  const uint32_t stage0 = wlo + kch * w_chunk_bytes;                 // stage s: [hi/raw 16 KB][lo 16 KB]
  const uint32_t bars = stage0 + nst * 2 * TC_CHUNK_BYTES;
  // barrier slots (8 bytes each)
  const uint32_t bar_w = bars;                                       // weights landed
  auto bar_full = [&](int s) { return bars + 8u * (1 + s); };                         // TMA landed raw chunk
  auto bar_conv = [&](int s) { return bars + 8u * (1 + TC_MAX_STAGES + s); };         // converters done
  auto bar_empty = [&](int s) { return bars + 8u * (1 + 2 * TC_MAX_STAGES + s); };    // MMAs reading the stage retired
  auto bar_dfull = [&](int j) { return bars + 8u * (1 + 3 * TC_MAX_STAGES + j); };    // accumulator j complete
  auto bar_dempty = [&](int j) { return bars + 8u * (3 + 3 * TC_MAX_STAGES + j); };   // accumulator j drained
  const uint32_t tmem_slot = bars + 8u * (5 + 3 * TC_MAX_STAGES);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));     // [bn] bias of this CTA's slice

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = blockIdx.x / p.nsplit, split = blockIdx.x - worker * p.nsplit;
  const uint32_t tmem_cols = (2 * bn <= 32) ? 32 : (2 * bn <= 64) ? 64 : (2 * bn <= 128) ? 128 : (2 * bn <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < nst; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_conv(s), 128); mbar_init(bar_empty(s), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_dfull(j), 1); mbar_init(bar_dempty(j), TC_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (threadIdx.x < 128) {                                            // bias of the slice, in accumulator-column order
    constexpr bool GLU_ = EPI == EIGB200_EPI_GLU_RESIDUAL;
    const int nout_ = GLU_ ? p.N / 2 : p.N;
    for (int c = threadIdx.x; c < bn; c += 128) {
      int n;
      if (GLU_) n = glu_weight_row(c, split, p.bg, nout_);
      else { n = split * bn + c; if (n >= p.N) n = -1; }
      float bv = (p.bias && n >= 0) ? p.bias[n] : 0.f;
      if (GLU_ && (c & 16)) bv *= -1.4426950408889634f;            // gate columns of each 32-column accumulator group: tc_epilogue folds the scale into an FMA
      bias_s[c] = bv;
    }
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  if (warp == TC_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmapA); tma_prefetch_desc(&tmapWhi); tma_prefetch_desc(&tmapWlo); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == TC_TMA_WARP) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, 2u * kch * w_chunk_bytes);
      for (int c = 0; c < kch; ++c) {
        tma_load_2d(&tmapWhi, bar_w, whi + c * w_chunk_bytes, c * TC_KC, split * bn);
        tma_load_2d(&tmapWlo, bar_w, wlo + c * w_chunk_bytes, c * TC_KC, split * bn);
      }
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
        for (int c = 0; c < kch; ++c) {
          mbar_wait_one(bar_empty(s), ph ^ 1);
          mbar_arrive_expect_tx(bar_full(s), TC_CHUNK_BYTES);
          tma_load_2d(&tmapA, bar_full(s), stage0 + s * 2 * TC_CHUNK_BYTES, c * TC_KC, (int)(tile * TC_BM));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= TC_CONV_WARP0 && warp < TC_CONV_WARP0 + 4) {
    // ===================================== converters: raw fp32 -> tf32 hi (in place) + tf32 lo ==============================
    const int ct = threadIdx.x - TC_CONV_WARP0 * 32;                               // 0..127
    const bool ln = p.ln_stats != nullptr;
    int s = 0; uint32_t ph = 0;
    for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
      // float4 #idx = i*128 + ct of a chunk is row idx>>3, physical 16-byte slot idx&7 = logical slot ^ (row & 7) (SWIZZLE_128B):
      // this thread's 8 rows are the same for every chunk of the tile, so their LayerNorm statistics are fetched once per tile
      float2 st[8];
      if (ln) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t m = tile * TC_BM + ((i * 128 + ct) >> 3);
          st[i] = m < p.M ? __ldg(p.ln_stats + m) : make_float2(0.f, 0.f);
          st[i].x = -st[i].x * st[i].y;                               // (a - mu) rstd = fma(a, rstd, -mu rstd)
        }
      }
      for (int c = 0; c < kch; ++c) {
        mbar_wait(bar_full(s), ph);
        uint8_t* hi_ptr = smem_raw + (stage0 + s * 2 * TC_CHUNK_BYTES - smem_u32(smem_raw));
        float4* h4 = reinterpret_cast<float4*>(hi_ptr);
        float4* l4 = reinterpret_cast<float4*>(hi_ptr + TC_CHUNK_BYTES);
        if (ln) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + ct;
            float4 a = h4[idx];
            a.x = fmaf(a.x, st[i].y, st[i].x); a.y = fmaf(a.y, st[i].y, st[i].x);
            a.z = fmaf(a.z, st[i].y, st[i].x); a.w = fmaf(a.w, st[i].y, st[i].x);
            float4 h, l;
            split_tf32_4(a, h, l);
            h4[idx] = h; l4[idx] = l;
          }
        } else if (p.nterms == 3) {
#pragma unroll
          for (int i = 0; i < TC_CHUNK_BYTES / 16 / 128; ++i) {      // 8 float4 per thread; the split is elementwise, layout agnostic
            const int idx = i * 128 + ct;
            const float4 a = h4[idx];
            float4 h, l;
            split_tf32_4(a, h, l);
            h4[idx] = h; l4[idx] = l;
          }
        } else {
#pragma unroll
          for (int i = 0; i < TC_CHUNK_BYTES / 16 / 128; ++i) {
            const int idx = i * 128 + ct;
            const float4 a = h4[idx];
            h4[idx] = make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
          }
        }
        fence_proxy_async();                                         // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_conv(s));
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {                                               // ONE thread runs the whole issue loop
      const uint32_t idesc = umma_idesc_tf32(TC_BM, bn);
      const bool three = p.nterms == 3;
      mbar_wait_one(bar_w, 0);
      int s = 0; uint32_t ph = 0;
      int j = 0; uint32_t dph = 0;
      for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
        mbar_wait_one(bar_dempty(j), dph ^ 1);                      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(j * bn);
        for (int c = 0; c < kch; ++c) {
          mbar_wait_one(bar_conv(s), ph);
          tc_fence_after();
          const uint32_t a_hi = stage0 + s * 2 * TC_CHUNK_BYTES;
          const uint64_t dah0 = umma_desc_k_sw128(a_hi), dal0 = umma_desc_k_sw128(a_hi + TC_CHUNK_BYTES);
          const uint64_t dbh0 = umma_desc_k_sw128(whi + c * w_chunk_bytes), dbl0 = umma_desc_k_sw128(wlo + c * w_chunk_bytes);
#pragma unroll
          for (int k = 0; k < TC_KC / 8; ++k) {                      // UMMA K = 8 tf32 = 32 bytes along the swizzled row = +2 in the address field
            umma_tf32(d_tmem, dah0 + 2u * k, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
            if (three) {
              umma_tf32(d_tmem, dah0 + 2u * k, dbl0 + 2u * k, idesc, 1u);
              umma_tf32(d_tmem, dal0 + 2u * k, dbh0 + 2u * k, idesc, 1u);
            }
          }
          umma_commit(bar_empty(s));                                 // stage reusable once these MMAs retire
          if (c == kch - 1) umma_commit(bar_dfull(j));               // accumulator complete
          if (++s == nst) { s = 0; ph ^= 1; }
        }
        if (++j == 2) { j = 0; dph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_EPI_WARPS) {
    // ===================================== epilogue ======================================
    tc_epilogue<EPI>(p, tmem_base, bn, bar_dfull(0), bar_dempty(0), bias_s, worker, split, warp - TC_EPI_WARP0, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}


// ---------------------------------------------------------------------------------------------------------------------------
// TS variant: the A operand lives in TMEM.
// Shared memory then holds only the resident weight slice and a deep ring of RAW 16 KB chunks (TMA destinations): a converter thread
// owns one row (= one TMEM lane), reads its 32 floats of the chunk from the swizzled tile, releases the ring slot immediately,
// applies the optional LayerNorm, splits into tf32 hi / lo and writes both with tcgen05.st into one of TS_ASTAGES TMEM operand
// stages (64 columns each); the MMA thread issues D += A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T with A read from TMEM.
// Bytes in flight per SM: nstages x 16 KB (6-8 stages) instead of 3-4, and the ring slot turns around without waiting for the MMAs.
// TMEM columns: [0, 2*bn) accumulators, then TS_ASTAGES x [hi 32 | lo 32].
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int TS_MAX_STAGES = 8;
constexpr int TS_ASTAGES = 4;

// fp16 split (F16 = true).  kind::f16 runs at twice the kind::tf32 rate and its operands take half the shared memory / TMEM, and an fp16 significand
// has the same 11 bits as a tf32 one, so  a S_a = a_hi + a_lo,  w S_w = w_hi + w_lo  (hi = fp16(x), lo = fp16(x - hi), products exact in the fp32
// accumulator) gives the 3xTF32 accuracy -- PROVIDED both halves stay in fp16's normal range, which tf32 (8 exponent bits) gave for free.  Hence the
// power-of-two scales: S_w puts max |w| at 2^14 (chosen on the device by the weight preparation), S_a is 2^10 behind a LayerNorm (|a| <= sqrt(K))
// and 2^4 otherwise; 1 / (S_a S_w) rides in the epilogue's bias FMA.  Elements below 2^-12 / S_a lose relative (not absolute) accuracy: absolute
// error <= 2^-25 / S_a per element (tools/split_error_study.py: 5e-8 of sum |a||w| at unit scale like 3xTF32, 7e-7 for activations of 1e-3).
// |a| S_a > 65504 would overflow to inf: the converter raises a sticky flag (eigb200_gemm_overflow) and the caller reruns with 3xTF32.
template <int EPI, bool DEFER, int AST = TS_ASTAGES, bool F16 = false>   // AST operand stages in TMEM (2 for the wide single-CTA tf32 plan: 2 * bn + 64 * AST <= 512 columns)
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_ts_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapWhi,
                  const __grid_constant__ CUtensorMap tmapWlo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int bn = p.bn, kch = p.kchunks, nst = p.nstages;
  const int kchw = F16 ? p.kch_w : kch;                              // weight chunks: bn rows x 128 bytes = 32 tf32 or 64 fp16 K-values
  constexpr uint32_t ACOLS = F16 ? 32u : 64u;                        // TMEM columns of one operand stage [hi | lo]
  const uint32_t w_chunk_bytes = (uint32_t)bn * 128u;
  const uint32_t whi = base;
  const uint32_t wlo = whi + kchw * w_chunk_bytes;
  const uint32_t stage0 = wlo + kchw * w_chunk_bytes;                // stage s: raw 16 KB chunk
  const uint32_t epi_stage0 = stage0 + nst * TC_CHUNK_BYTES;         // optional staging tiles of the epilogue warps (p.epi_stage bytes)
  const uint32_t bars = epi_stage0 + (uint32_t)p.epi_stage;
  const uint32_t bar_w = bars;
  auto bar_full = [&](int s) { return bars + 8u * (1 + s); };                              // TMA landed the raw chunk
  auto bar_free = [&](int s) { return bars + 8u * (1 + TS_MAX_STAGES + s); };              // converters have read it
  auto bar_afull = [&](int t) { return bars + 8u * (1 + 2 * TS_MAX_STAGES + t); };         // TMEM operand stage written
  auto bar_aempty = [&](int t) { return bars + 8u * (1 + 2 * TS_MAX_STAGES + TS_ASTAGES + t); };   // MMAs reading it retired
  auto bar_dfull = [&](int j) { return bars + 8u * (1 + 2 * TS_MAX_STAGES + 2 * TS_ASTAGES + j); };
  auto bar_dempty = [&](int j) { return bars + 8u * (3 + 2 * TS_MAX_STAGES + 2 * TS_ASTAGES + j); };
  const uint32_t tmem_slot = bars + 8u * (5 + 2 * TS_MAX_STAGES + 2 * TS_ASTAGES);          // slot 29 -> byte 232 (< 256)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = blockIdx.x / p.nsplit, split = blockIdx.x - worker * p.nsplit;
  const uint32_t a_col0 = (uint32_t)(2 * bn);                        // first TMEM column of the operand stages
  const uint32_t need_cols = a_col0 + AST * ACOLS;
  const uint32_t tmem_cols = need_cols <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < nst; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_free(s), 128); }
    for (int t = 0; t < AST; ++t) { mbar_init(bar_afull(t), 128); mbar_init(bar_aempty(t), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_dfull(j), 1); mbar_init(bar_dempty(j), TC_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (threadIdx.x < 128) {
    constexpr bool GLU_ = EPI == EIGB200_EPI_GLU_RESIDUAL;
    const int nout_ = GLU_ ? p.N / 2 : p.N;
    for (int c = threadIdx.x; c < bn; c += 128) {
      int n;
      if (GLU_) n = glu_weight_row(c, split, p.bg, nout_);
      else { n = split * bn + c; if (n >= p.N) n = -1; }
      float bv = (p.bias && n >= 0) ? p.bias[n] : 0.f;
      if (GLU_ && (c & 16)) bv *= -1.4426950408889634f;            // gate columns of each 32-column accumulator group: tc_epilogue folds the scale into an FMA
      bias_s[c] = bv;
    }
  }
  if (EPI == EIGB200_EPI_GLU_RESIDUAL && p.eig_part && threadIdx.x >= 128 && threadIdx.x < 256) {   // gate weights of this CTA's bg output columns
    float* ew = bias_s + 192;
    for (int c = threadIdx.x - 128; c < p.bg; c += 128) { const int n = split * p.bg + c; ew[c] = n < p.N / 2 ? p.eig_w[n] : 0.f; }
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  if (warp == TC_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmapA); tma_prefetch_desc(&tmapWhi); tma_prefetch_desc(&tmapWlo); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == TC_TMA_WARP) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, 2u * kchw * w_chunk_bytes);
      for (int c = 0; c < kchw; ++c) {
        tma_load_2d(&tmapWhi, bar_w, whi + c * w_chunk_bytes, c * (F16 ? 64 : TC_KC), split * bn);
        tma_load_2d(&tmapWlo, bar_w, wlo + c * w_chunk_bytes, c * (F16 ? 64 : TC_KC), split * bn);
      }
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
        for (int c = 0; c < kch; ++c) {
          mbar_wait_one(bar_free(s), ph ^ 1);
          mbar_arrive_expect_tx(bar_full(s), TC_CHUNK_BYTES);
          tma_load_2d(&tmapA, bar_full(s), stage0 + s * TC_CHUNK_BYTES, c * TC_KC, (int)(tile * TC_BM));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= TC_CONV_WARP0 && warp < TC_CONV_WARP0 + 4) {
    // ===================================== converters: one row (TMEM lane) per thread ======================================
    const int r = threadIdx.x - TC_CONV_WARP0 * 32;                                 // row of the tile = TMEM lane; warp % 4 == r / 32 (lane quarter)
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const bool ln = p.ln_stats != nullptr;
    const uint32_t row_off = (uint32_t)r * 128u;
    const int sw = r & 7;
    int s = 0; uint32_t ph = 0;
    int t = 0; uint32_t aph = 0;
    int t_pending = -1;                                              // operand stage whose tcgen05.st are issued but not yet published
    for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
      // NOTE: this loop is latency-critical and its performance follows ptxas' schedule: prefetching the statistics one tile ahead, making
      // `ln` a template parameter or forcing the publish ahead of the split with a data dependency were each measured 10-18 % SLOWER on the
      // plain N-split shapes (A/B on the same GPU), for a 2 % gain on the LayerNorm shape.
      float2 st = make_float2(0.f, 1.f);
      if (ln) { const int64_t m = tile * TC_BM + r; st = m < p.M ? __ldg(p.ln_stats + m) : make_float2(0.f, 0.f); st.x = -st.x * st.y; }
      if (F16) { st.x *= p.a_scale; st.y *= p.a_scale; }             // S_a folded into the (a - mu) rstd FMA; without LayerNorm st = (0, S_a)
      float amax = 0.f;
      for (int c = 0; c < kch; ++c) {
        mbar_wait(bar_full(s), ph);
#ifdef EIGB_ABL_CONV                                                 // ablation build: barrier protocol only, no loads / split / tcgen05.st
        if (DEFER && t_pending >= 0) { tc_fence_before(); mbar_arrive(bar_afull(t_pending)); }
        mbar_arrive(bar_free(s));
        mbar_wait(bar_aempty(t), aph ^ 1);
        if (DEFER) t_pending = t; else { tc_fence_before(); mbar_arrive(bar_afull(t)); }
        if (++s == nst) { s = 0; ph ^= 1; }
        if (++t == AST) { t = 0; aph ^= 1; }
        continue;
#endif
        const float4* src = reinterpret_cast<const float4*>(smem_raw + (stage0 + s * TC_CHUNK_BYTES + row_off - smem_u32(smem_raw)));
        float a[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {                                // logical 16-byte slot q sits at physical slot q ^ (row & 7)
          const float4 v = src[q ^ sw];
          a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
        if (DEFER && t_pending >= 0) {                               // publish the PREVIOUS chunk while this chunk's shared-memory loads are in flight
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_afull(t_pending));
        }
        if constexpr (F16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], st.y, st.x);
          uint32_t hi2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            hi2[i] = pack_f16x2(a[2 * i], a[2 * i + 1]);
            amax = fmaxf(amax, fmaxf(fabsf(a[2 * i]), fabsf(a[2 * i + 1])));
          }
          {                                                          // the raw slot can be refilled: its values are in registers
            uint32_t dep = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) dep ^= hi2[2 * q];
            mbar_arrive_after(bar_free(s), dep, (uint32_t)p.zero);
          }
          mbar_wait(bar_aempty(t), aph ^ 1);                         // MMAs that read this operand stage have retired
          tc_fence_after();
          const uint32_t acol = tmem_base + lane_sel + a_col0 + (uint32_t)t * ACOLS;
          tmem_st_32x16(acol, hi2);
          if (p.nterms == 3) {
            uint32_t lo2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) lo2[i] = pack_f16x2(a[2 * i] - f16_lo_to_f32(hi2[i]), a[2 * i + 1] - f16_hi_to_f32(hi2[i]));
            tmem_st_32x16(acol + 16u, lo2);
          }
          if (DEFER) t_pending = t;
          else { tmem_st_wait(); tc_fence_before(); mbar_arrive(bar_afull(t)); }
          if (++s == nst) { s = 0; ph ^= 1; }
          if (++t == AST) { t = 0; aph ^= 1; }
          continue;
        }
        if (ln) {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], st.y, st.x);
        }
        float hi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) hi[i] = __uint_as_float((__float_as_uint(a[i]) + 0x1000u) & 0xffffe000u);
        {                                                            // the raw slot can be refilled: its values are in registers
          uint32_t dep = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) dep ^= __float_as_uint(hi[4 * q]);
          mbar_arrive_after(bar_free(s), dep, (uint32_t)p.zero);
        }
        mbar_wait(bar_aempty(t), aph ^ 1);                           // MMAs that read this operand stage have retired
        tc_fence_after();
        const uint32_t acol = tmem_base + lane_sel + a_col0 + (uint32_t)t * 64u;
        tmem_st_32x32(acol, hi);
        if (p.nterms == 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = a[i] - hi[i];
          tmem_st_32x32(acol + 32u, a);
        }
        if (DEFER) t_pending = t;                                    // publish behind the next chunk's loads (converter-bound shapes)
        else { tmem_st_wait(); tc_fence_before(); mbar_arrive(bar_afull(t)); }
        if (++s == nst) { s = 0; ph ^= 1; }
        if (++t == AST) { t = 0; aph ^= 1; }
      }
      if (F16 && !(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);        // |a| S_a beyond fp16 (or NaN): the result of this tile is inf / NaN, say so
    }
    if (DEFER && t_pending >= 0) { tmem_st_wait(); tc_fence_before(); mbar_arrive(bar_afull(t_pending)); }
  } else if (warp == TC_MMA_WARP) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {                                               // ONE thread runs the whole issue loop
      const uint32_t idesc = F16 ? umma_idesc_f16(TC_BM, bn) : umma_idesc_tf32(TC_BM, bn);
      const bool three = p.nterms == 3;
      mbar_wait_one(bar_w, 0);
      int t = 0; uint32_t aph = 0;
      int j = 0; uint32_t dph = 0;
      for (int64_t tile = worker; tile < p.ntiles; tile += p.workers) {
        mbar_wait_one(bar_dempty(j), dph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(j * bn);
        for (int c = 0; c < kch; ++c) {
          mbar_wait_one(bar_afull(t), aph);
          tc_fence_after();
          if constexpr (F16) {
            // activation chunk c = 32 K-values = half of weight chunk c / 2 (64 fp16 per 128-byte swizzle row): +64 bytes = +4 in the address field
            const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)t * ACOLS, a_lo = a_hi + 16u;
            const uint32_t wofs = (uint32_t)(c >> 1) * w_chunk_bytes;
            const uint64_t kofs = (c & 1) ? 4u : 0u;
            const uint64_t dbh0 = umma_desc_k_sw128(whi + wofs) + kofs, dbl0 = umma_desc_k_sw128(wlo + wofs) + kofs;
#pragma unroll
            for (int k = 0; k < 2; ++k) {                            // UMMA K = 16 fp16 = 32 bytes of the swizzled row, 8 TMEM columns of the operand
              umma_f16_ts(d_tmem, a_hi + 8u * k, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
              if (three) {
                umma_f16_ts(d_tmem, a_hi + 8u * k, dbl0 + 2u * k, idesc, 1u);
                umma_f16_ts(d_tmem, a_lo + 8u * k, dbh0 + 2u * k, idesc, 1u);
              }
            }
            umma_commit(bar_aempty(t));
            if (c == kch - 1) umma_commit(bar_dfull(j));
            if (++t == AST) { t = 0; aph ^= 1; }
            continue;
          }
          const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)t * 64u, a_lo = a_hi + 32u;
          const uint64_t dbh0 = umma_desc_k_sw128(whi + c * w_chunk_bytes), dbl0 = umma_desc_k_sw128(wlo + c * w_chunk_bytes);
#pragma unroll
          for (int k = 0; k < TC_KC / 8; ++k) {                      // +32 bytes along the swizzled row = +2 in the descriptor's address field
            umma_tf32_ts(d_tmem, a_hi + 8u * k, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
            if (three) {
              umma_tf32_ts(d_tmem, a_hi + 8u * k, dbl0 + 2u * k, idesc, 1u);
              umma_tf32_ts(d_tmem, a_lo + 8u * k, dbh0 + 2u * k, idesc, 1u);
            }
          }
          umma_commit(bar_aempty(t));
          if (c == kch - 1) umma_commit(bar_dfull(j));
          if (++t == AST) { t = 0; aph ^= 1; }
        }
        if (++j == 2) { j = 0; dph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_EPI_WARPS) {
    tc_epilogue<EPI, F16>(p, tmem_base, bn, bar_dfull(0), bar_dempty(0), bias_s, worker, split, warp - TC_EPI_WARP0, lane,
                          p.epi_stage ? reinterpret_cast<float*>(smem_raw + (epi_stage0 - smem_u32(smem_raw))) : nullptr);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Streamed-operand variant for the shapes the resident-weight kernels cannot take (K > 256 or a weight slice that does not fit shared memory:
// the LM-sized layers of BASELINE C5, d_model 512).  There the GEMM is compute-bound (2 N K / (K + N) flop per byte > 500), so the classic
// pipeline is the right shape: A is split ONCE into tf32 hi / lo in global memory (split_a_kernel, which also applies the LayerNorm row
// statistics), and a ring of stages [A_hi | A_lo | W_hi | W_lo] x 32 K-columns is filled by one TMA thread; one thread issues the 12
// tcgen05.mma per stage (both operands from shared memory); the 16 epilogue warps are the same as above.  CTAs walk (row tile, N tile) pairs with
// the N tile fastest, so the A tile of a row block is re-read from L2.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void split_a_kernel(const float* __restrict__ A, int64_t lda, const float2* __restrict__ stats, float* __restrict__ hi, float* __restrict__ lo,
                               int64_t M, int K, int kpad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;             // one float4 of the padded (M, kpad) array
  const int kq = kpad >> 2;
  if (idx >= M * kq) return;
  const int64_t m = idx / kq;
  const int k = (int)(idx - m * kq) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) {                                                                     // K % 4 == 0
    a = ldg_stream_f4(reinterpret_cast<const float4*>(A + m * lda + k));
    if (stats) {
      const float2 st = __ldg(stats + m);
      const float c = -st.x * st.y;
      a.x = fmaf(a.x, st.y, c); a.y = fmaf(a.y, st.y, c); a.z = fmaf(a.z, st.y, c); a.w = fmaf(a.w, st.y, c);
    }
  }
  float4 h, l;
  split_tf32_4(a, h, l);
  reinterpret_cast<float4*>(hi)[idx] = h;
  reinterpret_cast<float4*>(lo)[idx] = l;
}

// fp16 split of A for the streamed-operand kernel: a' = LN(a) S_a -> hi = fp16(a'), lo = fp16(a' - hi), rows of kp64 halfs (zero padded); raises the sticky
// overflow flag when |a'| leaves the fp16 range (the caller reruns with 3xTF32)
__global__ void split_a_f16_kernel(const float* __restrict__ A, int64_t lda, const float2* __restrict__ stats, __half* __restrict__ hi, __half* __restrict__ lo,
                                   int64_t M, int K, int kp64, float a_scale, int* __restrict__ ovf_flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;             // four halfs of the padded (M, kp64) array
  const int kq = kp64 >> 2;
  if (idx >= M * kq) return;
  const int64_t m = idx / kq;
  const int k = (int)(idx - m * kq) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) {                                                                     // K % 4 == 0
    a = ldg_stream_f4(reinterpret_cast<const float4*>(A + m * lda + k));
    float sc = a_scale, c = 0.f;
    if (stats) { const float2 st = __ldg(stats + m); sc = st.y * a_scale; c = -st.x * st.y * a_scale; }
    a.x = fmaf(a.x, sc, c); a.y = fmaf(a.y, sc, c); a.z = fmaf(a.z, sc, c); a.w = fmaf(a.w, sc, c);
  }
  const uint32_t h01 = pack_f16x2(a.x, a.y), h23 = pack_f16x2(a.z, a.w);
  const uint32_t l01 = pack_f16x2(a.x - f16_lo_to_f32(h01), a.y - f16_hi_to_f32(h01)), l23 = pack_f16x2(a.z - f16_lo_to_f32(h23), a.w - f16_hi_to_f32(h23));
  reinterpret_cast<uint2*>(hi)[idx] = make_uint2(h01, h23);
  reinterpret_cast<uint2*>(lo)[idx] = make_uint2(l01, l23);
  const float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
  if (!(amax <= 65504.f)) atomicOr(ovf_flag, 1);
}

// bias in accumulator-column order (nsplit * bn entries, zero for padding columns)
__global__ void perm_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int N, int bn, int bg, int nsplit, int glu) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nsplit * bn) return;
  const int split = idx / bn, local = idx - split * bn;
  int n;
  if (glu) n = glu_weight_row(local, split, bg, N / 2);
  else { n = split * bn + local; if (n >= N) n = -1; }
  out[idx] = (bias && n >= 0) ? bias[n] : 0.f;
}

constexpr int SK_MAX_STAGES = 4;

__device__ __forceinline__ float4 sk_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sk_sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// F16: operands are the scaled fp16 split (kind::f16; a stage row of 128 bytes holds 64 K values instead of 32: half the L2 traffic per product, twice the
// tensor rate), the accumulator carries S_a S_w and the epilogue's bias FMA removes it.
// RAWA (F16 only): A arrives as RAW fp32 -- tmapAhi is the fp32 map of A (row stride lda), two 32-column boxes per stage in the 32 KB the hi / lo operands
// took -- and the four converter warps rewrite every landed row IN PLACE as [fp16 hi of its 32 values (64 B) | fp16 lo (64 B)] (thread = row; the box stays a
// SWIZZLE_128B K-major tile whose first two 32-byte K steps are the hi operand and whose last two are the lo operand, as in k7_front_fused.cu), LayerNorm
// folded in.  That removes split_a_f16_kernel -- A read, written as hi / lo and read again: a quarter of every call at K = 512 -- and its 2 M K workspace bytes.
template <int EPI, bool F16 = false, bool RAWA = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_stream_kernel(const __grid_constant__ CUtensorMap tmapAhi, const __grid_constant__ CUtensorMap tmapAlo,
                      const __grid_constant__ CUtensorMap tmapWhi, const __grid_constant__ CUtensorMap tmapWlo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int bn = p.bn, kch = p.kchunks, nst = p.nstages;
  const uint32_t w_chunk_bytes = (uint32_t)bn * 128u;
  const uint32_t stage_bytes = RAWA ? 2u * w_chunk_bytes : 2u * TC_CHUNK_BYTES + 2u * w_chunk_bytes;   // [A_hi 16K][A_lo 16K][W_hi][W_lo]; RAWA: [W_hi][W_lo]
  // RAWA: the raw A boxes (32 K columns, 16 KB) have their own, deeper ring in front of the W ring: a box goes TMA -> converters -> MMA, the weights only
  // TMA -> MMA, and with both in one 96 KB stage (two stages at 256 columns) the conversion sat in the critical path (tensor pipe 55 %)
  const int na = RAWA ? p.na_stages : 0;
  const uint32_t aring = base;
  const uint32_t stage0 = base + (uint32_t)na * TC_CHUNK_BYTES;
  const uint32_t bars = stage0 + nst * stage_bytes;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (SK_MAX_STAGES + s); };
  auto bar_dfull = [&](int j) { return bars + 8u * (2 * SK_MAX_STAGES + j); };
  auto bar_dempty = [&](int j) { return bars + 8u * (2 * SK_MAX_STAGES + 2 + j); };
  const uint32_t tmem_slot = bars + 8u * (2 * SK_MAX_STAGES + 4);
  auto bar_araw = [&](int i) { return bars + 8u * (24 + i); };      // RAWA: TMA landed raw box i | converted (128 arrivals) | its MMAs retired
  auto bar_afull = [&](int i) { return bars + 8u * (32 + i); };
  auto bar_aempty = [&](int i) { return bars + 8u * (40 + i); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * bn <= 32) ? 32 : (2 * bn <= 64) ? 64 : (2 * bn <= 128) ? 128 : (2 * bn <= 256) ? 256 : 512;
  const int64_t npairs = p.ntiles * p.nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int i = 0; i < na; ++i) { mbar_init(bar_araw(i), 1); mbar_init(bar_afull(i), 128); mbar_init(bar_aempty(i), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_dfull(j), 1); mbar_init(bar_dempty(j), TC_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
  if (warp == TC_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmapAhi); tma_prefetch_desc(&tmapAlo); tma_prefetch_desc(&tmapWhi); tma_prefetch_desc(&tmapWlo); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == TC_TMA_WARP) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      int ai = 0; uint32_t aph = 0;
      for (int64_t it = blockIdx.x; it < npairs; it += gridDim.x) {
        const int64_t tile = it / p.nsplit;
        const int split = (int)(it - tile * p.nsplit);
        for (int c = 0; c < kch; ++c) {
          constexpr int KSTAGE = F16 ? 64 : TC_KC;                 // K values per 128-byte row
          if constexpr (RAWA) {                                    // two raw fp32 boxes: K columns [64 c, 64 c + 32) and [64 c + 32, 64 c + 64) (zero-filled beyond K)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              mbar_wait_one(bar_aempty(ai), aph ^ 1);
              mbar_arrive_expect_tx(bar_araw(ai), TC_CHUNK_BYTES);
              tma_load_2d(&tmapAhi, bar_araw(ai), aring + (uint32_t)ai * TC_CHUNK_BYTES, c * KSTAGE + b * TC_KC, (int)(tile * TC_BM));
              if (++ai == na) { ai = 0; aph ^= 1; }
            }
          }
          mbar_wait_one(bar_empty(s), ph ^ 1);
          const uint32_t st0 = stage0 + s * stage_bytes;
          mbar_arrive_expect_tx(bar_full(s), stage_bytes);
          if constexpr (RAWA) {
            tma_load_2d(&tmapWhi, bar_full(s), st0, c * KSTAGE, split * bn);
            tma_load_2d(&tmapWlo, bar_full(s), st0 + w_chunk_bytes, c * KSTAGE, split * bn);
          } else {
            tma_load_2d(&tmapAhi, bar_full(s), st0, c * KSTAGE, (int)(tile * TC_BM));
            tma_load_2d(&tmapAlo, bar_full(s), st0 + TC_CHUNK_BYTES, c * KSTAGE, (int)(tile * TC_BM));
            tma_load_2d(&tmapWhi, bar_full(s), st0 + 2 * TC_CHUNK_BYTES, c * KSTAGE, split * bn);
            tma_load_2d(&tmapWlo, bar_full(s), st0 + 2 * TC_CHUNK_BYTES + w_chunk_bytes, c * KSTAGE, split * bn);
          }
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == TC_MMA_WARP) {
    if (elect_one()) {
      const uint32_t idesc = F16 ? umma_idesc_f16(TC_BM, bn) : umma_idesc_tf32(TC_BM, bn);
      const bool three = p.nterms == 3;
      int s = 0; uint32_t ph = 0;
      int ai = 0; uint32_t aph = 0;
      int j = 0; uint32_t dph = 0;
      for (int64_t it = blockIdx.x; it < npairs; it += gridDim.x) {
        mbar_wait_one(bar_dempty(j), dph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(j * bn);
        for (int c = 0; c < kch; ++c) {
          mbar_wait_one(bar_full(s), ph);
          tc_fence_after();
          const uint32_t st0 = stage0 + s * stage_bytes;
          if constexpr (F16 && RAWA) {
            const uint64_t dbh0 = umma_desc_k_sw128(st0), dbl0 = umma_desc_k_sw128(st0 + w_chunk_bytes);
#pragma unroll
            for (int b = 0; b < 2; ++b) {                            // box b: [hi K steps 0, 1 | lo K steps 0, 1] of K values [64 c + 32 b, + 32)
              mbar_wait_one(bar_afull(ai), aph);
              tc_fence_after();
              const uint64_t da = umma_desc_k_sw128(aring + (uint32_t)ai * TC_CHUNK_BYTES);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const uint32_t k = 2u * b + jj;
                umma_f16_ss(d_tmem, da + 2u * jj, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
                umma_f16_ss(d_tmem, da + 2u * jj, dbl0 + 2u * k, idesc, 1u);
                umma_f16_ss(d_tmem, da + 4u + 2u * jj, dbh0 + 2u * k, idesc, 1u);
              }
              umma_commit(bar_aempty(ai));
              if (++ai == na) { ai = 0; aph ^= 1; }
            }
          } else {
          const uint64_t dah0 = umma_desc_k_sw128(st0), dal0 = umma_desc_k_sw128(st0 + TC_CHUNK_BYTES);
          const uint64_t dbh0 = umma_desc_k_sw128(st0 + 2 * TC_CHUNK_BYTES), dbl0 = umma_desc_k_sw128(st0 + 2 * TC_CHUNK_BYTES + w_chunk_bytes);
#pragma unroll
          for (int k = 0; k < TC_KC / 8; ++k) {                      // 4 K steps of 32 bytes per stage row in either precision
            if constexpr (F16) {
              umma_f16_ss(d_tmem, dah0 + 2u * k, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
              if (three) {
                umma_f16_ss(d_tmem, dah0 + 2u * k, dbl0 + 2u * k, idesc, 1u);
                umma_f16_ss(d_tmem, dal0 + 2u * k, dbh0 + 2u * k, idesc, 1u);
              }
            } else {
              umma_tf32(d_tmem, dah0 + 2u * k, dbh0 + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
              if (three) {
                umma_tf32(d_tmem, dah0 + 2u * k, dbl0 + 2u * k, idesc, 1u);
                umma_tf32(d_tmem, dal0 + 2u * k, dbh0 + 2u * k, idesc, 1u);
              }
            }
          }
          }
          umma_commit(bar_empty(s));
          if (c == kch - 1) umma_commit(bar_dfull(j));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
        if (++j == 2) { j = 0; dph ^= 1; }
      }
    }
    __syncwarp();
  } else if (RAWA && warp >= TC_CONV_WARP0 && warp < TC_CONV_WARP0 + 4) {
    // ===================================== converters: one row of the tile per thread, in place ======================================
    const int r = threadIdx.x - TC_CONV_WARP0 * 32;
    const int sw = r & 7;
    const float S_a = p.ln_stats ? 1024.f : 16.f;                    // activation pre-scale (split_a_f16_kernel's)
    float amax = 0.f;
    int ai = 0; uint32_t aph = 0;
    for (int64_t it = blockIdx.x; it < npairs; it += gridDim.x) {
      const int64_t tile = it / p.nsplit;
      const int64_t m = tile * TC_BM + r;
      float sc = S_a, cc = 0.f;
      if (p.ln_stats && m < p.M) { const float2 stt = __ldg(p.ln_stats + m); sc = stt.y * S_a; cc = -stt.x * stt.y * S_a; }
      for (int cb = 0; cb < 2 * kch; ++cb) {                         // raw boxes in K order
        const uint32_t row = aring + (uint32_t)ai * TC_CHUNK_BYTES + (uint32_t)r * 128u;
        mbar_wait(bar_araw(ai), aph);
#ifndef SK_ABL_CONV                                                  // timing-only ablation: operands left as raw bytes (wrong results)
        // two half rows (16 values each): every slot is read before it is overwritten -- hi goes to logical slots 0-3, lo to 4-7 -- and the live set stays
        // at 48 registers (the whole row at once spilled under this kernel's 80-register cap)
        uint32_t hi2[16], lo2[16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float a[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {                              // logical 16-byte slot q sits at physical slot q ^ (row & 7)
            const float4 v = sk_lds_f4(row + (uint32_t)(((4 * hf + q) ^ sw) * 16));
            a[4 * q] = fmaf(v.x, sc, cc); a[4 * q + 1] = fmaf(v.y, sc, cc); a[4 * q + 2] = fmaf(v.z, sc, cc); a[4 * q + 3] = fmaf(v.w, sc, cc);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t h = pack_f16x2(a[2 * e], a[2 * e + 1]);
            hi2[8 * hf + e] = h;
            amax = fmaxf(amax, fmaxf(fabsf(a[2 * e]), fabsf(a[2 * e + 1])));
            lo2[8 * hf + e] = pack_f16x2(a[2 * e] - f16_lo_to_f32(h), a[2 * e + 1] - f16_hi_to_f32(h));
          }
          if (hf == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sk_sts_u4(row + (uint32_t)((j ^ sw) * 16), hi2[4 * j], hi2[4 * j + 1], hi2[4 * j + 2], hi2[4 * j + 3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sk_sts_u4(row + (uint32_t)(((4 + j) ^ sw) * 16), lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
#endif
        fence_proxy_async();                                         // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_afull(ai));
        if (++ai == na) { ai = 0; aph ^= 1; }
      }
    }
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_EPI_WARPS) {
    tc_epilogue_stream<EPI, true, F16>(p, tmem_base, bn, bar_dfull(0), bar_dempty(0), p.bias, 0, 0, warp - TC_EPI_WARP0, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor map: inner dimension `cols` (contiguous), outer `rows` with row stride ld (elements); box = 32 x box_rows, SWIZZLE_128B
static int make_tmap(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EIGB200_ECUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_KC, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", (int)r,
                                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows); return EIGB200_ECUDA; }
  return EIGB200_OK;
}

// 2-D fp16 tensor map of the split weights: inner dimension `cols` halfs, box = 64 halfs (one 128-byte swizzle row) x box_rows
static int make_tmap_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EIGB200_ECUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp16 weights) failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
                                     (unsigned long long)rows, (unsigned long long)cols, box_rows); return EIGB200_ECUDA; }
  return EIGB200_OK;
}

// ---- operand precision of the resident-weight kernels ---------------------------------------------------------------------------------
// 0 = 3xTF32 (kind::tf32), 1 = fp16 split (kind::f16).  EIGB200_GEMM_PRECISION = tf32x3 | f16x3 selects what EIGB200_GEMM_AUTO, eigb200_linear_ln,
// eigb200_linear_glu_extract and eigb200_linear_prepare use; eigb200_linear(mode = TC_3XTF32 / TC_F16X3) asks for one explicitly.
static int g_kind_override = -1;
void tc_set_default_kind(int kind) { g_kind_override = (kind == 0 || kind == 1) ? kind : -1; }
int tc_default_kind() {
  static int v = -1;
  if (g_kind_override >= 0) return g_kind_override;
  if (v < 0) { const char* e = getenv("EIGB200_GEMM_PRECISION"); v = (e && e[0] == 'f') ? 1 : (e && e[0] == 't') ? 0 : EIGB200_GEMM_DEFAULT_KIND; }
  return v;
}
__device__ int g_gemm_overflow = 0;
static int* overflow_flag_ptr() {
  int* ptr = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&ptr), g_gemm_overflow) != cudaSuccess) return nullptr;   // per device (cudaSetDevice'd by the caller)
  return ptr;
}
int tc_overflow_query(cudaStream_t st, int reset, int* h_flag) {
  int* d = overflow_flag_ptr();
  if (!d) { set_error("gemm_overflow: cannot resolve the device flag"); return EIGB200_ECUDA; }
  EIGB_CUDA(cudaMemcpyAsync(h_flag, d, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (reset) EIGB_CUDA(cudaMemsetAsync(d, 0, sizeof(int), st));
  EIGB_CUDA(cudaStreamSynchronize(st));
  return EIGB200_OK;
}

struct TcPlan { int bn, bg, nsplit, kchunks, kpad, nstages, ast; size_t smem; bool ok; bool ts; int kp64, kch_w; };

static bool use_ts_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EIGB200_GEMM_VARIANT"); v = (e && e[0] == 's' && e[1] == 's') ? 0 : 1; }   // default: TMEM-operand kernel; "ss" selects the smem-operand kernel
  return v == 1;
}

static int gemm_wide_mode() {       // EIGB200_GEMM_WIDE: 0 = N-split plan only, 1 = wide plan (default), 2 = wide plan with the deferred publish
  static int v = -1;
  if (v < 0) { const char* e = getenv("EIGB200_GEMM_WIDE"); v = e ? atoi(e) : 1; }
  return v;
}

static TcPlan make_plan(int N, int K, int epilogue, int kind = 0) {
  TcPlan pl{};
  pl.ok = false;
  pl.ts = use_ts_variant() || kind == 1;                              // the fp16 split exists in the TMEM-operand kernel only
  if (K % 4 != 0 || K <= 0) return pl;
  pl.kpad = (K + TC_KC - 1) / TC_KC * TC_KC;
  pl.kchunks = pl.kpad / TC_KC;
  pl.kp64 = (K + 63) / 64 * 64;
  pl.kch_w = pl.kp64 / 64;
  if (pl.kchunks > 8) return pl;
  const bool glu = epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const int nout = glu ? N / 2 : N;
  pl.ast = TS_ASTAGES;
  const int wchunks = kind == 1 ? pl.kch_w : pl.kchunks;             // 128-byte weight chunks per row and per hi / lo
  // Wide single-CTA plan (129..192 output columns, e.g. the Mamba in_proj with N = 161): the whole weight matrix stays resident in ONE CTA
  // (bn = N rounded up to 16), so every A tile is read and converted once instead of once per N split; what it costs is ring depth
  // (the weights leave room for 2-3 raw stages) and TMEM operand stages (2 * bn accumulator columns leave room for 2).
  if (pl.ts && !glu && nout > 128 && nout <= 192 && gemm_wide_mode() != 0) {
    const int bn = (nout + 15) / 16 * 16;
    const size_t wbytes = (size_t)2 * wchunks * bn * 128;
    const long room = (long)TC_SMEM_LIMIT - 2048 - (long)wbytes;
    int nst = room > 0 ? (int)(room / TC_CHUNK_BYTES) : 0;
    if (nst > TS_MAX_STAGES) nst = TS_MAX_STAGES;
    if (nst >= 2 && 2 * bn + 2 * 64 <= 512) {
      pl.bn = pl.bg = bn; pl.nsplit = 1; pl.nstages = nst; pl.ast = kind == 1 ? TS_ASTAGES : 2;   // fp16 operand stages are 32 columns: 4 of them fit
      pl.smem = wbytes + (size_t)nst * TC_CHUNK_BYTES + 1024 + 1024;
      pl.ok = true;
      return pl;
    }
  }
  const int per_col_bytes = wchunks * 128 * 2 * (glu ? 2 : 1);       // hi+lo bytes per OUTPUT column
  const int stage_bytes = pl.ts ? TC_CHUNK_BYTES : 2 * TC_CHUNK_BYTES;
  const int max_w_bytes = TC_SMEM_LIMIT - 2048 - (pl.ts ? 4 : 2) * stage_bytes;   // keep room for the minimum ring
  int max_cols = max_w_bytes / per_col_bytes;
  const int cap = glu ? 64 : 128;
  if (max_cols > cap) max_cols = cap;
  const int gran = glu ? 16 : 32;                                    // bn must be a multiple of 32
  max_cols = max_cols / gran * gran;
  if (max_cols < gran) return pl;
  pl.nsplit = (nout + max_cols - 1) / max_cols;
  int cols = (nout + pl.nsplit - 1) / pl.nsplit;
  cols = (cols + gran - 1) / gran * gran;
  pl.bg = cols;
  pl.bn = glu ? 2 * cols : cols;
  const size_t wbytes = (size_t)2 * wchunks * pl.bn * 128;
  int nst = (int)((TC_SMEM_LIMIT - 2048 - (long)wbytes) / stage_bytes);
  if (nst > (pl.ts ? TS_MAX_STAGES : TC_MAX_STAGES)) nst = pl.ts ? TS_MAX_STAGES : TC_MAX_STAGES;
  if (nst < 2) return pl;
  pl.nstages = nst;
  pl.smem = wbytes + (size_t)nst * stage_bytes + 1024 /*alignment*/ + 1024 /*barriers + bias*/;
  pl.ok = true;
  return pl;
}

size_t tc_workspace_bytes(int N, int K);
// plan of the streamed-operand kernel (any K % 4 == 0): 128 accumulator columns per N tile (GLU: 64 value + 64 gate columns)
struct StreamPlan { int bn, bg, nsplit, kchunks, kpad, nstages; size_t smem; bool ok; };
static StreamPlan make_stream_plan(int N, int K, int epilogue) {
  StreamPlan pl{};
  pl.ok = false;
  if (K % 4 != 0 || K <= 0 || K > 16384) return pl;
  pl.kpad = (K + TC_KC - 1) / TC_KC * TC_KC;
  pl.kchunks = pl.kpad / TC_KC;
  const bool glu = epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const int nout = glu ? N / 2 : N;
  if (glu) { pl.bg = 64; pl.bn = 128; }
  else { pl.bn = nout >= 128 ? 128 : (nout + 31) / 32 * 32; pl.bg = pl.bn; }
  {
    // Wider N tiles: the A stage (32 KB per 64 K values, converted once per (row tile, N tile) pair) is shared by more products -- 48 instead of 64 KB from L2
    // per 128 x 128 x 64 block at 256 columns, which is what bounds this kernel (the L2 slices deliver ~43 B per clock and SM, the tensor core wants 80 at 128
    // columns).  Measured cost per PADDED output column at K = 512 (profiles/r2_stream_bn_sweep.txt): 1.00 / 0.88 / 0.76 at 128 / 192 / 256 columns; the plan
    // takes the width with the smallest padded width x cost (N 552 -> 3 x 192, N 1544 -> 7 x 256, N 512 / GLU 1024 -> 2 / 4 x 256).  EIGB200_STREAM_BN forces one.
    static const char* force = getenv("EIGB200_STREAM_BN");
    const int widths[3] = {128, 192, 256};
    const double cost[3] = {1.0, 0.88, 0.76};
    double best = 1e30;
    for (int i = 0; i < 3; ++i) {
      const int bn = widths[i], bg = glu ? bn / 2 : bn;
      if (nout < bg && i > 0) continue;
      if (force && atoi(force) != bn && nout >= bg) continue;
      if (!glu && i == 0 && nout < 128) { best = 0; break; }          // narrow outputs keep the rounded-up single tile chosen above
      const double c = (double)((nout + bg - 1) / bg * bg) * cost[i];
      if (c < best) { best = c; pl.bn = bn; pl.bg = bg; }
    }
  }
  pl.nsplit = (nout + pl.bg - 1) / pl.bg;
  const size_t stage_bytes = 2 * (size_t)TC_CHUNK_BYTES + 2 * (size_t)pl.bn * 128;
  int nst = (int)((TC_SMEM_LIMIT - 2048) / stage_bytes);
  if (nst > SK_MAX_STAGES) nst = SK_MAX_STAGES;
  if (nst < 2) return pl;
  pl.nstages = nst;
  pl.smem = (size_t)nst * stage_bytes + 1024 /*alignment*/ + 1024 /*barriers*/;
  pl.ok = true;
  return pl;
}

static size_t stream_workspace_bytes(int64_t M, int N, int K, int epilogue) {
  const StreamPlan pl = make_stream_plan(N, K, epilogue);
  if (!pl.ok) return 0;
  const size_t wrows = (size_t)pl.nsplit * pl.bn;
  return (2 * wrows * pl.kpad + (size_t)((N + 3) / 4 * 4) + wrows + 2 * (size_t)M * pl.kpad) * sizeof(float) + 64;   // + the fp16 form's scale block
}

size_t tc_workspace_bytes_m(int64_t M, int N, int K) {
  size_t best = tc_workspace_bytes(N, K);
  for (int epi : {EIGB200_EPI_NONE, EIGB200_EPI_GLU_RESIDUAL}) {
    if (epi == EIGB200_EPI_GLU_RESIDUAL && N % 2) continue;
    if (make_plan(N, K, epi).ok) continue;                          // the resident-weight kernel takes this shape
    const size_t b = stream_workspace_bytes(M, N, K, epi);
    if (b > best) best = b;
  }
  return best;
}

size_t tc_workspace_bytes(int N, int K) {
  // sized for the worst case of either epilogue family
  size_t best = 0;
  for (int epi : {EIGB200_EPI_NONE, EIGB200_EPI_GLU_RESIDUAL}) {
    if (epi == EIGB200_EPI_GLU_RESIDUAL && N % 2) continue;
    for (int kind = 0; kind < 2; ++kind) {
      TcPlan pl = make_plan(N, K, epi, kind);
      if (!pl.ok) continue;
      // tf32: hi + lo floats; fp16: hi + lo halfs of kp64 columns; then the folded LayerNorm bias and 4 scale words
      const size_t wb = kind == 1 ? (size_t)2 * pl.nsplit * pl.bn * pl.kp64 * 2 : (size_t)2 * pl.nsplit * pl.bn * pl.kpad * sizeof(float);
      const size_t b = wb + (size_t)((N + 3) / 4 * 4) * sizeof(float) + 16;
      if (b > best) best = b;
    }
  }
  return best;
}

bool tc_supported(const LinearParams& p) {
  if (p.lda % 4 != 0 || p.ldc % 4 != 0 || (p.R && p.ldr % 4 != 0)) return false;
  if (((uintptr_t)p.A & 15) || ((uintptr_t)p.C & 15) || (p.R && ((uintptr_t)p.R & 15))) return false;
  if (p.M >= (1LL << 31)) return false;
  return make_plan(p.N, p.K, p.epilogue).ok || make_stream_plan(p.N, p.K, p.epilogue).ok;
}

// Weight preparation of either plan: tf32 hi / lo split in the row order the CTAs consume (LayerNorm gamma folded in, GLU value / gate rows interleaved)
// and, with LayerNorm, the folded bias b + W beta behind it.  Run per call by launch_linear_tc, or once by eigb200_linear_prepare.
int tc_prepare(cudaStream_t st, const LinearParams& lp, void* workspace, int kind) {
  const bool glu = lp.epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const bool ln = lp.ln_stats != nullptr || lp.ln_gamma != nullptr;
  int bn, bg, nsplit, kpad;
  const TcPlan pl = make_plan(lp.N, lp.K, lp.epilogue, kind);
  if (pl.ok) { bn = pl.bn; bg = pl.bg; nsplit = pl.nsplit; kpad = pl.kpad; }
  else {
    const StreamPlan sp = make_stream_plan(lp.N, lp.K, lp.epilogue);
    if (!sp.ok) { set_error("tcgen05 GEMM: unsupported shape N=%d K=%d", lp.N, lp.K); return EIGB200_EUNSUPPORTED; }
    bn = sp.bn; bg = sp.bg; nsplit = sp.nsplit; kpad = sp.kpad;
  }
  const size_t wrows = (size_t)nsplit * bn;
  const int kp64 = (lp.K + 63) / 64 * 64;
  if (kind == 1) {
    // [hi halfs wrows x kp64][lo halfs][bias2 floats][scal: 1/(S_a S_w), max|w| bits, S_w, S_a]
    __half* w_hi = reinterpret_cast<__half*>(workspace);
    __half* w_lo = w_hi + wrows * kp64;
    float* bias2 = reinterpret_cast<float*>(w_lo + wrows * kp64);
    float* scal = bias2 + (lp.N + 3) / 4 * 4;
    const float a_scale = ln ? 1024.f : 16.f;
    EIGB_CUDA(cudaMemsetAsync(scal, 0, 16, st));
    const int nk = lp.N * lp.K;
    absmax_weights_kernel<<<(nk + 1023) / 1024 > 64 ? 64 : (nk + 1023) / 1024, 256, 0, st>>>(lp.W, ln ? lp.ln_gamma : nullptr, lp.N, lp.K, reinterpret_cast<unsigned*>(scal) + 1);
    EIGB_LAUNCH_CHECK("absmax_weights_kernel");
    const int total = (int)(wrows * kp64);
    split_weights_f16_kernel<<<(total + 255) / 256, 256, 0, st>>>(lp.W, w_hi, w_lo, lp.N, lp.K, kp64, bn, bg, nsplit, glu ? 1 : 0, ln ? lp.ln_gamma : nullptr, scal, a_scale);
    EIGB_LAUNCH_CHECK("split_weights_f16_kernel");
    if (ln) {
      ln_bias_kernel<<<(lp.N + 7) / 8, 256, 0, st>>>(lp.W, lp.bias, lp.ln_beta, bias2, lp.N, lp.K);
      EIGB_LAUNCH_CHECK("ln_bias_kernel");
    }
    return EIGB200_OK;
  }
  float* w_hi = reinterpret_cast<float*>(workspace);
  float* w_lo = w_hi + wrows * kpad;
  float* bias2 = w_lo + wrows * kpad;
  const int total = (int)(wrows * kpad);
  split_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(lp.W, w_hi, w_lo, lp.N, lp.K, kpad, bn, bg, nsplit, glu ? 1 : 0, ln ? lp.ln_gamma : nullptr);
  EIGB_LAUNCH_CHECK("split_weights_kernel");
  if (ln) {
    ln_bias_kernel<<<(lp.N + 7) / 8, 256, 0, st>>>(lp.W, lp.bias, lp.ln_beta, bias2, lp.N, lp.K);
    EIGB_LAUNCH_CHECK("ln_bias_kernel");
  }
  return EIGB200_OK;
}

// ---- helpers for the fused out_proj -> GLU kernel (k4_gemm_fused.cu) -------------------------------------------------------------------------
int tc_make_tmap_f32(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) { return make_tmap(map, ptr, rows, cols, ld, box_rows); }
int tc_make_tmap_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) { return make_tmap_f16(map, ptr, rows, cols, box_rows); }
int* tc_overflow_flag() { return overflow_flag_ptr(); }
bool tc_prepared_layout_f16(int N, int K, int epilogue, const void* ws, TcPrepared* out) {
  const TcPlan pl = make_plan(N, K, epilogue, 1);
  if (!pl.ok) return false;
  const size_t wrows = (size_t)pl.nsplit * pl.bn;
  const __half* w_hi = reinterpret_cast<const __half*>(ws);
  const __half* w_lo = w_hi + wrows * pl.kp64;
  out->w_hi = w_hi; out->w_lo = w_lo;
  out->bias2 = reinterpret_cast<const float*>(w_lo + wrows * pl.kp64);
  out->scal = out->bias2 + (N + 3) / 4 * 4;
  out->bn = pl.bn; out->bg = pl.bg; out->nsplit = pl.nsplit; out->kp64 = pl.kp64; out->kch_w = pl.kch_w; out->wrows = (int)wrows;
  return true;
}

static int launch_linear_stream(cudaStream_t st, const LinearParams& lp, int nterms, void* workspace, int kind) {
  if (lp.eig_part) { set_error("linear_glu_extract: not available on the streamed-operand kernel (K > 256)"); return EIGB200_EUNSUPPORTED; }
  const StreamPlan pl = make_stream_plan(lp.N, lp.K, lp.epilogue);
  if (!pl.ok) { set_error("tcgen05 GEMM: unsupported shape N=%d K=%d", lp.N, lp.K); return EIGB200_EUNSUPPORTED; }
  const bool glu = lp.epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const size_t wrows = (size_t)pl.nsplit * pl.bn;
  const bool f16 = kind == 1 && nterms == 3;
  const int kp64 = (lp.K + 63) / 64 * 64;
  // workspace: [W hi][W lo][bias2 = b + W beta][scal (fp16 split only)][bias in accumulator-column order][A hi][A lo]; the fp16 form needs at most the fp32 form's bytes
  float *bias2, *bias_perm; const float* scal = nullptr;
  void *w_hi, *w_lo, *a_hi, *a_lo;
  if (f16) {
    __half* wh = reinterpret_cast<__half*>(workspace);
    __half* wl = wh + wrows * kp64;
    bias2 = reinterpret_cast<float*>(wl + wrows * kp64);
    scal = bias2 + (lp.N + 3) / 4 * 4;
    bias_perm = bias2 + (lp.N + 3) / 4 * 4 + 4;
    __half* ah = reinterpret_cast<__half*>(bias_perm + wrows);
    w_hi = wh; w_lo = wl; a_hi = ah; a_lo = ah + (size_t)lp.M * kp64;
  } else {
    float* wh = reinterpret_cast<float*>(workspace);
    float* wl = wh + wrows * pl.kpad;
    bias2 = wl + wrows * pl.kpad;
    bias_perm = bias2 + (lp.N + 3) / 4 * 4;
    float* ah = bias_perm + wrows;
    w_hi = wh; w_lo = wl; a_hi = ah; a_lo = ah + (size_t)lp.M * pl.kpad;
  }
  if (lp.W != nullptr) {                                             // not prepared by eigb200_linear_prepare
    int rc0 = tc_prepare(st, lp, workspace, f16 ? 1 : 0);
    if (rc0 != EIGB200_OK) return rc0;
  }
  const float* bias_eff = lp.ln_stats ? bias2 : lp.bias;
  perm_bias_kernel<<<(unsigned)((wrows + 255) / 256), 256, 0, st>>>(bias_eff, bias_perm, lp.N, pl.bn, pl.bg, pl.nsplit, glu ? 1 : 0);
  EIGB_LAUNCH_CHECK("perm_bias_kernel");
  int* ovf = nullptr;
  static const bool rawa_off = getenv("EIGB200_STREAM_RAWA") != nullptr && atoi(getenv("EIGB200_STREAM_RAWA")) == 0;
  const bool rawa = f16 && !rawa_off;                                // A converted on the SM (gemm_tc_stream_kernel<EPI, true, true>): no split pre-pass
  if (f16 && rawa) {
    ovf = overflow_flag_ptr();
    if (!ovf) { set_error("tcgen05 GEMM: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  } else if (f16) {
    ovf = overflow_flag_ptr();
    if (!ovf) { set_error("tcgen05 GEMM: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
    const int64_t nq = lp.M * (kp64 / 4);
    split_a_f16_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(lp.A, lp.lda, reinterpret_cast<const float2*>(lp.ln_stats), reinterpret_cast<__half*>(a_hi),
                                                                    reinterpret_cast<__half*>(a_lo), lp.M, lp.K, kp64, lp.ln_stats ? 1024.f : 16.f, ovf);
    EIGB_LAUNCH_CHECK("split_a_f16_kernel");
  } else {
    const int64_t nq = lp.M * (pl.kpad / 4);
    split_a_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(lp.A, lp.lda, reinterpret_cast<const float2*>(lp.ln_stats), reinterpret_cast<float*>(a_hi),
                                                                reinterpret_cast<float*>(a_lo), lp.M, lp.K, pl.kpad);
    EIGB_LAUNCH_CHECK("split_a_kernel");
  }
  CUtensorMap tAh, tAl, tWh, tWl;
  int rc;
  if (f16) {
    if (rawa) {
      if ((rc = make_tmap(&tAh, lp.A, (uint64_t)lp.M, (uint64_t)lp.K, (uint64_t)lp.lda, TC_BM))) return rc;
      tAl = tAh;
    } else {
      if ((rc = make_tmap_f16(&tAh, a_hi, (uint64_t)lp.M, (uint64_t)kp64, TC_BM))) return rc;
      if ((rc = make_tmap_f16(&tAl, a_lo, (uint64_t)lp.M, (uint64_t)kp64, TC_BM))) return rc;
    }
    if ((rc = make_tmap_f16(&tWh, w_hi, wrows, (uint64_t)kp64, (uint32_t)pl.bn))) return rc;
    if ((rc = make_tmap_f16(&tWl, w_lo, wrows, (uint64_t)kp64, (uint32_t)pl.bn))) return rc;
  } else {
    if ((rc = make_tmap(&tAh, reinterpret_cast<float*>(a_hi), (uint64_t)lp.M, (uint64_t)pl.kpad, (uint64_t)pl.kpad, TC_BM))) return rc;
    if ((rc = make_tmap(&tAl, reinterpret_cast<float*>(a_lo), (uint64_t)lp.M, (uint64_t)pl.kpad, (uint64_t)pl.kpad, TC_BM))) return rc;
    if ((rc = make_tmap(&tWh, reinterpret_cast<float*>(w_hi), wrows, (uint64_t)pl.kpad, (uint64_t)pl.kpad, (uint32_t)pl.bn))) return rc;
    if ((rc = make_tmap(&tWl, reinterpret_cast<float*>(w_lo), wrows, (uint64_t)pl.kpad, (uint64_t)pl.kpad, (uint32_t)pl.bn))) return rc;
  }
  TcParams p{};
  p.bias = bias_perm; p.C = lp.C; p.ldc = lp.ldc; p.R = lp.R; p.ldr = lp.ldr; p.M = lp.M; p.N = lp.N; p.K = lp.K; p.epilogue = lp.epilogue;
  p.bn = pl.bn; p.bg = pl.bg; p.nsplit = pl.nsplit; p.kchunks = f16 ? kp64 / 64 : pl.kchunks; p.nstages = pl.nstages; p.nterms = nterms;
  p.ntiles = (lp.M + TC_BM - 1) / TC_BM;
  p.workers = 1; p.zero = 0; p.out_scale = scal; p.ovf_flag = ovf;
  size_t smem_bytes = pl.smem;
  if (rawa) {                                                        // separate rings: raw A boxes (16 KB each) in front of the [W hi | W lo] stages
    p.ln_stats = reinterpret_cast<const float2*>(lp.ln_stats);
    const size_t wstage = 2 * (size_t)pl.bn * 128;
    int nw = pl.bn <= 128 ? 3 : 2;
    if (const char* e = getenv("EIGB200_STREAM_NW")) { const int v = atoi(e); if (v >= 2 && v <= SK_MAX_STAGES) nw = v; }   // experiment hook
    int na = (int)((TC_SMEM_LIMIT - 2048 - nw * wstage) / TC_CHUNK_BYTES);
    if (na > 8) na = 8;
    if (na < 2) { set_error("tcgen05 GEMM: no room for the raw A ring (bn %d)", pl.bn); return EIGB200_EUNSUPPORTED; }
    p.nstages = nw; p.na_stages = na;
    smem_bytes = (size_t)na * TC_CHUNK_BYTES + nw * wstage + 1024 /*alignment*/ + 1024 /*barriers*/;
  }
  p.r_v8 = (lp.R && (((uintptr_t)lp.R & 31) == 0) && lp.ldr % 8 == 0) ? 1 : 0;
  const int64_t npairs = p.ntiles * p.nsplit;
  const unsigned grid = (unsigned)(npairs < (int64_t)num_sms() ? npairs : (int64_t)num_sms());
#define SK_LAUNCH(EPI_)                                                                                                                \
  do {                                                                                                                                 \
    if (f16 && rawa) {                                                                                                                 \
      EIGB_CUDA(cudaFuncSetAttribute(gemm_tc_stream_kernel<EPI_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)); \
      gemm_tc_stream_kernel<EPI_, true, true><<<grid, TC_THREADS, smem_bytes, st>>>(tAh, tAl, tWh, tWl, p);                           \
    } else if (f16) {                                                                                                                  \
      EIGB_CUDA(cudaFuncSetAttribute(gemm_tc_stream_kernel<EPI_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));  \
      gemm_tc_stream_kernel<EPI_, true><<<grid, TC_THREADS, pl.smem, st>>>(tAh, tAl, tWh, tWl, p);                                    \
    } else {                                                                                                                           \
      EIGB_CUDA(cudaFuncSetAttribute(gemm_tc_stream_kernel<EPI_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
      gemm_tc_stream_kernel<EPI_, false><<<grid, TC_THREADS, pl.smem, st>>>(tAh, tAl, tWh, tWl, p);                                   \
    }                                                                                                                                  \
  } while (0)
  switch (lp.epilogue) {
    case EIGB200_EPI_NONE: SK_LAUNCH(EIGB200_EPI_NONE); break;
    case EIGB200_EPI_GELU: SK_LAUNCH(EIGB200_EPI_GELU); break;
    case EIGB200_EPI_GLU_RESIDUAL: SK_LAUNCH(EIGB200_EPI_GLU_RESIDUAL); break;
    case EIGB200_EPI_RESIDUAL: SK_LAUNCH(EIGB200_EPI_RESIDUAL); break;
    default: set_error("linear: unknown epilogue %d", lp.epilogue); return EIGB200_EINVAL;
  }
#undef SK_LAUNCH
  EIGB_LAUNCH_CHECK("gemm_tc_stream_kernel");
  return EIGB200_OK;
}

int launch_linear_tc(cudaStream_t st, const LinearParams& lp, int nterms, void* workspace, int kind) {
  TcPlan pl = make_plan(lp.N, lp.K, lp.epilogue, kind);
  if (!pl.ok) return launch_linear_stream(st, lp, nterms, workspace, kind);   // K > 256 or a weight slice too large to stay resident
  if (!pl.ts) kind = 0;
  const bool glu = lp.epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const size_t wrows = (size_t)pl.nsplit * pl.bn;
  const bool prepared = lp.W == nullptr;                             // eigb200_linear_prepare filled the workspace for this (N, K, epilogue, LayerNorm, precision)
  if (!prepared) {
    int rc0 = tc_prepare(st, lp, workspace, kind);
    if (rc0 != EIGB200_OK) return rc0;
  }
  CUtensorMap tA, tWh, tWl;
  int rc;
  if ((rc = make_tmap(&tA, lp.A, (uint64_t)lp.M, (uint64_t)lp.K, (uint64_t)lp.lda, TC_BM))) return rc;
  const float* bias2;
  const float* scal = nullptr;
  if (kind == 1) {
    const __half* w_hi = reinterpret_cast<const __half*>(workspace);
    const __half* w_lo = w_hi + wrows * pl.kp64;
    bias2 = reinterpret_cast<const float*>(w_lo + wrows * pl.kp64);
    scal = bias2 + (lp.N + 3) / 4 * 4;
    if ((rc = make_tmap_f16(&tWh, w_hi, wrows, (uint64_t)pl.kp64, (uint32_t)pl.bn))) return rc;
    if ((rc = make_tmap_f16(&tWl, w_lo, wrows, (uint64_t)pl.kp64, (uint32_t)pl.bn))) return rc;
  } else {
    float* w_hi = reinterpret_cast<float*>(workspace);
    float* w_lo = w_hi + wrows * pl.kpad;
    bias2 = w_lo + wrows * pl.kpad;                                  // bias + W beta, behind the split weights in the workspace
    if ((rc = make_tmap(&tWh, w_hi, wrows, (uint64_t)pl.kpad, (uint64_t)pl.kpad, (uint32_t)pl.bn))) return rc;
    if ((rc = make_tmap(&tWl, w_lo, wrows, (uint64_t)pl.kpad, (uint64_t)pl.kpad, (uint32_t)pl.bn))) return rc;
  }
  const float* bias_eff = lp.ln_stats ? bias2 : lp.bias;

  TcParams p{};
  p.ln_stats = reinterpret_cast<const float2*>(lp.ln_stats);
  p.bias = bias_eff; p.C = lp.C; p.ldc = lp.ldc; p.R = lp.R; p.ldr = lp.ldr; p.M = lp.M; p.N = lp.N; p.K = lp.K; p.epilogue = lp.epilogue;
  p.bn = pl.bn; p.bg = pl.bg; p.nsplit = pl.nsplit; p.kchunks = pl.kchunks; p.nstages = pl.nstages; p.nterms = nterms;
  p.ntiles = (lp.M + TC_BM - 1) / TC_BM;
  p.zero = 0;
  p.kch_w = pl.kch_w; p.a_scale = lp.ln_stats ? 1024.f : 16.f; p.out_scale = scal; p.ovf_flag = nullptr;
  if (kind == 1) {
    p.ovf_flag = overflow_flag_ptr();
    if (!p.ovf_flag) { set_error("tcgen05 GEMM: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  }
  // TMEM-operand converter: DEFER publishes a chunk after the next chunk's shared-memory loads were issued.  Measured: +10 % on the converter-bound
  // N-split shape (bn = 96, in_proj), -6 % on the tensor-bound GLU shape (bn = 128) where the MMA wants its operand at once; a run-time flag
  // instead of the template parameter costs both shapes 3 % (the converter loop is latency-critical), hence two instantiations.
  const bool defer = pl.bn < 128;
  p.r_v8 = (lp.R && (((uintptr_t)lp.R & 31) == 0) && lp.ldr % 8 == 0) ? 1 : 0;
  {
    // EIGB200_GEMM_DIRECT_STORE: 0 = shuffle-transposed 128-bit stores (a warp instruction writes 4 full 128-byte lines), 1 = every thread stores its own
    // row segment with 256-bit stores, 2 (default) = direct stores in the GLU epilogue only.  Measured on B200: direct stores cost the plain / GELU
    // epilogues 10-50 % (32 scattered sectors per store instruction: DRAM write efficiency), the GLU + extractor epilogue gains 2 %.
    static int direct = -1;
    if (direct < 0) { const char* e = getenv("EIGB200_GEMM_DIRECT_STORE"); direct = e ? atoi(e) : 2; }
    const bool res_ok = !lp.R || p.r_v8 || !(glu || lp.epilogue == EIGB200_EPI_RESIDUAL);
    const bool want = direct == 1 || (direct == 2 && glu);
    p.c_v8 = (want && (((uintptr_t)lp.C & 31) == 0) && lp.ldc % 8 == 0 && res_ok) ? 1 : 0;
  }
  {
    // EIGB200_GEMM_EPI_STAGE: the non-GLU epilogues of the TMEM-operand kernel can transpose through 32 KB of shared memory taken from the raw ring
    // (two 16 KB stages) when at least 4 stages remain.  0 = shuffle transposes, 1 (default) = staged for the GELU epilogue only, 2 = staged for
    // all of them.  Measured: GELU (issue-bound epilogue) -2...-4 %, bias-only (87 % of the HBM peak, wants the ring depth) +1...+5 %.
    static int stage = -1;
    if (stage < 0) { const char* e = getenv("EIGB200_GEMM_EPI_STAGE"); stage = e ? atoi(e) : 1; }
    const bool res_ok = !lp.R || p.r_v8 || lp.epilogue != EIGB200_EPI_RESIDUAL;
    p.epi_stage = 0;
    // default: GELU whenever 4 stages remain; the HBM-paced bias-only / residual epilogues only when 6 remain (narrow N: -2...-3.5 % at N = 64)
    const bool want = stage == 2 || (stage == 1 && (lp.epilogue == EIGB200_EPI_GELU || pl.nstages - 2 >= 6));
    if (want && pl.ts && !glu && pl.ast == TS_ASTAGES && pl.nstages - 2 >= 4 && res_ok && !p.c_v8) { p.epi_stage = 2 * TC_CHUNK_BYTES; p.nstages = pl.nstages - 2; }
  }
  p.eig_w = nullptr; p.eig_part = nullptr;
  if (lp.eig_part) {                                                 // extractor partials ride in the GLU epilogue of the TMEM-operand kernel only
    if (!(glu && pl.ts && p.r_v8 && (lp.N / 2) % 16 == 0 && lp.eig_w)) {
      set_error("linear_glu_extract: needs the GLU epilogue with a 32-byte aligned residual and N/2 a multiple of 16 (N=%d)", lp.N);
      return EIGB200_EUNSUPPORTED;
    }
    p.eig_w = lp.eig_w; p.eig_part = lp.eig_part;
  }
  int workers = num_sms() / pl.nsplit;
  if (workers < 1) workers = 1;
  if ((int64_t)workers > p.ntiles) workers = (int)p.ntiles;
  p.workers = workers;
  dim3 grid(workers * pl.nsplit);
#define TC_LAUNCH_K(KERNEL_)                                                                                                    \
  do {                                                                                                                          \
    EIGB_CUDA(cudaFuncSetAttribute(KERNEL_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));                        \
    KERNEL_<<<grid, TC_THREADS, pl.smem, st>>>(tA, tWh, tWl, p);                                                                \
  } while (0)
#define TC_LAUNCH(EPI_)                                                                                                         \
  do {                                                                                                                          \
    if (kind == 1) {                                                                                                            \
      if (defer) TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, true, TS_ASTAGES, true>));                                                \
      else TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, false, TS_ASTAGES, true>));                                                     \
    } else if (pl.ts && pl.ast == 2 && EPI_ != EIGB200_EPI_GLU_RESIDUAL) {                                                      \
      if (gemm_wide_mode() == 2) TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, true, 2>));                                               \
      else TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, false, 2>));                                                                    \
    } else if (pl.ts && defer) {                                                                                                \
      TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, true>));                                                                             \
    } else if (pl.ts) {                                                                                                         \
      TC_LAUNCH_K((gemm_tc_ts_kernel<EPI_, false>));                                                                            \
    } else {                                                                                                                    \
      TC_LAUNCH_K((gemm_tc_kernel<EPI_>));                                                                                      \
    }                                                                                                                           \
  } while (0)
  switch (lp.epilogue) {
    case EIGB200_EPI_NONE: TC_LAUNCH(EIGB200_EPI_NONE); break;
    case EIGB200_EPI_GELU: TC_LAUNCH(EIGB200_EPI_GELU); break;
    case EIGB200_EPI_GLU_RESIDUAL: TC_LAUNCH(EIGB200_EPI_GLU_RESIDUAL); break;
    case EIGB200_EPI_RESIDUAL: TC_LAUNCH(EIGB200_EPI_RESIDUAL); break;
    default: set_error("linear: unknown epilogue %d", lp.epilogue); return EIGB200_EINVAL;
  }
#undef TC_LAUNCH
#undef TC_LAUNCH_K
  EIGB_LAUNCH_CHECK("gemm_tc_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200
