// k1_eigen.cu -- K1 family: fused gate-projection -> discretisation -> eigenvalue -> on-device bin counts.
//
// Reference operators replaced (see include/eigb200.h for the per-entry citations):
//   get_eig_mamba2 / get_eig_mamba2_LTI / get_eig_att_norm (first half)   analysis/eval_eig.py:137-205
//   threshold_analysis                                                      analysis/eval_eig.py:335-362
//
// Roofline: HBM.  Algorithmic bytes per eigenvalue = D*s/H (read x once) + 4 (write lambda); ~0.5 flop/B.
// Data layout: x (B,T,D) row-major, rows 16-byte aligned.  One warp owns a contiguous run of rows of one sequence,
// walks it R=4 rows at a time (4 independent 128-bit streaming loads per lane in flight per 128 columns), keeps the
// H gate rows of W in shared memory (one LDS.128 feeds 4 rows), reduces the R*HT partial dot products with a
// butterfly transpose-reduce so that lane (r,h) ends up with exactly one full dot product, applies the
// transcendental epilogue once per eigenvalue, stores lambda coalesced and bins it in registers.  Bin counts are
// flushed once per CTA with integer atomics (order independent => bit reproducible).
#include "common.cuh"

namespace eigb200 {

constexpr int K1_R = 4;          // rows per warp iteration
constexpr int K1_WARPS = 4;      // warps per CTA

enum { K1_EPI_MAMBA2 = 0, K1_EPI_NORMGATE = 1 };

struct K1Params {
  const void* x; const float* W; const float* p0; const float* p1; const float* p2;   // W (H,D); per-head vectors
  float* out; int64_t out_stride; int* counts;
  float2* rowstats; float ln_eps;                   // optional (mean, rstd) of every row of x: LayerNorm statistics for the next block
  int T, D, H, rows_per_warp, norm_fn;
  EdgesF e;
};

template <bool BF16> struct XLoad;
template <> struct XLoad<false> {
  static constexpr int VEC = 4;                      // floats per 128-bit load
  __device__ static __forceinline__ void load(const void* row, int c, float (&v)[8]) {
    float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(row) + c);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct XLoad<true> {
  static constexpr int VEC = 8;                      // bf16 per 128-bit load
  __device__ static __forceinline__ void load(const void* row, int c, float (&v)[8]) {
    uint4 t = ldg_stream_u4(reinterpret_cast<const uint4*>(row) + c);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};

__device__ __forceinline__ float norm_fn_apply(int fn, float z) {
  switch (fn) {
    case EIGB200_NORM_EXP: return expf(z);
    case EIGB200_NORM_ELU: return elu_f(z);
    case EIGB200_NORM_SOFTPLUS: return softplus_f(z);
    default: return sigmoid_f(z);
  }
}

template <int HT, bool BF16, int EPI, bool STATS>
__global__ void __launch_bounds__(K1_WARPS * 32) k1_gate_kernel(const K1Params p) {
  constexpr int VEC = XLoad<BF16>::VEC;
  constexpr int NV = K1_R * HT;                       // partial sums per lane
  constexpr int SH = 5 - Log2<NV>::value;             // lanes sharing one result = 1 << SH
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                   // [HT][D]
  __shared__ int hist[HT][EIGB200_NSLOT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int hbase = blockIdx.z * HT;
  const int D = p.D, T = p.T, H = p.H;

  for (int i = threadIdx.x; i < HT * D; i += blockDim.x) {
    const int h = i / D, c = i - h * D;
    Ws[i] = (hbase + h < H) ? p.W[(size_t)(hbase + h) * D + c] : 0.f;
  }
  if (threadIdx.x < HT * EIGB200_NSLOT) (&hist[0][0])[threadIdx.x] = 0;
  __syncthreads();

  // this lane's (row-in-group, head) after the transpose-reduce
  const int idx = lane >> SH;
  const int r_own = idx / HT, h_own = idx - r_own * HT;
  const int hg = hbase + h_own;
  const bool lead = (lane & ((1 << SH) - 1)) == 0 && hg < H;
  float e0 = 0.f, e1 = 0.f, e2 = 0.f;
  if (hg < H) {
    if (EPI == K1_EPI_MAMBA2) { e0 = p.p0[hg]; e1 = -expf(p.p1[hg]); }               // dt_bias, A = -exp(A_log)
    else { e0 = p.p0[hg] + (p.p2 ? p.p2[hg] : 0.f); }                                  // bias (+ offset)
  }
  // NOTE: the reference adds bias inside the Linear and the offset afterwards, both in fp32: (dot + b) + offset.
  const float bias_only = (EPI == K1_EPI_NORMGATE && hg < H) ? p.p0[hg] : 0.f;
  const float offs_only = (EPI == K1_EPI_NORMGATE && hg < H && p.p2) ? p.p2[hg] : 0.f;
  (void)e0;

  int cnt[EIGB_NCNT];
#pragma unroll
  for (int j = 0; j < EIGB_NCNT; ++j) cnt[j] = 0;

  const int wid = blockIdx.x * K1_WARPS + warp;
  const int t0 = wid * p.rows_per_warp;
  const int t1 = min(T, t0 + p.rows_per_warp);
  const size_t row_bytes = (size_t)D * (BF16 ? 2 : 4);
  const char* xb = reinterpret_cast<const char*>(p.x) + (size_t)b * T * row_bytes;
  const int nvec = D / VEC;
  const float4* Ws4 = reinterpret_cast<const float4*>(Ws);

  const bool want_stats = STATS && blockIdx.z == 0;   // compile-time off for the plain extractor: no extra registers in its hot loop
  for (int t = t0; t < t1; t += K1_R) {
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    float s1[K1_R], s2[K1_R], shift[K1_R];            // shifted moments: sum(x - x_0), sum((x - x_0)^2)
#pragma unroll
    for (int r = 0; r < K1_R; ++r) { s1[r] = 0.f; s2[r] = 0.f; shift[r] = 0.f; }
    const char* rows[K1_R];
#pragma unroll
    for (int r = 0; r < K1_R; ++r) rows[r] = xb + (size_t)min(t + r, t1 - 1) * row_bytes;   // clamp: tail rows re-read a valid row
    if (want_stats) {                                   // shift = first element of the row (same broadcast load in every lane)
#pragma unroll
      for (int r = 0; r < K1_R; ++r)
        shift[r] = BF16 ? __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(rows[r])) << 16)
                        : __ldg(reinterpret_cast<const float*>(rows[r]));
    }
#pragma unroll 2
    for (int c = lane; c < nvec; c += 32) {
      float xv[K1_R][8];
#pragma unroll
      for (int r = 0; r < K1_R; ++r) XLoad<BF16>::load(rows[r], c, xv[r]);
      if (want_stats) {
#pragma unroll
        for (int r = 0; r < K1_R; ++r)
#pragma unroll
          for (int e = 0; e < VEC; ++e) { const float dlt = xv[r][e] - shift[r]; s1[r] += dlt; s2[r] = fmaf(dlt, dlt, s2[r]); }
      }
#pragma unroll
      for (int h = 0; h < HT; ++h) {
#pragma unroll
        for (int q = 0; q < VEC / 4; ++q) {
          const float4 w = Ws4[(h * D + c * VEC) / 4 + q];
#pragma unroll
          for (int r = 0; r < K1_R; ++r) {
            float a = acc[r * HT + h];
            a = fmaf(xv[r][4 * q + 0], w.x, a);
            a = fmaf(xv[r][4 * q + 1], w.y, a);
            a = fmaf(xv[r][4 * q + 2], w.z, a);
            a = fmaf(xv[r][4 * q + 3], w.w, a);
            acc[r * HT + h] = a;
          }
        }
      }
    }
    const float dot = transpose_reduce<NV>(acc, lane);
    if (want_stats) {
      const float m1 = transpose_reduce<K1_R>(s1, lane), m2 = transpose_reduce<K1_R>(s2, lane);
      const int rr = lane >> 3;                          // K1_R = 4 partials: lanes 8r..8r+7 hold row r
      const float sh = rr == 0 ? shift[0] : (rr == 1 ? shift[1] : (rr == 2 ? shift[2] : shift[3]));
      if ((lane & 7) == 0 && t + rr < t1) {
        const float invD = 1.f / (float)D;
        const float md = m1 * invD;                      // mean - shift
        const float var = fmaxf(m2 * invD - md * md, 0.f);
        p.rowstats[(size_t)b * T + t + rr] = make_float2(sh + md, rsqrtf(var + p.ln_eps));
      }
    }
    const int row = t + r_own;
    if (lead && row < t1) {
      float val;
      if (EPI == K1_EPI_MAMBA2) {
        const float dt = softplus_f(dot + e0);
        val = expf(dt * e1);
      } else {
        const float raw = (dot + bias_only) + offs_only;
        val = expf(-norm_fn_apply(p.norm_fn, raw));
      }
      if (p.out) p.out[(((size_t)b * T + row) * H + hg) * p.out_stride] = val;
      if (EPI == K1_EPI_MAMBA2) {
        // the reference bins the float32 radius sqrt(fl(re^2) + fl(im^2)) of the real lambda (eval_eig.py:605-606)
        bin_f32(sqrtf(__fmul_rn(val, val)), p.e, cnt);
        cnt[8] += (val == val) ? 1 : 0;                 // arctan2(0, lambda) = 0 unless lambda is NaN
      }
    }
  }

  if (EPI == K1_EPI_MAMBA2 && p.counts) {
    if (lead) {
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j)
        if (cnt[j]) atomicAdd(&hist[h_own][slot_of(j, p.e.nb)], cnt[j]);
    }
    __syncthreads();
    if (threadIdx.x < HT * EIGB200_NSLOT) {
      const int h = threadIdx.x / EIGB200_NSLOT, s = threadIdx.x % EIGB200_NSLOT;
      const int v = hist[h][s];
      if (v && hbase + h < H) atomicAdd(&p.counts[((size_t)b * H + hbase + h) * EIGB200_NSLOT + s], v);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Small-H, small-D specialisation of the Mamba-2 extractor (BASELINE C2: D = 128, H = 1): 8 lanes share a row.
// Lane l of an 8-lane group loads the float4s l, l+8, l+16, ... of its row (a warp instruction reads 4 rows x 128 contiguous bytes), keeps
// its slice of the H gate rows of W in REGISTERS (no shared-memory traffic in the loop) and the three per-row sums -- the gate dot product and,
// for the LayerNorm statistics of the next block, sum(x - x0) and sum((x - x0)^2) -- are reduced with 3 xor-shuffles each inside the group
// (9 shuffles per 4 rows instead of 18 with the warp-per-4-rows butterfly).  Two row groups (8 rows) are in flight per trip.
// ---------------------------------------------------------------------------------------------------------------------------
template <int HT, int NF4, bool STATS>
__global__ void __launch_bounds__(K1_WARPS * 32) k1_row8_kernel(const K1Params p) {
  __shared__ int hist[HT][EIGB200_NSLOT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int D = p.D, T = p.T, H = p.H;
  const int g = lane >> 3, l = lane & 7;
  if (threadIdx.x < HT * EIGB200_NSLOT) (&hist[0][0])[threadIdx.x] = 0;
  __syncthreads();

  float4 w[HT][NF4];
#pragma unroll
  for (int h = 0; h < HT; ++h)
#pragma unroll
    for (int j = 0; j < NF4; ++j)
      w[h][j] = (h < H) ? __ldg(reinterpret_cast<const float4*>(p.W + (size_t)h * D) + l + 8 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
  // after the butterfly every lane of the group holds the full sums: lane l < H finishes head l
  const bool fin = l < H;
  const float e0 = fin ? p.p0[l] : 0.f;                               // dt_bias
  const float e1 = fin ? -expf(p.p1[l]) : 0.f;                        // A = -exp(A_log)
  int cnt[EIGB_NCNT];
#pragma unroll
  for (int j = 0; j < EIGB_NCNT; ++j) cnt[j] = 0;

  const int wid = blockIdx.x * K1_WARPS + warp;
  const int t0 = wid * p.rows_per_warp;
  const int t1 = min(T, t0 + p.rows_per_warp);
  const float4* xb = reinterpret_cast<const float4*>(p.x) + (size_t)b * T * (D / 4);
  constexpr int U = 2;                                                // row groups per trip
  for (int t = t0; t < t1; t += 4 * U) {
    float4 xv[U][NF4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = min(t + 4 * u + g, t1 - 1);                     // clamp: tail lanes re-read a valid row
      const float4* xr = xb + (size_t)row * (D / 4) + l;
#pragma unroll
      for (int j = 0; j < NF4; ++j) xv[u][j] = ldg_stream_f4(xr + 8 * j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = t + 4 * u + g;
      float acc[HT], s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll
      for (int h = 0; h < HT; ++h) acc[h] = 0.f;
      if (STATS) shift = __shfl_sync(0xffffffffu, xv[u][0].x, lane & ~7);   // first element of the row
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        const float4 v = xv[u][j];
#pragma unroll
        for (int h = 0; h < HT; ++h) {
          float a = acc[h];
          a = fmaf(v.x, w[h][j].x, a); a = fmaf(v.y, w[h][j].y, a); a = fmaf(v.z, w[h][j].z, a); a = fmaf(v.w, w[h][j].w, a);
          acc[h] = a;
        }
        if (STATS) {
          const float d0 = v.x - shift, d1 = v.y - shift, d2 = v.z - shift, d3 = v.w - shift;
          s1 += (d0 + d1) + (d2 + d3);
          s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
        }
      }
#pragma unroll
      for (int ofs = 4; ofs >= 1; ofs >>= 1) {
#pragma unroll
        for (int h = 0; h < HT; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], ofs);
        if (STATS) { s1 += __shfl_xor_sync(0xffffffffu, s1, ofs); s2 += __shfl_xor_sync(0xffffffffu, s2, ofs); }
      }
      if (row < t1) {
        if (STATS && l == 7) {                                        // a lane that finishes no head
          const float invD = 1.f / (float)D;
          const float md = s1 * invD;                                 // mean - shift
          const float var = fmaxf(s2 * invD - md * md, 0.f);
          p.rowstats[(size_t)b * T + row] = make_float2(shift + md, rsqrtf(var + p.ln_eps));
        }
        if (fin) {
          float dot = acc[0];
#pragma unroll
          for (int h = 1; h < HT; ++h) dot = (l == h) ? acc[h] : dot;
          const float dt = softplus_f(dot + e0);
          const float val = expf(dt * e1);
          if (p.out) p.out[(((size_t)b * T + row) * H + l) * p.out_stride] = val;
          bin_f32(sqrtf(__fmul_rn(val, val)), p.e, cnt);             // eval_eig.py:605-606: float32 radius sqrt(fl(re^2))
          cnt[8] += (val == val) ? 1 : 0;
        }
      }
    }
  }
  if (p.counts) {
    if (fin) {
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j)
        if (cnt[j]) atomicAdd(&hist[l][slot_of(j, p.e.nb)], cnt[j]);
    }
    __syncthreads();
    if (threadIdx.x < HT * EIGB200_NSLOT) {
      const int h = threadIdx.x / EIGB200_NSLOT, sl = threadIdx.x % EIGB200_NSLOT;
      const int v = hist[h][sl];
      if (v && h < H) atomicAdd(&p.counts[((size_t)b * H + h) * EIGB200_NSLOT + sl], v);
    }
  }
}

template <int HT, int NF4>
static void launch_row8(cudaStream_t st, const K1Params& q, dim3 grid, dim3 block) {
  if (q.rowstats) k1_row8_kernel<HT, NF4, true><<<grid, block, 0, st>>>(q);
  else k1_row8_kernel<HT, NF4, false><<<grid, block, 0, st>>>(q);
}


// ---------------------------------------------------------------------------------------------------------------------------
// Mamba-2 extractor from the partials the GLU epilogue left behind (eigb200_linear_glu_extract): per row, NG = D / 16 groups of
// (gate dot product, mean, M2) over 16 columns each, stored part[(g * 3 + c) * rows + row].  They are combined in a FIXED order (so the result
// does not depend on scheduling): dot = ((p0 + p1) + (p2 + p3)) + ..., moments by Chan's pairwise merge.  One thread per row; a warp's 32 rows
// belong to one sequence when T % 32 == 0 (bin counters then flush once per warp), otherwise every thread flushes its own.
// ---------------------------------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(256) k1_partials_kernel(const float* __restrict__ part, int64_t rows, int T, const float* __restrict__ dt_bias_p,
                                                           const float* __restrict__ A_log_p, float* __restrict__ lam,
                                                           int64_t lam_stride, int* __restrict__ counts, EdgesF e, float2* __restrict__ rowstats, float ln_eps) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = row < rows;
  const int64_t r = ok ? row : rows - 1;
  const float dt_bias = __ldg(dt_bias_p), A = -expf(__ldg(A_log_p));   // one head
  float dot[NG], mean[NG], m2[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    dot[g] = ldg_stream_f1(part + (size_t)(g * 3 + 0) * rows + r);
    mean[g] = ldg_stream_f1(part + (size_t)(g * 3 + 1) * rows + r);
    m2[g] = ldg_stream_f1(part + (size_t)(g * 3 + 2) * rows + r);
  }
  float n = 16.f;
#pragma unroll
  for (int w = 1; w < NG; w <<= 1) {                               // pairwise tree: (0,1) (2,3) ... then (01,23) ...
#pragma unroll
    for (int g = 0; g + w < NG; g += 2 * w) {
      dot[g] += dot[g + w];
      const float delta = mean[g + w] - mean[g];
      mean[g] = 0.5f * (mean[g] + mean[g + w]);                    // equal counts
      m2[g] = (m2[g] + m2[g + w]) + delta * delta * (0.5f * n);
    }
    n *= 2.f;
  }
  const float dt = softplus_f(dot[0] + dt_bias);
  const float val = expf(dt * A);
  int cnt[EIGB_NCNT];
#pragma unroll
  for (int j = 0; j < EIGB_NCNT; ++j) cnt[j] = 0;
  if (ok) {
    if (lam) lam[(size_t)row * lam_stride] = val;
    if (rowstats) rowstats[row] = make_float2(mean[0], rsqrtf(fmaxf(m2[0] / n, 0.f) + ln_eps));
    bin_f32(sqrtf(__fmul_rn(val, val)), e, cnt);
    cnt[8] = (val == val) ? 1 : 0;
  }
  if (counts) {
    const int64_t b = r / T;
    if (T % 32 == 0) {                                             // the whole warp shares the sequence: one atomic per slot per warp
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j) {
        const int v = __reduce_add_sync(0xffffffffu, cnt[j]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[(size_t)b * EIGB200_NSLOT + slot_of(j, e.nb)], v);
      }
    } else if (ok) {
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j)
        if (cnt[j]) atomicAdd(&counts[(size_t)b * EIGB200_NSLOT + slot_of(j, e.nb)], cnt[j]);
    }
  }
}

template <bool BF16, int EPI>
static int launch_k1(cudaStream_t st, const K1Params& p, int64_t B) {
  int ht = 1;
  while (ht < p.H && ht < 8) ht <<= 1;
  const int hgroups = (p.H + ht - 1) / ht;
  // rows per warp: aim for >= 8 CTAs per SM overall, but never fewer than 2 iterations per warp
  const int64_t target_ctas = (int64_t)num_sms() * 16;
  int ctas_per_seq = (int)((target_ctas + B * hgroups - 1) / (B * hgroups));
  const int max_cps = (p.T + K1_WARPS * 2 * K1_R - 1) / (K1_WARPS * 2 * K1_R);
  if (ctas_per_seq > max_cps) ctas_per_seq = max_cps;
  if (ctas_per_seq < 1) ctas_per_seq = 1;
  K1Params q = p;
  int rpw = (p.T + ctas_per_seq * K1_WARPS - 1) / (ctas_per_seq * K1_WARPS);
  rpw = (rpw + K1_R - 1) / K1_R * K1_R;
  q.rows_per_warp = rpw;
  ctas_per_seq = (p.T + rpw * K1_WARPS - 1) / (rpw * K1_WARPS);
  dim3 grid(ctas_per_seq, (unsigned)B, hgroups), block(K1_WARPS * 32);
  EIGB_CHECK_ARG(B <= 65535, "k1: batch %lld exceeds grid.y limit 65535; split the call", (long long)B);
  if (EPI == K1_EPI_MAMBA2 && !BF16 && p.H <= 2 && p.D % 32 == 0 && p.D <= 256 && (p.D & (p.D - 1)) == 0) {
    // 8-lanes-per-row specialisation (fp32, H <= 2, D in {32,64,128,256}); rows per warp rounded to whole trips of 8 rows
    rpw = (rpw + 7) / 8 * 8;
    q.rows_per_warp = rpw;
    grid.x = (p.T + rpw * K1_WARPS - 1) / (rpw * K1_WARPS);
    const int nf4 = p.D / 32;
#define K1_ROW8(HT_)                                                                                               \
    switch (nf4) { case 1: launch_row8<HT_, 1>(st, q, grid, block); break; case 2: launch_row8<HT_, 2>(st, q, grid, block); break;      \
                   case 4: launch_row8<HT_, 4>(st, q, grid, block); break; default: launch_row8<HT_, 8>(st, q, grid, block); break; }
    if (p.H == 1) { K1_ROW8(1) } else { K1_ROW8(2) }
#undef K1_ROW8
    EIGB_LAUNCH_CHECK("k1_row8_kernel");
    return EIGB200_OK;
  }
  const size_t smem = (size_t)ht * p.D * sizeof(float);
  const bool stats = (EPI == K1_EPI_MAMBA2) && p.rowstats != nullptr;
#define K1_LAUNCH(HT_, ST_)                                                                                       \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      EIGB_CUDA(cudaFuncSetAttribute(k1_gate_kernel<HT_, BF16, EPI, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k1_gate_kernel<HT_, BF16, EPI, ST_><<<grid, block, smem, st>>>(q);                                            \
  } while (0)
#define K1_CASE(HT_)                                                                                              \
  case HT_: {                                                                                                     \
    if (stats) K1_LAUNCH(HT_, (EPI == K1_EPI_MAMBA2)); else K1_LAUNCH(HT_, false);                                \
  } break;
  switch (ht) { K1_CASE(1) K1_CASE(2) K1_CASE(4) K1_CASE(8) default: set_error("k1: bad head tile"); return EIGB200_EINVAL; }
#undef K1_CASE
#undef K1_LAUNCH
  EIGB_LAUNCH_CHECK("k1_gate_kernel");
  return EIGB200_OK;
}

// ---- LTI: parameter-only eigenvalues broadcast over (B,T) ------------------------------------------------------
__global__ void k1_lti_kernel(const float* A, const float* beta, int64_t BT, int T, int H, float* lam, int64_t lam_stride, int* counts, EdgesF e) {
  // one thread per head computes lambda and its bins once; the broadcast store is a plain grid-stride fill
  __shared__ float lam_s[256];
  for (int h = threadIdx.x; h < H; h += blockDim.x) lam_s[h] = expf(beta[h] * -softplus_f(A[h]));
  __syncthreads();
  if (lam) {
    const int64_t n = BT * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) lam[i * lam_stride] = lam_s[i % H];
  }
  if (counts) {
    const int64_t Bn = BT / T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Bn * H; i += (int64_t)gridDim.x * blockDim.x) {
      const int h = (int)(i % H);
      int c[EIGB_NCNT];
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j) c[j] = 0;
      bin_f32(sqrtf(__fmul_rn(lam_s[h], lam_s[h])), e, c);
      c[8] = (lam_s[h] == lam_s[h]) ? 1 : 0;
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j)
        if (c[j]) atomicAdd(&counts[i * EIGB200_NSLOT + slot_of(j, e.nb)], c[j] * T);
    }
  }
}

// ---- ratios + bins over a (B,N,inner) array -------------------------------------------------------------------------
struct RatioParams {
  const void* a; double* out; int64_t out_stride; int* counts;
  int64_t N, inner; int mode; int chunk;             // chunk = values of n handled per CTA
  const float* mscale;                               // softmax eta: v *= exp(double(mscale[n] - mscale[n+1])) (eval_eig.py:83-88), no zero guard
  EdgesF ef; EdgesD ed; int cmp_f32_on_f32;
};

template <typename TIn>
__global__ void __launch_bounds__(256) ratio_hist_kernel(const RatioParams p) {
  // thread -> fixed inner index i (stride S is a multiple of `inner` when inner <= 256) so that bin counters live in registers
  extern __shared__ int hist_s[];                     // [min(inner,256)][8]
  const int64_t inner = p.inner;
  const int b = blockIdx.y;
  const int64_t Nout = p.mode == EIGB200_RATIO_NONE ? p.N : p.N - 1;
  const int64_t n0 = (int64_t)blockIdx.x * p.chunk, n1 = min(Nout, n0 + p.chunk);
  const TIn* a = reinterpret_cast<const TIn*>(p.a) + (size_t)b * p.N * inner;
  const bool small = inner <= 256;
  const int per = small ? (int)(256 / inner) : 1;     // n-values handled per CTA step
  const int S = small ? per * (int)inner : 256;
  if (small) { for (int i = threadIdx.x; i < inner * EIGB200_NSLOT; i += blockDim.x) hist_s[i] = 0; __syncthreads(); }
  int cnt[EIGB_NCNT];
#pragma unroll
  for (int j = 0; j < EIGB_NCNT; ++j) cnt[j] = 0;
  const int nb = p.ed.nb;
  if ((int)threadIdx.x < S) {
    const int64_t total = (n1 - n0) * inner;
    for (int64_t k = threadIdx.x; k < total; k += S) {
      const int64_t e = n0 * inner + k;
      double v; float vf = 0.f;
      if (p.mode == EIGB200_RATIO_NONE) {
        v = (double)a[e]; vf = (float)a[e];
      } else {
        double u0 = (double)a[e], u1 = (double)a[e + inner];
        if (!p.mscale) {
          if (u0 == 0.0) u0 = 2e-23;
          if (u1 == 0.0) u1 = 2e-23;
        }
        v = p.mode == EIGB200_RATIO_NEXT_OVER_CUR ? u1 / u0 : u0 / u1;
        if (p.mscale) {                                  // different row maxima: rescale by exp(m_t - m_{t+1}), difference formed in float32
          const float* ms = p.mscale + (size_t)b * p.N * inner;
          v *= exp((double)(ms[e] - ms[e + inner]));
        }
        if (p.out) p.out[((size_t)b * Nout * inner + e) * p.out_stride] = v;
      }
      int c[EIGB_NCNT];
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j) c[j] = 0;
      if (sizeof(TIn) == 4 && p.mode == EIGB200_RATIO_NONE) bin_f32(vf, p.ef, c); else bin_f64(v, p.ed, c);
      c[8] = (v - v == 0.0) ? 1 : 0;                   // finite: 0*v lands in the first phase bin (eval_eig.py:673-674)
      if (small) {
#pragma unroll
        for (int j = 0; j < EIGB_NCNT; ++j) cnt[j] += c[j];
      } else {
        const int64_t i = e % inner;
#pragma unroll
        for (int j = 0; j < EIGB_NCNT; ++j)
          if (c[j]) atomicAdd(&p.counts[((size_t)b * inner + i) * EIGB200_NSLOT + slot_of(j, nb)], c[j]);
      }
    }
    if (small && p.counts) {
      const int i = threadIdx.x % (int)inner;
#pragma unroll
      for (int j = 0; j < EIGB_NCNT; ++j)
        if (cnt[j]) atomicAdd(&hist_s[i * EIGB200_NSLOT + slot_of(j, nb)], cnt[j]);
    }
  }
  if (small && p.counts) {
    __syncthreads();
    for (int i = threadIdx.x; i < inner * EIGB200_NSLOT; i += blockDim.x)
      if (hist_s[i]) atomicAdd(&p.counts[(size_t)b * inner * EIGB200_NSLOT + i], hist_s[i]);
  }
}

// sum_b c and sum_b c^2 of the per-sample bin counts (eval_eig.py:620-623 forms mean / std over the batch axis from them): counts (L, B, inner8)
// -> sum, sumsq (L, inner8) int64.  A CTA takes CM_ROWS samples of one layer x up to 256 columns: thread = (row lane, column), partial sums meet in
// shared memory (integer atomics: order-independent, bit-reproducible) and leave with ONE global 64-bit atomic per column and CTA.
constexpr int CM_ROWS = 512;
__global__ void __launch_bounds__(256) count_moments_kernel(const int* __restrict__ counts, int64_t B, int64_t inner8,
                                                            unsigned long long* __restrict__ sum, unsigned long long* __restrict__ sumsq) {
  __shared__ unsigned long long s_sum[256], s_sq[256];
  const int ncol = (int)min((int64_t)256, inner8 - (int64_t)blockIdx.y * 256);     // columns of this CTA
  const int col = threadIdx.x % ncol, rlane = threadIdx.x / ncol, nrl = 256 / ncol;
  const int64_t c0 = (int64_t)blockIdx.y * 256, layer = blockIdx.z;
  s_sum[threadIdx.x] = 0ull; s_sq[threadIdx.x] = 0ull;
  __syncthreads();
  const int64_t b0 = (int64_t)blockIdx.x * CM_ROWS, b1 = min(B, b0 + CM_ROWS);
  const int* base = counts + (layer * B) * inner8 + c0 + col;
  unsigned long long s = 0ull, s2 = 0ull;
  if (rlane < nrl) {
    for (int64_t b = b0 + rlane; b < b1; b += nrl) { const long long c = __ldg(base + b * inner8); s += (unsigned long long)c; s2 += (unsigned long long)(c * c); }
    atomicAdd(&s_sum[col], s); atomicAdd(&s_sq[col], s2);
  }
  __syncthreads();
  if ((int)threadIdx.x < ncol) {
    atomicAdd(sum + layer * inner8 + c0 + threadIdx.x, s_sum[threadIdx.x]);
    atomicAdd(sumsq + layer * inner8 + c0 + threadIdx.x, s_sq[threadIdx.x]);
  }
}

__global__ void zero_i32_kernel(int* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_zero_i32(void* stream, int32_t* d_buf, size_t n) {
  EIGB_CHECK_ARG(d_buf || n == 0, "zero_i32: null buffer");
  if (n == 0) return EIGB200_OK;
  EIGB_CUDA(cudaMemsetAsync(d_buf, 0, n * sizeof(int32_t), (cudaStream_t)stream));
  return EIGB200_OK;
}

extern "C" int eigb200_mamba2_eig(void* stream, const void* d_x, int x_dtype, int64_t B, int64_t T, int D,
                                  const float* d_W_dt, const float* d_dt_bias, const float* d_A_log, int H,
                                  float* d_lam, int64_t lam_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode,
                                  float* d_rowstats, float ln_eps) {
  EIGB_CHECK_ARG(d_x && d_W_dt && d_dt_bias && d_A_log, "mamba2_eig: null input pointer");
  EIGB_CHECK_ARG(!d_lam || lam_stride >= 1, "mamba2_eig: lam_stride must be >= 1");
  EIGB_CHECK_ARG(B > 0 && T > 0 && D > 0 && H > 0, "mamba2_eig: bad shape B=%lld T=%lld D=%d H=%d", (long long)B, (long long)T, D, H);
  EIGB_CHECK_ARG(x_dtype == EIGB200_F32 || x_dtype == EIGB200_BF16, "mamba2_eig: x_dtype must be F32 or BF16");
  EIGB_CHECK_ARG(D % (x_dtype == EIGB200_BF16 ? 8 : 4) == 0, "mamba2_eig: D=%d must be a multiple of %d (128-bit rows)", D, x_dtype == EIGB200_BF16 ? 8 : 4);
  EIGB_CHECK_ARG(((uintptr_t)d_x & 15) == 0, "mamba2_eig: x must be 16-byte aligned");
  EIGB_CHECK_ARG(T < (1LL << 31), "mamba2_eig: T too large");
  K1Params p{};
  p.x = d_x; p.W = d_W_dt; p.p0 = d_dt_bias; p.p1 = d_A_log; p.p2 = nullptr; p.out = d_lam; p.out_stride = lam_stride; p.counts = d_counts;
  p.rowstats = reinterpret_cast<float2*>(d_rowstats); p.ln_eps = ln_eps;
  p.T = (int)T; p.D = D; p.H = H; p.norm_fn = 0;
  if (d_counts) { int rc = make_edges_f(thresholds, nthr, compare_mode, &p.e); if (rc) return rc; }
  else { double one = 1.0; make_edges_f(&one, 1, 0, &p.e); }
  return x_dtype == EIGB200_BF16 ? launch_k1<true, K1_EPI_MAMBA2>((cudaStream_t)stream, p, B)
                                 : launch_k1<false, K1_EPI_MAMBA2>((cudaStream_t)stream, p, B);
}

extern "C" int eigb200_normattn_gate(void* stream, const void* d_x, int x_dtype, int64_t B, int64_t T, int D,
                                     const float* d_W_n, const float* d_b_n, const float* d_offset, int H, int norm_fn,
                                     float* d_n) {
  EIGB_CHECK_ARG(d_x && d_W_n && d_b_n && d_n, "normattn_gate: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0 && D > 0 && H > 0, "normattn_gate: bad shape");
  EIGB_CHECK_ARG(norm_fn >= EIGB200_NORM_EXP && norm_fn <= EIGB200_NORM_SIGMOID, "normalization function %d not implemented!", norm_fn);
  EIGB_CHECK_ARG(x_dtype == EIGB200_F32 || x_dtype == EIGB200_BF16, "normattn_gate: x_dtype must be F32 or BF16");
  EIGB_CHECK_ARG(D % (x_dtype == EIGB200_BF16 ? 8 : 4) == 0, "normattn_gate: D=%d must be a multiple of %d", D, x_dtype == EIGB200_BF16 ? 8 : 4);
  EIGB_CHECK_ARG(((uintptr_t)d_x & 15) == 0, "normattn_gate: x must be 16-byte aligned");
  K1Params p{};
  p.x = d_x; p.W = d_W_n; p.p0 = d_b_n; p.p1 = nullptr; p.p2 = d_offset; p.out = d_n; p.out_stride = 1; p.counts = nullptr;
  p.T = (int)T; p.D = D; p.H = H; p.norm_fn = norm_fn;
  double one = 1.0; make_edges_f(&one, 1, 0, &p.e);
  return x_dtype == EIGB200_BF16 ? launch_k1<true, K1_EPI_NORMGATE>((cudaStream_t)stream, p, B)
                                 : launch_k1<false, K1_EPI_NORMGATE>((cudaStream_t)stream, p, B);
}

extern "C" int eigb200_mamba2_lti_eig(void* stream, const float* d_A, const float* d_beta, int64_t B, int64_t T, int H,
                                      float* d_lam, int64_t lam_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode) {
  EIGB_CHECK_ARG(d_A && d_beta, "mamba2_lti_eig: null pointer");
  EIGB_CHECK_ARG(!d_lam || lam_stride >= 1, "mamba2_lti_eig: lam_stride must be >= 1");
  EIGB_CHECK_ARG(B > 0 && T > 0 && H > 0 && H <= 256, "mamba2_lti_eig: bad shape (H <= 256)");
  EdgesF e; double one = 1.0;
  if (d_counts) { int rc = make_edges_f(thresholds, nthr, compare_mode, &e); if (rc) return rc; } else make_edges_f(&one, 1, 0, &e);
  const int64_t n = B * T * H;
  int grid = (int)((n + 255) / 256); if (grid > num_sms() * 8) grid = num_sms() * 8; if (grid < 1) grid = 1;
  k1_lti_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_A, d_beta, B * T, (int)T, H, d_lam, lam_stride, d_counts, e);
  EIGB_LAUNCH_CHECK("k1_lti_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_ratio_hist(void* stream, const void* d_a, int dtype, int mode, int64_t B, int64_t N, int64_t inner,
                                  double* d_out, int64_t out_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode) {
  EIGB_CHECK_ARG(d_a, "ratio_hist: null input");
  EIGB_CHECK_ARG(!d_out || out_stride >= 1, "ratio_hist: out_stride must be >= 1");
  EIGB_CHECK_ARG(dtype == EIGB200_F32 || dtype == EIGB200_F64, "ratio_hist: dtype must be F32 or F64");
  EIGB_CHECK_ARG(mode >= EIGB200_RATIO_NONE && mode <= EIGB200_RATIO_CUR_OVER_NEXT, "ratio_hist: bad mode %d", mode);
  EIGB_CHECK_ARG(B > 0 && inner > 0 && N > (mode == EIGB200_RATIO_NONE ? 0 : 1), "ratio_hist: bad shape");
  EIGB_CHECK_ARG(B <= 65535, "ratio_hist: batch exceeds 65535; split the call");
  RatioParams p{};
  p.a = d_a; p.out = d_out; p.out_stride = out_stride; p.counts = d_counts; p.N = N; p.inner = inner; p.mode = mode;
  int rc = make_edges_d(thresholds, nthr, &p.ed); if (rc) return rc;
  rc = make_edges_f(thresholds, nthr, compare_mode, &p.ef); if (rc) return rc;
  if (dtype == EIGB200_F32 && mode == EIGB200_RATIO_NONE && compare_mode == EIGB200_CMP_F64) { /* ef already encodes the f64 compare */ }
  const int64_t Nout = mode == EIGB200_RATIO_NONE ? N : N - 1;
  // CTAs per sequence: enough to fill the machine, at least ~2048 values per CTA
  int64_t per_seq = ((int64_t)num_sms() * 8 + B - 1) / B;
  const int64_t max_ps = (Nout * inner + 2047) / 2048;
  if (per_seq > max_ps) per_seq = max_ps;
  if (per_seq < 1) per_seq = 1;
  p.chunk = (int)((Nout + per_seq - 1) / per_seq);
  per_seq = (Nout + p.chunk - 1) / p.chunk;
  dim3 grid((unsigned)per_seq, (unsigned)B);
  const size_t smem = inner <= 256 ? (size_t)inner * EIGB200_NSLOT * sizeof(int) : 0;
  if (dtype == EIGB200_F32) ratio_hist_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  else ratio_hist_kernel<double><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  EIGB_LAUNCH_CHECK("ratio_hist_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_softmax_eta(void* stream, const double* d_nu, const float* d_m, int64_t B, int64_t T, int H,
                                   double* d_eta, int32_t* d_counts, const double* thresholds, int nthr) {
  EIGB_CHECK_ARG(d_nu && d_m, "softmax_eta: null input");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 1 && H > 0, "softmax_eta: bad shape (T >= 2)");
  RatioParams p{};
  p.a = d_nu; p.out = d_eta; p.out_stride = 1; p.counts = d_counts; p.N = T; p.inner = H; p.mode = EIGB200_RATIO_CUR_OVER_NEXT; p.mscale = d_m;
  int rc = make_edges_d(thresholds, nthr, &p.ed); if (rc) return rc;
  rc = make_edges_f(thresholds, nthr, EIGB200_CMP_F64, &p.ef); if (rc) return rc;
  const int64_t Nout = T - 1;
  int64_t per_seq = ((int64_t)num_sms() * 8 + B - 1) / B;
  const int64_t max_ps = (Nout * H + 2047) / 2048;
  if (per_seq > max_ps) per_seq = max_ps;
  if (per_seq < 1) per_seq = 1;
  p.chunk = (int)((Nout + per_seq - 1) / per_seq);
  per_seq = (Nout + p.chunk - 1) / p.chunk;
  dim3 grid((unsigned)per_seq, (unsigned)B);
  const size_t smem = H <= 256 ? (size_t)H * EIGB200_NSLOT * sizeof(int) : 0;
  ratio_hist_kernel<double><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  EIGB_LAUNCH_CHECK("ratio_hist_kernel");
  return EIGB200_OK;
}

static int launch_count_moments(cudaStream_t st, const int32_t* d_counts, int64_t L, int64_t B, int64_t inner, int64_t* d_sum, int64_t* d_sumsq) {
  const int64_t n = inner * EIGB200_NSLOT;
  EIGB_CHECK_ARG(L > 0 && L <= 65535 && (n + 255) / 256 <= 65535, "count_moments: too many layers / columns");
  EIGB_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)(L * n) * sizeof(int64_t), st));
  EIGB_CUDA(cudaMemsetAsync(d_sumsq, 0, (size_t)(L * n) * sizeof(int64_t), st));
  dim3 grid((unsigned)((B + CM_ROWS - 1) / CM_ROWS), (unsigned)((n + 255) / 256), (unsigned)L);
  count_moments_kernel<<<grid, 256, 0, st>>>(d_counts, B, n, (unsigned long long*)d_sum, (unsigned long long*)d_sumsq);
  EIGB_LAUNCH_CHECK("count_moments_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_count_moments(void* stream, const int32_t* d_counts, int64_t B, int64_t inner, int64_t* d_sum, int64_t* d_sumsq) {
  EIGB_CHECK_ARG(d_counts && d_sum && d_sumsq && B > 0 && inner > 0, "count_moments: bad arguments");
  return launch_count_moments((cudaStream_t)stream, d_counts, 1, B, inner, d_sum, d_sumsq);
}

extern "C" int eigb200_count_moments_layers(void* stream, const int32_t* d_counts, int64_t L, int64_t B, int64_t inner, int64_t* d_sum, int64_t* d_sumsq) {
  EIGB_CHECK_ARG(d_counts && d_sum && d_sumsq && L > 0 && B > 0 && inner > 0, "count_moments_layers: bad arguments");
  return launch_count_moments((cudaStream_t)stream, d_counts, L, B, inner, d_sum, d_sumsq);
}

extern "C" int eigb200_mamba2_eig_partials(void* stream, const float* d_partials, int ngroups16, int64_t B, int64_t T,
                                           const float* d_dt_bias, const float* d_A_log, float* d_lam, int64_t lam_stride, int32_t* d_counts,
                                           const double* thresholds, int nthr, int compare_mode, float* d_rowstats, float ln_eps) {
  EIGB_CHECK_ARG(d_partials && d_dt_bias && d_A_log, "mamba2_eig_partials: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0 && T < (1LL << 31), "mamba2_eig_partials: bad shape");
  EIGB_CHECK_ARG(ngroups16 == 2 || ngroups16 == 4 || ngroups16 == 8 || ngroups16 == 16, "mamba2_eig_partials: d_model / 16 = %d must be 2, 4, 8 or 16", ngroups16);
  EIGB_CHECK_ARG(!d_lam || lam_stride >= 1, "mamba2_eig_partials: lam_stride must be >= 1");
  EdgesF e; double one = 1.0;
  if (d_counts) { int rc = make_edges_f(thresholds, nthr, compare_mode, &e); if (rc) return rc; } else make_edges_f(&one, 1, 0, &e);
  const int64_t rows = B * T;
  const float* dtb = d_dt_bias; const float* A = d_A_log;
  const unsigned grid = (unsigned)((rows + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  float2* rs = reinterpret_cast<float2*>(d_rowstats);
  switch (ngroups16) {
    case 2: k1_partials_kernel<2><<<grid, 256, 0, st>>>(d_partials, rows, (int)T, dtb, A, d_lam, lam_stride, d_counts, e, rs, ln_eps); break;
    case 4: k1_partials_kernel<4><<<grid, 256, 0, st>>>(d_partials, rows, (int)T, dtb, A, d_lam, lam_stride, d_counts, e, rs, ln_eps); break;
    case 8: k1_partials_kernel<8><<<grid, 256, 0, st>>>(d_partials, rows, (int)T, dtb, A, d_lam, lam_stride, d_counts, e, rs, ln_eps); break;
    default: k1_partials_kernel<16><<<grid, 256, 0, st>>>(d_partials, rows, (int)T, dtb, A, d_lam, lam_stride, d_counts, e, rs, ln_eps); break;
  }
  EIGB_LAUNCH_CHECK("k1_partials_kernel");
  return EIGB200_OK;
}
