// block_ops.cu -- the glue that carries activations from one layer's extractor to the next: token embedding, LayerNorm,
// depthwise causal conv + SiLU, GELU / residual / gating elementwise ops.  All HBM-bound streaming kernels: 128-bit
// coalesced accesses, one pass over the data, grid sized from the SM count.
//
// Reference operators: TokenEmbeddings.forward models/common.py:160-176; nn.LayerNorm models/mamba.py:321, :331,
// transformer.py:84-94; Conv1d(groups=C,padding=k-1)+SiLU+truncate models/attention.py:153-156, norm_attention.py:236-239;
// nn.GELU models/mamba.py:318; residual adds / y*silu(z) models/transformer.py:96-109.
#include "common.cuh"

namespace eigb200 {

// ---- embedding: LPR lanes per (b,t) row (8 for D <= 256: a warp gathers 4 rows per trip, 32 otherwise) ------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) embedding_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                                        const float* __restrict__ pos, float* __restrict__ out,
                                                        int64_t rows, int64_t T, int D, int64_t vocab, int* __restrict__ err,
                                                        float2* __restrict__ rowstats, float ln_eps) {
  constexpr int RPW = 32 / LPR;                                       // rows per warp trip
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = D >> 2;
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    const int64_t rc = ok ? r : rows - 1;
    int64_t id = ids[rc];
    if (id < 0 || id >= vocab) { if (err) *err = 1; id = 0; }
    const float4* w = reinterpret_cast<const float4*>(word + id * D);
    const float4* pp = pos ? reinterpret_cast<const float4*>(pos + (rc % T) * D) : nullptr;
    float4* o = reinterpret_cast<float4*>(out + rc * D);
    float s1 = 0.f, s2 = 0.f;
    const float shift = rowstats ? (__ldg(word + id * D) + (pos ? __ldg(pos + (rc % T) * D) : 0.f)) : 0.f;
    for (int c = l; c < nv; c += LPR) {
      float4 v = __ldg(w + c);
      if (pp) { const float4 q = __ldg(pp + c); v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
      if (ok) stg_stream_f4(o + c, v);
      if (rowstats) {
        const float a = v.x - shift, bb = v.y - shift, cc = v.z - shift, d = v.w - shift;
        s1 += (a + bb) + (cc + d); s2 += (a * a + bb * bb) + (cc * cc + d * d);
      }
    }
    if (rowstats) {
#pragma unroll
      for (int ofs = LPR / 2; ofs >= 1; ofs >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, ofs); s2 += __shfl_xor_sync(0xffffffffu, s2, ofs); }
      if (l == 0 && ok) {
        const float md = s1 / (float)D;
        rowstats[r] = make_float2(shift + md, rsqrtf(fmaxf(s2 / (float)D - md * md, 0.f) + ln_eps));
      }
    }
  }
}

// D = 128 (the C2 / MQAR family): 8 lanes per row, the row's four float4 gathers (L2 hits) of a lane issued together and the token ids of the next trip fetched
// one trip ahead -- the generic kernel's dependent id -> gather -> store chain per 16 bytes left it latency-bound at 59 % of the copy peak
__global__ void __launch_bounds__(256) embedding128_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                                           const float* __restrict__ pos, float* __restrict__ out,
                                                           int64_t rows, int64_t T, int64_t vocab, float2* __restrict__ rowstats, float ln_eps) {
  constexpr int D = 128, LPR = 8, RPW = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t r0 = warp * RPW;
  int64_t id_next = 0;
  if (r0 < rows) id_next = ids[min(r0 + sub, rows - 1)];
  for (; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    const int64_t rc = ok ? r : rows - 1;
    int64_t id = id_next;
    const int64_t rn = r0 + nwarps * RPW;
    if (rn < rows) id_next = ids[min(rn + sub, rows - 1)];
    if (id < 0 || id >= vocab) id = 0;
    const float4* w = reinterpret_cast<const float4*>(word + id * D);
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(w + l + LPR * k);
    if (pos) {
      const float4* pp = reinterpret_cast<const float4*>(pos + (rc % T) * D);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float4 q = __ldg(pp + l + LPR * k); v[k].x += q.x; v[k].y += q.y; v[k].z += q.z; v[k].w += q.w; }
    }
    float4* o = reinterpret_cast<float4*>(out + rc * D);
    if (ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) stg_stream_f4(o + l + LPR * k, v[k]);
    }
    if (rowstats) {
      const float shift = __shfl_sync(0xffffffffu, v[0].x, sub * LPR);           // the row's first element (lane l = 0 holds columns 0-3)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = v[k].x - shift, bb = v[k].y - shift, cc = v[k].z - shift, d = v[k].w - shift;
        s1 += (a + bb) + (cc + d); s2 += (a * a + bb * bb) + (cc * cc + d * d);
      }
#pragma unroll
      for (int ofs = LPR / 2; ofs >= 1; ofs >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, ofs); s2 += __shfl_xor_sync(0xffffffffu, s2, ofs); }
      if (l == 0 && ok) {
        const float md = s1 / (float)D;
        rowstats[r] = make_float2(shift + md, rsqrtf(fmaxf(s2 / (float)D - md * md, 0.f) + ln_eps));
      }
    }
  }
}

static void launch_embedding(cudaStream_t st, const int64_t* ids, const float* word, const float* pos, float* out, int64_t rows, int64_t T, int D,
                             int64_t vocab, float2* rowstats, float ln_eps) {
  if (D == 128) {
    int64_t g = (rows + 31) / 32; const int64_t cap = (int64_t)num_sms() * 16; if (g > cap) g = cap;
    embedding128_kernel<<<(unsigned)g, 256, 0, st>>>(ids, word, pos, out, rows, T, vocab, rowstats, ln_eps);
    return;
  }
  const int lpr = D <= 256 ? 8 : 32;
  const int64_t rows_per_cta = 8 * (32 / lpr);
  int64_t g = (rows + rows_per_cta - 1) / rows_per_cta; const int64_t cap = (int64_t)num_sms() * 16; if (g > cap) g = cap;
  if (lpr == 8) embedding_kernel<8><<<(unsigned)g, 256, 0, st>>>(ids, word, pos, out, rows, T, D, vocab, nullptr, rowstats, ln_eps);
  else embedding_kernel<32><<<(unsigned)g, 256, 0, st>>>(ids, word, pos, out, rows, T, D, vocab, nullptr, rowstats, ln_eps);
}

// (mean, rstd) of every row: one warp per row, shifted moments
__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ x, int64_t rows, int D, float eps, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = D >> 2;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    const float shift = __ldg(x + r * D);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < nv; c += 32) {
      const float4 v = ldg_stream_f4(xr + c);
      const float a = v.x - shift, bb = v.y - shift, cc = v.z - shift, d = v.w - shift;
      s1 += (a + bb) + (cc + d); s2 += (a * a + bb * bb) + (cc * cc + d * d);
    }
#pragma unroll
    for (int ofs = 16; ofs >= 1; ofs >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, ofs); s2 += __shfl_xor_sync(0xffffffffu, s2, ofs); }
    if (lane == 0) { const float md = s1 / (float)D; stats[r] = make_float2(shift + md, rsqrtf(fmaxf(s2 / (float)D - md * md, 0.f) + eps)); }
  }
}

// ---- LayerNorm: one warp per row, row held in registers (D <= 128*NV4) -------------------------------------------
template <int NV4>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                        float eps, float* __restrict__ out, int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = D >> 2;
  const float invD = 1.f / (float)D;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    float4 v[NV4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) { v[i] = ldg_stream_f4(xr + c); sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
      else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mu = sum * invD;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) { const float a = v[i].x - mu, bb = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu; sq += (a * a + bb * bb) + (cc * cc + d * d); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * invD + eps);
    float4* orow = reinterpret_cast<float4*>(out + r * D);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c), be = __ldg(reinterpret_cast<const float4*>(b) + c);
        float4 o4;
        o4.x = (v[i].x - mu) * rstd * g.x + be.x; o4.y = (v[i].y - mu) * rstd * g.y + be.y;
        o4.z = (v[i].z - mu) * rstd * g.z + be.z; o4.w = (v[i].w - mu) * rstd * g.w + be.w;
        orow[c] = o4;
      }
    }
  }
}

// ---- depthwise causal conv + SiLU: thread per channel, sliding window over a chunk of tokens -------------------
constexpr int CONV_TCH = 64;
__global__ void __launch_bounds__(128) conv_silu_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                                                        int k, float* __restrict__ out, int64_t ldo, int64_t T, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.y * CONV_TCH, t1 = min(T, t0 + CONV_TCH);
  // taps right-aligned in an 8-slot window: slot 7 is the current token, slot 7-i the token i steps back
  float wk[8], win[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wk[j] = (j >= 8 - k) ? w[(size_t)c * k + (j - (8 - k))] : 0.f; win[j] = 0.f; }
  const float bb = bias[c];
  const float* xr = x + (b * T) * ldx + c;
#pragma unroll
  for (int j = 1; j < 8; ++j) {                       // prime slots 1..7 with x_{t0-7..t0-1} (zeros before the sequence start)
    const int64_t t = t0 - 8 + j;
    win[j] = (t >= 0 && j >= 8 - k) ? __ldg(xr + t * ldx) : 0.f;
  }
  float* orow = out + (b * T) * ldo + c;
  // 8 tokens per trip: their loads are issued together (one load in flight per thread left the kernel latency-bound at 40 % of the copy peak)
  constexpr int U = 8;
  int64_t t = t0;
  for (; t + U <= t1; t += U) {
    float xn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xn[u] = ldg_stream_f1(xr + (t + u) * ldx);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < 7; ++j) win[j] = win[j + 1];
      win[7] = xn[u];
      float acc = bb;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(wk[j], win[j], acc);
      orow[(t + u) * ldo] = silu_f(acc);
    }
  }
  for (; t < t1; ++t) {
#pragma unroll
    for (int j = 0; j < 7; ++j) win[j] = win[j + 1];
    win[7] = __ldg(xr + t * ldx);
    float acc = bb;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(wk[j], win[j], acc);
    orow[t * ldo] = silu_f(acc);
  }
}

// ---- elementwise ---------------------------------------------------------------------------------------------------
enum { EW_ADD = 0, EW_MUL_SILU = 1, EW_GELU = 2 };
template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 u = ldg_stream_f4(reinterpret_cast<const float4*>(a) + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), r;
    if (OP != EW_GELU) v = ldg_stream_f4(reinterpret_cast<const float4*>(b) + i);
    if (OP == EW_ADD) { r.x = u.x + v.x; r.y = u.y + v.y; r.z = u.z + v.z; r.w = u.w + v.w; }
    else if (OP == EW_MUL_SILU) { r.x = u.x * silu_f(v.x); r.y = u.y * silu_f(v.y); r.z = u.z * silu_f(v.z); r.w = u.w * silu_f(v.w); }
    else { r.x = gelu_f(u.x); r.y = gelu_f(u.y); r.z = gelu_f(u.z); r.w = gelu_f(u.w); }
    reinterpret_cast<float4*>(out)[i] = r;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float u = a[i], v = OP != EW_GELU ? b[i] : 0.f;
    out[i] = OP == EW_ADD ? u + v : (OP == EW_MUL_SILU ? u * silu_f(v) : gelu_f(u));
  }
}

// out[m, j] = a[m, j] * s[j]  (the D * u feed-through of LRU / S5, models/lru.py:97, s5.py:247-248)
__global__ void __launch_bounds__(256) scale_cols_kernel(const float* __restrict__ a, const float* __restrict__ s, float* __restrict__ out, int64_t n, int cols) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] * __ldg(s + (int)(i % cols));
}

static int ew_grid(int64_t n) {
  int64_t g = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}


// SSD_LTI.forward (models/mamba.py:262-281): dt = softplus(dt_raw + dt_bias[h]) tiled over the d_state columns (column j belongs to head j / khead), B <- dt * B,
// in place on the conv'd projection buffer.  One thread per (row, state column).
__global__ void lti_scale_b_kernel(float* __restrict__ buf, int64_t ld, int col_b, int col_dt, const float* __restrict__ dt_bias,
                                   int64_t rows, int N, int khead) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * N) return;
  const int64_t m = idx / N;
  const int j = (int)(idx - m * N);
  const float dt = softplus_f(buf[m * ld + col_dt] + dt_bias[j / khead]);
  buf[m * ld + col_b + j] *= dt;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_embedding(void* stream, const int64_t* d_ids, const float* d_word, const float* d_pos, float* d_out,
                                 int64_t B, int64_t T, int D, int64_t vocab) {
  EIGB_CHECK_ARG(d_ids && d_word && d_out, "embedding: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0 && D > 0 && D % 4 == 0 && vocab > 0, "embedding: bad shape (D %% 4 == 0 required)");
  launch_embedding((cudaStream_t)stream, d_ids, d_word, d_pos, d_out, B * T, T, D, vocab, nullptr, 0.f);
  EIGB_LAUNCH_CHECK("embedding_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_embedding_stats(void* stream, const int64_t* d_ids, const float* d_word, const float* d_pos, float* d_out,
                                       int64_t B, int64_t T, int D, int64_t vocab, float* d_rowstats, float ln_eps) {
  EIGB_CHECK_ARG(d_ids && d_word && d_out && d_rowstats, "embedding_stats: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0 && D > 0 && D % 4 == 0 && vocab > 0, "embedding_stats: bad shape (D %% 4 == 0 required)");
  launch_embedding((cudaStream_t)stream, d_ids, d_word, d_pos, d_out, B * T, T, D, vocab, reinterpret_cast<float2*>(d_rowstats), ln_eps);
  EIGB_LAUNCH_CHECK("embedding_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_rowstats(void* stream, const float* d_x, int64_t rows, int D, float eps, float* d_stats) {
  EIGB_CHECK_ARG(d_x && d_stats && rows > 0 && D > 0 && D % 4 == 0, "rowstats: bad arguments (D %% 4 == 0 required)");
  int64_t g = (rows + 7) / 8; const int64_t cap = (int64_t)num_sms() * 16; if (g > cap) g = cap;
  rowstats_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(d_x, rows, D, eps, reinterpret_cast<float2*>(d_stats));
  EIGB_LAUNCH_CHECK("rowstats_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_layernorm(void* stream, const float* d_x, const float* d_w, const float* d_b, float eps, float* d_out, int64_t rows, int D) {
  EIGB_CHECK_ARG(d_x && d_w && d_b && d_out, "layernorm: null pointer");
  EIGB_CHECK_ARG(rows > 0 && D > 0 && D % 4 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 4 and <= 2048", D);
  int64_t g = (rows + 7) / 8; const int64_t cap = (int64_t)num_sms() * 8; if (g > cap) g = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv4 = (D / 4 + 31) / 32;
  if (nv4 <= 1) layernorm_kernel<1><<<(unsigned)g, 256, 0, st>>>(d_x, d_w, d_b, eps, d_out, rows, D);
  else if (nv4 <= 2) layernorm_kernel<2><<<(unsigned)g, 256, 0, st>>>(d_x, d_w, d_b, eps, d_out, rows, D);
  else if (nv4 <= 4) layernorm_kernel<4><<<(unsigned)g, 256, 0, st>>>(d_x, d_w, d_b, eps, d_out, rows, D);
  else if (nv4 <= 8) layernorm_kernel<8><<<(unsigned)g, 256, 0, st>>>(d_x, d_w, d_b, eps, d_out, rows, D);
  else layernorm_kernel<16><<<(unsigned)g, 256, 0, st>>>(d_x, d_w, d_b, eps, d_out, rows, D);
  EIGB_LAUNCH_CHECK("layernorm_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_conv_silu(void* stream, const float* d_x, int64_t ldx, const float* d_w, const float* d_b, int k,
                                 float* d_out, int64_t ldo, int64_t B, int64_t T, int C) {
  EIGB_CHECK_ARG(d_x && d_w && d_b && d_out, "conv_silu: null pointer");
  EIGB_CHECK_ARG(k >= 1 && k <= 8, "conv_silu: kernel size %d not in 1..8", k);
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && C > 0, "conv_silu: bad shape");
  dim3 grid((C + 127) / 128, (unsigned)((T + CONV_TCH - 1) / CONV_TCH), (unsigned)B);
  conv_silu_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_x, ldx, d_w, d_b, k, d_out, ldo, T, C);
  EIGB_LAUNCH_CHECK("conv_silu_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_add(void* stream, const float* d_a, const float* d_b, float* d_out, int64_t n) {
  EIGB_CHECK_ARG(d_a && d_b && d_out && n > 0, "add: bad arguments");
  ew_kernel<EW_ADD><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, d_out, n);
  EIGB_LAUNCH_CHECK("ew_kernel<add>");
  return EIGB200_OK;
}
extern "C" int eigb200_mul_silu(void* stream, const float* d_y, const float* d_z, float* d_out, int64_t n) {
  EIGB_CHECK_ARG(d_y && d_z && d_out && n > 0, "mul_silu: bad arguments");
  ew_kernel<EW_MUL_SILU><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_y, d_z, d_out, n);
  EIGB_LAUNCH_CHECK("ew_kernel<mul_silu>");
  return EIGB200_OK;
}
extern "C" int eigb200_gelu(void* stream, const float* d_x, float* d_out, int64_t n) {
  EIGB_CHECK_ARG(d_x && d_out && n > 0, "gelu: bad arguments");
  ew_kernel<EW_GELU><<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_x, nullptr, d_out, n);
  EIGB_LAUNCH_CHECK("ew_kernel<gelu>");
  return EIGB200_OK;
}
extern "C" int eigb200_scale_cols(void* stream, const float* d_a, const float* d_s, float* d_out, int64_t rows, int cols) {
  EIGB_CHECK_ARG(d_a && d_s && d_out && rows > 0 && cols > 0, "scale_cols: bad arguments");
  scale_cols_kernel<<<ew_grid(rows * cols), 256, 0, (cudaStream_t)stream>>>(d_a, d_s, d_out, rows * cols, cols);
  EIGB_LAUNCH_CHECK("scale_cols_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_lti_scale_b(void* stream, float* d_buf, int64_t ld, int col_b, int col_dt, const float* d_dt_bias,
                                   int64_t rows, int N, int khead) {
  EIGB_CHECK_ARG(d_buf && d_dt_bias, "lti_scale_b: null pointer");
  EIGB_CHECK_ARG(rows > 0 && N > 0 && khead > 0 && N % khead == 0 && col_b >= 0 && col_dt >= 0 && ld > col_dt && ld >= col_b + N, "lti_scale_b: bad arguments");
  const int64_t n = rows * N;
  lti_scale_b_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_buf, ld, col_b, col_dt, d_dt_bias, rows, N, khead);
  EIGB_LAUNCH_CHECK("lti_scale_b_kernel");
  return EIGB200_OK;
}
