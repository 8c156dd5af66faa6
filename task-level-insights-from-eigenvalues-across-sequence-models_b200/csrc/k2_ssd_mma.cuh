// k2_ssd_mma.cuh -- K2b in its chunked ("dual") form on the warp-level tensor-core path (mma.sync.m16n8k8 tf32, error-compensated 3xTF32).
// Included by k2_ssd_scan.cu (needs SsdParams, silu_fast_f, cp_async_16_zfill).
//
// Same recurrence as the scan kernels (models/mamba.py:138-150, mamba_chunk_scan_combined):
//   S_t[p,n] = exp(dt_t A) S_{t-1}[p,n] + dt_t xc_t[p] B_t[n];    y_t[p] = sum_n C_t[n] S_t[p,n] + D xc_t[p],    xc = SiLU(conv(x))
// unrolled over a chunk of Q = 16 tokens with cum_t = sum_{i<=t} dt_i A (inclusive, <= 0) and S0 the state at the chunk start:
//   Y^T[p][t] = sum_n S0[p][n] Cs[n][t]  +  sum_s xc[p][s] Gm[s][t]                 Cs[n][t] = exp(cum_t) C_t[n]
//   S_Q[p][n] = exp(cum_15) S0[p][n]     +  sum_s xc[p][s] Bt[s][n]                 Gm[s][t] = [s<=t] exp(cum_t - cum_s) dt_s (C_t . B_s) + [s==t] D
//                                                                                   Bt[s][n] = exp(cum_15 - cum_s) dt_s B_s[n]
// Every exponential has a non-positive argument, so nothing is rescaled and nothing can overflow.  Channels are the M dimension of the MMAs:
// a warp owns 32 channels (two 16-row tiles) of one (sequence, head), its state tile lives in the accumulator registers of the third product and
// is fed back as the A operand of the first one without a shuffle (the accumulator's column order is a permutation of the K index, applied to
// the rows of Cs instead).  Per (16 channels x 16 tokens): 33 mma.sync (3 per product term, the all-zero block of Gm skipped) replace 8192 FMAs.
//
// CTA = PB channels of one (b, h), PB threads (one warp per 32 channels).  Per chunk: every thread convolves its own channel (rolling window in
// registers) into u_s, the CTA builds the three 16x16 operand matrices (hi / lo tf32 parts, stored fragment-major so that a thread reads its
// B fragments with 128-bit loads), then each warp runs its MMAs and writes y through a warp-private staging tile with full-line stores.
// Software pipeline: iteration c computes chunk c, builds the operands of chunk c+1 and convolves the [B|C] rows of chunk c+2 while the raw rows of
// chunk c+3 and the x tile of chunk c+1 are in flight (cp.async); one __syncthreads per chunk.
//
// STATUS (measured on B200, BASELINE C2 shape: 4096 sequences x 512 tokens, 1 head x 128 channels, d_state 16): 0.885 ms per layer against 0.793 ms
// for the recurrent kernel ssd_scan_v3 -- bit-for-bit the same tests pass, but with one head of 128 channels the per-chunk operand preparation
// (conv of B / C, the 16 x 16 Gram matrix, the tf32 splits) is amortised over only four warps and the 3xTF32 error compensation triples the MMAs:
// 813 issued instructions per (32 channels x 16 tokens) against 1176 for the scan, at 52 % instead of 70 % issue utilisation.  It is therefore NOT
// the default; EIGB200_SSD_FORM=mma selects it (tests/test_scans_gpu.py runs both forms).  It should win for heads with more channels per (B, C) group.
#pragma once

namespace eigb200 {

constexpr int SM_Q = 16;                 // tokens per chunk
constexpr int SM_BCS = 36;               // row stride of the conv'd [B|C] tile (floats)
constexpr int SM_YS = 36;                // row stride of the warp-private y staging tile

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tf32 split by truncation: the tensor core reads only the upper 19 bits of an operand register, so `hi` is the fp32 value itself (no
// instruction) and lo = v - trunc19(v) is exact (2 instructions per element; the dropped lo x lo term and lo's own truncation are ~2^-21 relative)
__device__ __forceinline__ void split_tf32_u(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v);
  lo = __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// D += A(16x8: hi, lo) * B(8x8: hi, lo), error-compensated: small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32_16x8x8(d, al, bh0, bh1);
  mma_tf32_16x8x8(d, ah, bl0, bl1);
  mma_tf32_16x8x8(d, ah, bh0, bh1);
}

template <int PB, bool CONV>
__global__ void __launch_bounds__(PB, 512 / PB) ssd_chunk_mma_kernel(const SsdParams p) {
  constexpr int N = 16, Q = SM_Q, NW = PB / 32;
  constexpr int US = PB + 8;                                         // u_s row stride: 8 mod 32 -> conflict-free A-fragment loads
  constexpr int RAW_ROWS = Q + SSD_HIST;
  constexpr int RAW_F4 = RAW_ROWS * 8;                               // 16-byte pieces of the raw [B|C] rows of a chunk
  constexpr int RAW_PER_T = (RAW_F4 + PB - 1) / PB;
  __shared__ __align__(16) float raw_s[2][RAW_ROWS][2 * N];          // raw [B|C] rows (+3 history rows), chunk k in buffer k & 1
  __shared__ __align__(16) float bc_s[2][Q][SM_BCS];                 // conv'd [B|C]
  __shared__ __align__(16) float2 dd_s[2][Q];                        // (dt_t, cum_t)
  __shared__ float dec_s[2];                                         // exp(cum_15) of the chunk whose operands sit in op_s[buf]
  __shared__ __align__(16) float op_s[2][3][2][2][32][4];           // [buffer][Cs, Gm, Bt][hi, lo][k-step][lane][n-tile 0: b0 b1, n-tile 1: b0 b1]
  __shared__ __align__(16) float u_s[2][Q][US];                      // raw x lands here (cp.async) and is convolved in place: [token][channel]
  __shared__ __align__(16) float y_s[NW][Q][SM_YS];

  const int P = p.P;
  const int b = blockIdx.z, h = blockIdx.y, pblk = blockIdx.x;
  const int g_ = h / (p.H / p.G);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;                         // mma fragment coordinates
  const float Ah = p.fused ? -expf(p.A[h]) : p.A[h];
  const float Dh = p.D ? p.D[h] : 0.f;
  const float dtb = p.fused ? p.dt_bias[h] : 0.f;
  const int64_t T = p.T;
  const int nch = (int)((T + Q - 1) / Q);

  float cw[4] = {0.f, 0.f, 0.f, 1.f}, cb = 0.f;
  float bw[4] = {0.f, 0.f, 0.f, 1.f}, bbias = 0.f;
  const int bc_col = tid & 31;
  if (CONV) {
    const int HP = p.H * P, GN = p.G * N;
    const int ch = h * P + pblk * PB + tid;
#pragma unroll
    for (int j = 0; j < 4; ++j) cw[j] = (j >= 4 - p.kconv) ? p.conv_w[(size_t)ch * p.kconv + j - (4 - p.kconv)] : 0.f;
    cb = p.conv_b[ch];
    const int chb = HP + (bc_col < N ? g_ * N + bc_col : GN + g_ * N + (bc_col - N));
#pragma unroll
    for (int j = 0; j < 4; ++j) bw[j] = (j >= 4 - p.kconv) ? p.conv_w[(size_t)chb * p.kconv + j - (4 - p.kconv)] : 0.f;
    bbias = p.conv_b[chb];
  }

  const size_t rowbase = (size_t)b * p.T;
  // x tile of this WARP (32 channels x 16 tokens = 16 rows x 8 pieces of 16 bytes): lane copies pieces (row lane/8 + 4k, piece lane%8), k = 0..3
  const float* xsrc = p.x + (rowbase + (lane >> 3)) * (size_t)p.ldx + (size_t)h * P + pblk * PB + warp * 32 + (lane & 7) * 4;
  const size_t x_step4 = 4 * (size_t)p.ldx, x_step_chunk = (size_t)Q * p.ldx;
  // raw [B|C] pieces of this thread: e = tid + k PB -> row e / 8 (token t0 - 3 + row), piece e % 8 (0-3: B, 4-7: C)
  const float* rsrc[RAW_PER_T];
#pragma unroll
  for (int k = 0; k < RAW_PER_T; ++k) {
    const int e = tid + k * PB, q = e & 7;
    rsrc[k] = ((q < 4) ? p.Bm + (size_t)g_ * N + 4 * q : p.Cm + (size_t)g_ * N + 4 * (q - 4)) + rowbase * (size_t)p.ldbc;
  }
  float* ydst = p.y + (rowbase + (lane >> 3)) * (size_t)p.ldy + (size_t)h * P + pblk * PB + warp * 32 + (lane & 7) * 4;
  const size_t y_step4 = 4 * (size_t)p.ldy, y_step_chunk = (size_t)Q * p.ldy;

  const float* xnext = xsrc;                                         // running source pointers: the chunks are requested in order
  auto x_async = [&](int c, int buf) {                               // raw x of chunk c -> this warp's columns of u_s[buf]
    if (c < nch) {
      const int64_t t0 = (int64_t)c * Q;
      float* dst = &u_s[buf][lane >> 3][warp * 32 + (lane & 7) * 4];
      if (t0 + Q <= T) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cp_async_16_zfill(dst + k * 4 * US, xnext + k * x_step4, true);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool ok = t0 + (lane >> 3) + 4 * k < T;
          cp_async_16_zfill(dst + k * 4 * US, ok ? xnext + k * x_step4 : p.x, ok);
        }
      }
      xnext += x_step_chunk;
    }
  };
  const size_t r_step_chunk = (size_t)Q * p.ldbc;
#pragma unroll
  for (int k = 0; k < RAW_PER_T; ++k) rsrc[k] += ((int64_t)((tid + k * PB) >> 3) - SSD_HIST) * p.ldbc;   // row of chunk 0 (may point before the sequence)
  auto raw_async = [&](int c) {                                      // rows [16c-3, 16c+16) of [B|C] -> raw_s[c & 1], zero-filled outside the sequence
    if (c < nch) {
      const int64_t t0 = (int64_t)c * Q;
      float4* dst = reinterpret_cast<float4*>(&raw_s[c & 1][0][0]) + tid;
      if (c > 0 && t0 + Q <= T) {
#pragma unroll
        for (int k = 0; k < RAW_PER_T; ++k)
          if (tid + k * PB < RAW_F4) cp_async_16_zfill(dst + k * PB, rsrc[k], true);
      } else {
#pragma unroll
        for (int k = 0; k < RAW_PER_T; ++k) {
          const int e = tid + k * PB;
          if (e < RAW_F4) {
            const int64_t t = t0 - SSD_HIST + (e >> 3);
            const bool ok = t >= 0 && t < T;
            cp_async_16_zfill(dst + k * PB, ok ? rsrc[k] : p.Bm, ok);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < RAW_PER_T; ++k) rsrc[k] += r_step_chunk;
    }
  };
  float dtreg = 0.f;
  auto dt_fetch = [&](int c) {
    const int64_t t = (int64_t)c * Q + lane;
    dtreg = 0.f;
    if (warp == 0 && lane < Q && c < nch && t < T) dtreg = __ldg(p.dt + (rowbase + t) * p.lddt + h);
  };

  // conv'd [B|C] rows and (dt, cum) of chunk c from raw_s[c & 1] (landed and published by a barrier) -> bc_s[c & 1], dd_s[c & 1]
  auto conv_bc = [&](int c) {
    if (c >= nch) return;
    const int buf = c & 1;
    const int tc = (int)min((int64_t)Q, T - (int64_t)c * Q);
#pragma unroll
    for (int k = 0; k < Q * 2 * N / PB; ++k) {
      const int r = (tid + k * PB) >> 5;
      float v;
      if (CONV) {
        v = fmaf(bw[3], raw_s[buf][r + 3][bc_col], fmaf(bw[2], raw_s[buf][r + 2][bc_col], fmaf(bw[1], raw_s[buf][r + 1][bc_col], fmaf(bw[0], raw_s[buf][r][bc_col], bbias))));
        v = silu_fast_f(v);
      } else v = raw_s[buf][r + 3][bc_col];
      bc_s[buf][r][bc_col] = v;
    }
    if (warp == 0) {
      float d = 0.f;
      if (lane < tc) d = p.fused ? softplus_f(dtreg + dtb) : dtreg;
      float cum = d * Ah;                                            // lanes >= tc contribute 0: padded tokens neither decay nor feed the state
#pragma unroll
      for (int o = 1; o < Q; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += v; }
      if (lane < Q) dd_s[buf][lane] = make_float2(d, cum);
    }
  };

  // the three operand matrices of chunk c from bc_s / dd_s[c & 1], split into tf32 hi / lo and scattered into fragment order -> op_s[c & 1]
  auto build_op = [&](int c) {
    if (c >= nch) return;
    const int buf = c & 1;
    const float cum_last = dd_s[buf][Q - 1].y;
    if (tid == 0) dec_s[buf] = fast_exp_f(cum_last);
#pragma unroll
    for (int k = 0; k < Q * Q / PB; ++k) {
      const int e = tid + k * PB;
      const int i = e >> 4, j = e & 15;
      const float2 di = dd_s[buf][i];
      const float cj = dd_s[buf][j].y;
      const int ldn = (j & 7) * 4 + (i & 3), idxn = (j >> 3) * 2 + ((i & 7) >> 2), ks = i >> 3;   // K index i in natural order, N index j
      uint32_t hi, lo;
      {                                                              // Gm[s = i][t = j]
        float v = 0.f;
        if (i <= j) {
          const float4* Bs = reinterpret_cast<const float4*>(&bc_s[buf][i][0]);
          const float4* Ct = reinterpret_cast<const float4*>(&bc_s[buf][j][N]);
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < N / 4; ++q) {
            const float4 bv = Bs[q], cv = Ct[q];
            dot = fmaf(bv.x, cv.x, dot); dot = fmaf(bv.y, cv.y, dot); dot = fmaf(bv.z, cv.z, dot); dot = fmaf(bv.w, cv.w, dot);
          }
          v = fast_exp_f(cj - di.y) * di.x * dot;
          if (i == j) v += Dh;
        }
        split_tf32_u(v, hi, lo);
        op_s[buf][1][0][ks][ldn][idxn] = __uint_as_float(hi);
        op_s[buf][1][1][ks][ldn][idxn] = __uint_as_float(lo);
      }
      {                                                              // Bt[s = i][n = j]
        const float v = fast_exp_f(cum_last - di.y) * di.x * bc_s[buf][i][j];
        split_tf32_u(v, hi, lo);
        op_s[buf][2][0][ks][ldn][idxn] = __uint_as_float(hi);
        op_s[buf][2][1][ks][ldn][idxn] = __uint_as_float(lo);
      }
      {                                                              // Cs[n = i][t = j]; the K slot of n follows the accumulator's column order
        const float v = fast_exp_f(cj) * bc_s[buf][j][N + i];
        split_tf32_u(v, hi, lo);
        const int r = i & 7;
        const int ld = (j & 7) * 4 + (r >> 1), idx = (j >> 3) * 2 + (r & 1);
        op_s[buf][0][0][ks][ld][idx] = __uint_as_float(hi);
        op_s[buf][0][1][ks][ld][idx] = __uint_as_float(lo);
      }
    }
  };

  // this thread's channel of chunk c: raw x (landed in u_s[c & 1], copied by this warp) -> SiLU(conv(x)) in place
  float hist0 = 0.f, hist1 = 0.f, hist2 = 0.f;                       // raw x of tokens t-3, t-2, t-1
  auto prepass = [&](int c) {
    if (c >= nch) return;
    const int buf = c & 1;
    const int tc = (int)min((int64_t)Q, T - (int64_t)c * Q);
    float* col = &u_s[buf][0][tid];
    float xr[Q + 3];
    xr[0] = hist0; xr[1] = hist1; xr[2] = hist2;
#pragma unroll
    for (int j = 0; j < Q; ++j) xr[j + 3] = col[j * US];
    hist0 = xr[Q]; hist1 = xr[Q + 1]; hist2 = xr[Q + 2];
    if (CONV) {
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const float v = silu_fast_f(fmaf(cw[3], xr[j + 3], fmaf(cw[2], xr[j + 2], fmaf(cw[1], xr[j + 1], fmaf(cw[0], xr[j], cb)))));
        col[j * US] = v;
      }
    }
    if (tc < Q) {                                                    // padded tokens of the last chunk feed nothing
#pragma unroll
      for (int j = 0; j < Q; ++j) if (j >= tc) col[j * US] = 0.f;
    }
  };

  float S[2][2][4];                                                  // state tile [m-tile][n-tile][accumulator register]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) S[mt][nt][j] = 0.f;

  // ---- compute chunk c from op_s / u_s / dec_s[c & 1] ---------------------------------------------------------------------------------------
  auto compute = [&](int c) {
    const int buf = c & 1;
    const int64_t t0 = (int64_t)c * Q;
    float Y[2][2][4];
    {                                                                // (1) Y = S0 . Cs
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint4 vh = *reinterpret_cast<const uint4*>(&op_s[buf][0][0][ks][lane][0]);
        const uint4 vl = *reinterpret_cast<const uint4*>(&op_s[buf][0][1][ks][lane][0]);
        bh[ks][0] = vh.x; bh[ks][1] = vh.y; bh[ks][2] = vh.z; bh[ks][3] = vh.w;
        bl[ks][0] = vl.x; bl[ks][1] = vl.y; bl[ks][2] = vl.z; bl[ks][3] = vl.w;
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) Y[mt][nt][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {                             // accumulator (row, col 2 tig + e) -> A fragment (row, k slot tig + 4 e)
          uint32_t ah[4], al[4];
          split_tf32_u(S[mt][ks][0], ah[0], al[0]);
          split_tf32_u(S[mt][ks][2], ah[1], al[1]);
          split_tf32_u(S[mt][ks][1], ah[2], al[2]);
          split_tf32_u(S[mt][ks][3], ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma3(Y[mt][nt], ah, al, bh[ks][2 * nt], bh[ks][2 * nt + 1], bl[ks][2 * nt], bl[ks][2 * nt + 1]);
        }
      }
    }
    {                                                                // (2) Y += xc . Gm;   (3) S = exp(cum_15) S + xc . Bt
      uint32_t gh[2][4], gl[2][4], th[2][4], tl[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint4 a = *reinterpret_cast<const uint4*>(&op_s[buf][1][0][ks][lane][0]);
        const uint4 b2 = *reinterpret_cast<const uint4*>(&op_s[buf][1][1][ks][lane][0]);
        const uint4 c2 = *reinterpret_cast<const uint4*>(&op_s[buf][2][0][ks][lane][0]);
        const uint4 d2 = *reinterpret_cast<const uint4*>(&op_s[buf][2][1][ks][lane][0]);
        gh[ks][0] = a.x; gh[ks][1] = a.y; gh[ks][2] = a.z; gh[ks][3] = a.w;
        gl[ks][0] = b2.x; gl[ks][1] = b2.y; gl[ks][2] = b2.z; gl[ks][3] = b2.w;
        th[ks][0] = c2.x; th[ks][1] = c2.y; th[ks][2] = c2.z; th[ks][3] = c2.w;
        tl[ks][0] = d2.x; tl[ks][1] = d2.y; tl[ks][2] = d2.z; tl[ks][3] = d2.w;
      }
      const float dec = dec_s[buf];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* ub = &u_s[buf][tig][warp * 32 + mt * 16 + gid];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) S[mt][nt][j] *= dec;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t ah[4], al[4];
          split_tf32_u(ub[(8 * ks) * US], ah[0], al[0]);
          split_tf32_u(ub[(8 * ks) * US + 8], ah[1], al[1]);
          split_tf32_u(ub[(8 * ks + 4) * US], ah[2], al[2]);
          split_tf32_u(ub[(8 * ks + 4) * US + 8], ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (!(ks == 1 && nt == 0))                               // Gm[s >= 8][t < 8] = 0
              mma3(Y[mt][nt], ah, al, gh[ks][2 * nt], gh[ks][2 * nt + 1], gl[ks][2 * nt], gl[ks][2 * nt + 1]);
            mma3(S[mt][nt], ah, al, th[ks][2 * nt], th[ks][2 * nt + 1], tl[ks][2 * nt], tl[ks][2 * nt + 1]);
          }
        }
        float* yw = &y_s[warp][2 * tig][mt * 16 + gid];              // accumulator -> staging tile [token][channel]
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          yw[(8 * nt) * SM_YS] = Y[mt][nt][0];
          yw[(8 * nt + 1) * SM_YS] = Y[mt][nt][1];
          yw[(8 * nt) * SM_YS + 8] = Y[mt][nt][2];
          yw[(8 * nt + 1) * SM_YS + 8] = Y[mt][nt][3];
        }
      }
    }
    __syncwarp();
    {                                                                // 4 tokens x 128 contiguous bytes per store instruction
      float* yb = ydst;
      ydst += y_step_chunk;
      const float* ys = &y_s[warp][lane >> 3][(lane & 7) * 4];
      if (t0 + Q <= T) {
#pragma unroll
        for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(yb + r * y_step4) = *reinterpret_cast<const float4*>(ys + 4 * r * SM_YS);
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (t0 + 4 * r + (lane >> 3) < T) *reinterpret_cast<float4*>(yb + r * y_step4) = *reinterpret_cast<const float4*>(ys + 4 * r * SM_YS);
      }
    }
    __syncwarp();                                                    // y_s is rewritten by the next chunk
  };

  // ---- software pipeline: iteration c computes chunk c, builds the operands of chunk c+1, convolves [B|C] of chunk c+2; raw rows of chunk c+3
  //      and the x tile of chunk c+1 are in flight.  One CTA barrier per chunk; u_s columns are private to their warp.
  raw_async(0); dt_fetch(0);
  cp_async_commit_wait_all();
  __syncthreads();
  conv_bc(0);
  dt_fetch(1);
  raw_async(1); x_async(0, 0);
  cp_async_commit_wait_all();
  __syncthreads();
  build_op(0);
  conv_bc(1);
  dt_fetch(2);
  prepass(0);
  raw_async(2);
  cp_async_commit_wait_all();
  __syncthreads();
  for (int c = 0; c < nch; ++c) {
    x_async(c + 1, (c + 1) & 1);
    raw_async(c + 3);
    compute(c);
    build_op(c + 1);
    conv_bc(c + 2);
    dt_fetch(c + 3);
    cp_async_commit_wait_all();
    __syncwarp();
    prepass(c + 1);
    __syncthreads();
  }

  if (p.final_state) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = pblk * PB + warp * 32 + mt * 16 + gid + ((j & 2) ? 8 : 0);
          const int n = 8 * nt + 2 * tig + (j & 1);
          p.final_state[(((size_t)b * p.H + h) * P + ch) * N + n] = S[mt][nt][j];
        }
  }
}

static bool ssd_mma_ok(const SsdParams& p) {
  if (p.N != 16 || p.P % 64 != 0) return false;
  if (p.ldbc % 4 != 0 || ((uintptr_t)p.Bm & 15) || ((uintptr_t)p.Cm & 15)) return false;
  if (p.ldy % 4 != 0 || ((uintptr_t)p.y & 15) || p.ldx % 4 != 0 || ((uintptr_t)p.x & 15)) return false;
  return true;
}

static int launch_ssd_mma(cudaStream_t st, const SsdParams& p, int64_t B) {
  const bool conv = p.fused && p.kconv > 0;
  static int pb64 = -1;                                              // EIGB200_SSD_PB=64: two-warp CTAs even when the head has 128 channels (tuning hook)
  if (pb64 < 0) { const char* e = getenv("EIGB200_SSD_PB"); pb64 = (e && atoi(e) == 64) ? 1 : 0; }
  if (p.P % 128 == 0 && !pb64) {
    dim3 grid(p.P / 128, p.H, (unsigned)B);
    if (conv) ssd_chunk_mma_kernel<128, true><<<grid, 128, 0, st>>>(p);
    else ssd_chunk_mma_kernel<128, false><<<grid, 128, 0, st>>>(p);
  } else {
    dim3 grid(p.P / 64, p.H, (unsigned)B);
    if (conv) ssd_chunk_mma_kernel<64, true><<<grid, 64, 0, st>>>(p);
    else ssd_chunk_mma_kernel<64, false><<<grid, 64, 0, st>>>(p);
  }
  EIGB_LAUNCH_CHECK("ssd_chunk_mma_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200
