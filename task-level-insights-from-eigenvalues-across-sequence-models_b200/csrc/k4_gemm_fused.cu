// k4_gemm_fused.cu -- the tail of a Mamba block as ONE tcgen05 kernel:   x_out = GLU(GELU(y W_out^T + b_out) W_glu^T + b_glu) + x_in   (+ extractor partials)
//
// Reference operators: MambaBlock.forward, models/mamba.py:333-337 (`self.activation(self.mamba(...))` = GELU(out_proj(y)), then `self.glu(x) + skip`),
// GLU, models/common.py:50-58; the extractor partials feed get_eig_mamba2 of the block OUTPUT (analysis/eval_eig.py:176-190, :512-520).
//
// Why: as two kernels (eigb200_linear[gelu] + eigb200_linear_glu_extract) the intermediate o = GELU(out_proj(y)) is written to HBM and read back
// (2 x 1.07 GB per layer at BASELINE C2, of 11.3 GB per layer in total), every tile pays two TMA pipelines, two converters and two epilogue store paths,
// and the GLU's 256 accumulator columns force an N-split pair of CTAs that both convert the same A tile.  Here o never leaves the SM -- it never even
// reaches shared memory: the GELU epilogue of GEMM 1 splits it into fp16 hi / lo and writes it with tcgen05.st into the TMEM A operand of GEMM 2.
//
// Shape (the C2 / MQAR family): d_model D = 128, d_inner K1 <= 128.  fp16-split operands only (kind::f16, see k4_gemm_tc.cu), which is what lets ONE CTA
// keep both weight matrices resident: W_out 128 x K1 (hi + lo, 64 KB) and all of W_glu, 256 x 128 (hi + lo, 128 KB).  Every CTA owns whole 128-row tiles:
// each y tile is read and converted once, each o element gets ONE GELU (the first version of this kernel split the GLU columns over a CTA pair and paid
// the out_proj GEMM, its conversion and the GELU twice: 1.03 ms per layer at C2, issue-bound at 68 %).
//
// Per CTA (704 threads = 16 epilogue warps, 4 converter warps, TMA warp, MMA warp), per 128-row tile i:
//   TMA     : y chunks (32 fp32 columns, 16 KB) into a 2-stage ring
//   convert : thread = row: S_a y -> fp16 hi / lo -> tcgen05.st into one of 4 TMEM operand stages                      (as in gemm_tc_ts_kernel<F16>)
//   MMA 1   : D1 (128 columns) = Y W_out^T, 6 kind::f16 MMAs per chunk, A from TMEM.  Issued for tile i+1 as soon as epi 1 of tile i has pulled D1 into
//             registers, i.e. it runs under the GELU math of tile i.
//   epi 1   : tcgen05.ld D1 -> (1 / S_a S_w1) acc + b_out -> S_a GELU -> fp16 hi / lo -> tcgen05.st into O (TMEM, 64 + 64 columns), one barrier per
//             32-column group so that MMA 2 starts on the first finished group
//   MMA 2   : two passes over the GLU columns, D2 (128 columns: per 32-column group 16 value + 16 gate columns) = O W_glu[pass]^T, A from TMEM;
//             pass b is issued when epi 2a has pulled D2 into registers and runs under its math
//   epi 2   : tcgen05.ld D2 -> value * sigmoid(gate) + residual -> 256-bit row-segment stores + extractor partials  (as the GLU epilogue of k4_gemm_tc.cu)
// TMEM (512 columns): D1 128 | D2 128 | O hi 64 | O lo 64 | 4 operand stages x 32.  Shared memory: 192 KB weights + 32 KB ring = 224 KB.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#ifndef EIGB200_TAIL_DEFAULT_T
#define EIGB200_TAIL_DEFAULT_T 0
#endif

namespace eigb200 {

constexpr int FG_BM = 128;
constexpr int FG_D = 128;                       // d_model: N of GEMM 1, K of GEMM 2, output columns
constexpr int FG_CHUNK_BYTES = FG_BM * 32 * 4;  // 16 KB: a raw fp32 chunk, and 128 rows x 128 bytes of an fp16 operand
constexpr int FG_EPI_WARPS = 16;
constexpr int FG_CONV_WARP0 = FG_EPI_WARPS;
constexpr int FG_TMA_WARP = FG_EPI_WARPS + 4;
constexpr int FG_MMA_WARP = FG_EPI_WARPS + 5;
constexpr int FG_THREADS = (FG_EPI_WARPS + 6) * 32;
constexpr int FG_NST = 2;                       // raw ring stages
constexpr int FG_AST = 2;                       // TMEM operand stages of GEMM 1
// TMEM: D1 128 | D2 64 (one pass of 32 output columns) | O double-buffered, per buffer hi 64 + lo 64 | 2 operand stages x 32
constexpr uint32_t FG_COL_D1 = 0, FG_COL_D2 = 128, FG_COL_O = 192, FG_COL_A = 448;
constexpr float FG_SA = 16.f;                   // activation pre-scale of both GEMMs (no LayerNorm in front of either)

struct FgParams {
  const float* bias1; const float* bias2;       // out_proj bias (D) | NULL, GLU bias (2 D) | NULL
  const float* osc1; const float* osc2;         // 1 / (S_a S_w) of the two prepared weight sets (device scalars)
  float* C; int64_t ldc; const float* R; int64_t ldr;
  const float* eig_w; float* eig_part;          // extractor partials (nullable)
  int64_t M, ntiles; int kch1, kchw1, zero; int* ovf_flag;
};

__global__ void __launch_bounds__(FG_THREADS, 1)
gemm_out_glu_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapW1h, const __grid_constant__ CUtensorMap tmapW1l,
                    const __grid_constant__ CUtensorMap tmapW2h, const __grid_constant__ CUtensorMap tmapW2l, const FgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w1hi = base, w1lo = w1hi + 2 * FG_CHUNK_BYTES;                     // [K chunk of 64][128 rows][128 B]
  const uint32_t w2hi = w1lo + 2 * FG_CHUNK_BYTES, w2lo = w2hi + 4 * FG_CHUNK_BYTES; // [K chunk of 64][256 rows][128 B]: rows 0-127 pass a, 128-255 pass b
  const uint32_t ring = w2lo + 4 * FG_CHUNK_BYTES;
  const uint32_t bars = ring + FG_NST * FG_CHUNK_BYTES;
  const uint32_t bar_w = bars;
  auto bar_full = [&](int s) { return bars + 8u * (1 + s); };                       // TMA landed a raw chunk
  auto bar_free = [&](int s) { return bars + 8u * (3 + s); };                       // converters have it in registers
  auto bar_afull = [&](int t) { return bars + 8u * (5 + t); };                      // TMEM operand stage written
  auto bar_aempty = [&](int t) { return bars + 8u * (9 + t); };                     // MMAs reading it retired
  const uint32_t bar_d1full = bars + 8u * 13, bar_d1empty = bars + 8u * 14;
  auto bar_ofull = [&](int b, int g) { return bars + 8u * (15 + 4 * b + g); };      // 32-column group g of O[b] written (4 warps)
  auto bar_oempty = [&](int b) { return bars + 8u * (23 + b); };                    // all passes of MMA 2 have read O[b]
  auto bar_d2full = [&](int h) { return bars + 8u * (25 + h); };                    // passes of parity h (read by epilogue warp set h)
  auto bar_d2empty = [&](int h) { return bars + 8u * (27 + h); };
  const uint32_t tmem_slot = bars + 8u * 29;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* b1s = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));   // [128] out_proj bias
  float* b2s = b1s + 128;                                                            // [256] GLU bias in accumulator-column order of the two passes (gate entries x -log2 e)
  // (barriers 256 B + biases 1536 B: the 2 KB behind the ring; the extractor gate weights are read through L1 instead)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kch1 = p.kch1;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < FG_NST; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_free(s), 128); }
    for (int t = 0; t < FG_AST; ++t) { mbar_init(bar_afull(t), 128); mbar_init(bar_aempty(t), 1); }
    mbar_init(bar_d1full, 1); mbar_init(bar_d1empty, FG_EPI_WARPS * 32);
    for (int b = 0; b < 2; ++b) {
      for (int g = 0; g < 4; ++g) mbar_init(bar_ofull(b, g), 128);
      mbar_init(bar_oempty(b), 1);
      mbar_init(bar_d2full(b), 1); mbar_init(bar_d2empty(b), FG_EPI_WARPS * 16);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 256) {
    const int c = threadIdx.x;                                                       // accumulator column c & 127 of pass c >> 7
    if (c < 128) b1s[c] = p.bias1 ? p.bias1[c] : 0.f;
    const int n = glu_weight_row(c & 127, c >> 7, 64, FG_D);                         // -> row of W_glu / entry of b_glu
    float bv = (p.bias2 && n >= 0) ? p.bias2[n] : 0.f;
    if (c & 16) bv *= -1.4426950408889634f;
    b2s[c] = bv;
  }
  if (warp == FG_MMA_WARP) tmem_alloc(tmem_slot, 512);
  if (warp == FG_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmapA); tma_prefetch_desc(&tmapW1h); tma_prefetch_desc(&tmapW1l); tma_prefetch_desc(&tmapW2h); tma_prefetch_desc(&tmapW2l);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == FG_TMA_WARP) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)(2 * p.kchw1 + 8) * FG_CHUNK_BYTES);
      for (int c = 0; c < p.kchw1; ++c) {
        tma_load_2d(&tmapW1h, bar_w, w1hi + c * FG_CHUNK_BYTES, c * 64, 0);
        tma_load_2d(&tmapW1l, bar_w, w1lo + c * FG_CHUNK_BYTES, c * 64, 0);
      }
      for (int c = 0; c < 2; ++c) {                                  // 256 rows per K chunk (box 64 x 256 = 32 KB)
        tma_load_2d(&tmapW2h, bar_w, w2hi + c * 2 * FG_CHUNK_BYTES, c * 64, 0);
        tma_load_2d(&tmapW2l, bar_w, w2lo + c * 2 * FG_CHUNK_BYTES, c * 64, 0);
      }
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int c = 0; c < kch1; ++c) {
          mbar_wait_one(bar_free(s), ph ^ 1);
          mbar_arrive_expect_tx(bar_full(s), FG_CHUNK_BYTES);
          tma_load_2d(&tmapA, bar_full(s), ring + s * FG_CHUNK_BYTES, c * 32, (int)(tile * FG_BM));
          if (++s == FG_NST) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= FG_CONV_WARP0 && warp < FG_CONV_WARP0 + 4) {
    // ===================================== converters: one row of y (TMEM lane) per thread ======================================
    const int r = threadIdx.x - FG_CONV_WARP0 * 32;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t row_off = (uint32_t)r * 128u;
    const int sw = r & 7;
    int s = 0; uint32_t ph = 0;
    int t = 0; uint32_t aph = 0;
    float amax = 0.f;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int c = 0; c < kch1; ++c) {
        mbar_wait(bar_full(s), ph);
        const float4* src = reinterpret_cast<const float4*>(smem_raw + (ring + s * FG_CHUNK_BYTES + row_off - smem_u32(smem_raw)));
        float a[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {                                // logical 16-byte slot q sits at physical slot q ^ (row & 7)
          const float4 v = src[q ^ sw];
          a[4 * q] = v.x * FG_SA; a[4 * q + 1] = v.y * FG_SA; a[4 * q + 2] = v.z * FG_SA; a[4 * q + 3] = v.w * FG_SA;
        }
        uint32_t hi2[16], lo2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi2[i] = pack_f16x2(a[2 * i], a[2 * i + 1]);
          amax = fmaxf(amax, fmaxf(fabsf(a[2 * i]), fabsf(a[2 * i + 1])));
        }
        {                                                            // the raw slot can be refilled: its values are in registers
          uint32_t dep = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) dep ^= hi2[2 * q];
          mbar_arrive_after(bar_free(s), dep, (uint32_t)p.zero);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) lo2[i] = pack_f16x2(a[2 * i] - f16_lo_to_f32(hi2[i]), a[2 * i + 1] - f16_hi_to_f32(hi2[i]));
        mbar_wait(bar_aempty(t), aph ^ 1);
        tc_fence_after();
        const uint32_t acol = tmem_base + lane_sel + FG_COL_A + (uint32_t)t * 32u;
        tmem_st_32x16(acol, hi2);
        tmem_st_32x16(acol + 16u, lo2);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_afull(t));
        if (++s == FG_NST) { s = 0; ph ^= 1; }
        if (++t == FG_AST) { t = 0; aph ^= 1; }
      }
    }
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  } else if (warp == FG_MMA_WARP) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(FG_BM, 128);
      mbar_wait_one(bar_w, 0);
      int t = 0; uint32_t aph = 0;
      const uint32_t d1 = tmem_base + FG_COL_D1, d2 = tmem_base + FG_COL_D2;
      auto mma1 = [&](int it) {                                      // D1 = Y(it) W_out^T
        mbar_wait_one(bar_d1empty, (uint32_t)(it & 1) ^ 1u);        // epi 1 of the previous tile has D1 in registers
        tc_fence_after();
        for (int c = 0; c < kch1; ++c) {
          mbar_wait_one(bar_afull(t), aph);
          tc_fence_after();
          const uint32_t a_hi = tmem_base + FG_COL_A + (uint32_t)t * 32u, a_lo = a_hi + 16u;
          const uint32_t wofs = (uint32_t)(c >> 1) * FG_CHUNK_BYTES;
          const uint64_t kofs = (c & 1) ? 4u : 0u;                   // second half of the 128-byte row: +64 bytes
          const uint64_t dbh = umma_desc_k_sw128(w1hi + wofs) + kofs, dbl = umma_desc_k_sw128(w1lo + wofs) + kofs;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_f16_ts(d1, a_hi + 8u * k, dbh + 2u * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_f16_ts(d1, a_hi + 8u * k, dbl + 2u * k, idesc, 1u);
            umma_f16_ts(d1, a_lo + 8u * k, dbh + 2u * k, idesc, 1u);
          }
          umma_commit(bar_aempty(t));
          if (c == kch1 - 1) umma_commit(bar_d1full);
          if (++t == FG_AST) { t = 0; aph ^= 1; }
        }
      };
      const uint32_t idesc2 = umma_idesc_f16(FG_BM, 64);
      auto mma2 = [&](int it, int q) {                               // D2 = O(it) W_glu[rows 64 q .. 64 q + 63]^T: output columns [32 q, 32 q + 32)
        const int ob = it & 1;
        if (4 * it + q > 0) {                                        // the epilogue warps that read the PREVIOUS pass have it in registers
          const int pq = 4 * it + q - 1;
          mbar_wait_one(bar_d2empty(pq & 1), (uint32_t)((pq >> 1) & 1));
        }
        tc_fence_after();
        const uint32_t o_hi = tmem_base + FG_COL_O + (uint32_t)ob * 128u, o_lo = o_hi + 64u;
        for (int g = 0; g < 4; ++g) {
          if (q == 0) { mbar_wait_one(bar_ofull(ob, g), (uint32_t)((it >> 1) & 1)); tc_fence_after(); }
          const uint32_t cofs = (uint32_t)(g >> 1) * 2u * FG_CHUNK_BYTES + (uint32_t)q * (FG_CHUNK_BYTES / 2);   // 64 rows x 128 B per pass
          const uint64_t kofs = (g & 1) ? 4u : 0u;
          const uint64_t dbh = umma_desc_k_sw128(w2hi + cofs) + kofs, dbl = umma_desc_k_sw128(w2lo + cofs) + kofs;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_f16_ts(d2, o_hi + 16u * g + 8u * k, dbh + 2u * k, idesc2, (g > 0 || k > 0) ? 1u : 0u);
            umma_f16_ts(d2, o_hi + 16u * g + 8u * k, dbl + 2u * k, idesc2, 1u);
            umma_f16_ts(d2, o_lo + 16u * g + 8u * k, dbh + 2u * k, idesc2, 1u);
          }
        }
        umma_commit(bar_d2full(q & 1));
        if (q == 3) umma_commit(bar_oempty(ob));
      };
      // Issue order = the order in which the operands become available.  While the epilogue warps do the GELU math of tile it (epi 1), the tensor pipe
      // runs MMA 1 of tile it + 1 (D1 is free as soon as epi 1 has pulled it into registers); passes 1-3 of tile it - 1 interleave with the epi 2 passes
      // that drain D2 one after the other; pass 0 of tile it follows once O(it) is written and is consumed first thing in the next epi 2.
      int it = 0;
      if ((int64_t)blockIdx.x < p.ntiles) mma1(0);
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        if (tile + (int64_t)gridDim.x < p.ntiles) mma1(it + 1);
        if (it > 0) { mma2(it - 1, 1); mma2(it - 1, 2); mma2(it - 1, 3); }
        mma2(it, 0);
      }
      if (it > 0) { mma2(it - 1, 1); mma2(it - 1, 2); mma2(it - 1, 3); }
    }
    __syncwarp();
  } else if (warp < FG_EPI_WARPS) {
    // ===================================== epilogue warps ======================================
    const int quarter = warp & 3, g = warp >> 2;                     // TMEM lane quarter (32 rows); 32-column group of D1 / O and of D2
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;                               // row of the tile
    const float osc1 = __ldg(p.osc1), osc2 = __ldg(p.osc2);
    const float osc2_gate = -1.4426950408889634f * osc2;
    float amax = 0.f;
    const int wset = warp >> 3, grp2 = (warp >> 2) & 1;              // epi 2: warp set (passes of parity wset), 32-column accumulator group of the pass
    auto epi2 = [&](int it, int q, int64_t tile) {                   // pass q of tile it: output columns [32 q, 32 q + 32); this warp: 16 of them
      const int64_t own_row = tile * FG_BM + r;
      const int oc = 32 * q + 16 * grp2;
      float rr[16];
      if (own_row < p.M) {                                           // residual prefetch: its DRAM latency hides behind the accumulator wait
        const float* rptr = p.R + own_row * p.ldr + oc;
        ldg_stream_v8(rptr, rr); ldg_stream_v8(rptr + 8, rr + 8);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) rr[i] = 0.f;
      }
      mbar_wait(bar_d2full(wset), (uint32_t)((2 * it + (q >> 1)) & 1));
      tc_fence_after();
      float a[32];
      tmem_ld_32x32(tmem_base + FG_COL_D2 + lane_sel + 32u * grp2, a);
      tc_fence_before();
      mbar_arrive(bar_d2empty(wset));
      const float* b2 = b2s + 64 * q + 32 * grp2;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {                                 // sigmoid(g + b) = 1 / (1 + 2^((g + b) * -log2 e)), operand scales folded into the FMAs
        float e2;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(a[16 + i], osc2_gate, b2[16 + i])));
        v[i] = fmaf(a[i], osc2, b2[i]) * fast_rcp_f(1.f + e2) + rr[i];
      }
      if (p.eig_part) {                                              // extractor partials of the finished 16-column row segment (see tc_epilogue)
        float we[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.eig_w + oc) + q4);
          we[4 * q4] = w4.x; we[4 * q4 + 1] = w4.y; we[4 * q4 + 2] = w4.z; we[4 * q4 + 3] = w4.w;
        }
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
      if (own_row < p.M) {
        float* cptr = p.C + own_row * p.ldc + oc;
        stg_v8(cptr, v); stg_v8(cptr + 8, v + 8);
      }
    };
    int it = 0;
    int64_t prev_tile = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      // ---- epi 1: D1 -> S_a GELU -> fp16 hi / lo -> O[it & 1] (TMEM A operand of GEMM 2) ----
      const int ob = it & 1;
      mbar_wait(bar_d1full, (uint32_t)(it & 1));
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(tmem_base + FG_COL_D1 + lane_sel + 32u * g, v);
      tc_fence_before();
      mbar_arrive(bar_d1empty);                                      // MMA 1 of the next tile may overwrite D1 from here on
      uint32_t hi2[16], lo2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float z0 = fmaf(v[2 * i], osc1, b1s[32 * g + 2 * i]), z1 = fmaf(v[2 * i + 1], osc1, b1s[32 * g + 2 * i + 1]);
        const float o0 = gelu_fast_scaled_f(z0, z0 * FG_SA, FG_SA);      // S_a GELU(z): the operand scale of GEMM 2 rides in the constants
        const float o1 = gelu_fast_scaled_f(z1, z1 * FG_SA, FG_SA);
        amax = fmaxf(amax, fmaxf(fabsf(o0), fabsf(o1)));
        hi2[i] = pack_f16x2(o0, o1);
        lo2[i] = pack_f16x2(o0 - f16_lo_to_f32(hi2[i]), o1 - f16_hi_to_f32(hi2[i]));
      }
      mbar_wait(bar_oempty(ob), (uint32_t)((it >> 1) & 1) ^ 1u);    // MMA 2 of tile it - 2 has read this O buffer
      tc_fence_after();
      const uint32_t o_hi = tmem_base + lane_sel + FG_COL_O + (uint32_t)ob * 128u;
      tmem_st_32x16(o_hi + 16u * g, hi2);
      tmem_st_32x16(o_hi + 64u + 16u * g, lo2);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_ofull(ob, g));
      // ---- epi 2 of the PREVIOUS tile (its MMA 2 passes ran under the GELU math above): this warp set takes the passes of its parity ----
      if (it > 0) { epi2(it - 1, wset, prev_tile); epi2(it - 1, 2 + wset, prev_tile); }
      prev_tile = tile;
    }
    if (it > 0) { epi2(it - 1, wset, prev_tile); epi2(it - 1, 2 + wset, prev_tile); }
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FG_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

bool out_glu_fused_supported(int D, int K1) { return D == FG_D && K1 % 32 == 0 && K1 >= 32 && K1 <= 128; }

int launch_out_glu_fused(cudaStream_t st, const float* A, int64_t lda, const void* ws1, const float* bias1, const void* ws2, const float* bias2,
                         float* C, int64_t ldc, const float* R, int64_t ldr, int64_t M, int D, int K1, const float* eig_w, float* eig_part) {
  if (!out_glu_fused_supported(D, K1)) { set_error("out_glu_fused: needs d_model = 128 and d_inner a multiple of 32 <= 128 (D=%d K1=%d)", D, K1); return EIGB200_EUNSUPPORTED; }
  if (lda % 4 != 0 || ((uintptr_t)A & 15) || ldc % 8 != 0 || ((uintptr_t)C & 31) || ldr % 8 != 0 || ((uintptr_t)R & 31) || M >= (1LL << 31)) {
    set_error("out_glu_fused: y rows must be 16-byte aligned, residual and output rows 32-byte aligned"); return EIGB200_EUNSUPPORTED;
  }
  TcPrepared p1, p2;
  if (!tc_prepared_layout_f16(D, K1, EIGB200_EPI_GELU, ws1, &p1) || !tc_prepared_layout_f16(2 * D, D, EIGB200_EPI_GLU_RESIDUAL, ws2, &p2) ||
      p1.nsplit != 1 || p1.bn != 128 || p2.nsplit != 2 || p2.bn != 128 || p2.bg != 64 || p2.kp64 != 128) {
    set_error("out_glu_fused: unexpected operand plan for D=%d K1=%d", D, K1); return EIGB200_EUNSUPPORTED;
  }
  CUtensorMap tA, tW1h, tW1l, tW2h, tW2l;
  int rc;
  if ((rc = tc_make_tmap_f32(&tA, A, (uint64_t)M, (uint64_t)K1, (uint64_t)lda, FG_BM))) return rc;
  if ((rc = tc_make_tmap_f16(&tW1h, p1.w_hi, (uint64_t)p1.wrows, (uint64_t)p1.kp64, 128))) return rc;
  if ((rc = tc_make_tmap_f16(&tW1l, p1.w_lo, (uint64_t)p1.wrows, (uint64_t)p1.kp64, 128))) return rc;
  if ((rc = tc_make_tmap_f16(&tW2h, p2.w_hi, (uint64_t)p2.wrows, (uint64_t)p2.kp64, 256))) return rc;   // all 256 rows of a K chunk in one box
  if ((rc = tc_make_tmap_f16(&tW2l, p2.w_lo, (uint64_t)p2.wrows, (uint64_t)p2.kp64, 256))) return rc;
  FgParams p{};
  p.bias1 = bias1; p.bias2 = bias2; p.osc1 = p1.scal; p.osc2 = p2.scal;
  p.C = C; p.ldc = ldc; p.R = R; p.ldr = ldr; p.eig_w = eig_w; p.eig_part = eig_part;
  p.M = M; p.ntiles = (M + FG_BM - 1) / FG_BM;
  p.kch1 = K1 / 32; p.kchw1 = p1.kch_w; p.zero = 0;
  p.ovf_flag = tc_overflow_flag();
  if (!p.ovf_flag) { set_error("out_glu_fused: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  int64_t grid = p.ntiles < (int64_t)num_sms() ? p.ntiles : (int64_t)num_sms();
  const size_t smem = (size_t)(4 + 8 + FG_NST) * FG_CHUNK_BYTES + 1024 /*alignment*/ + 2048 /*barriers, biases*/;
  EIGB_CUDA(cudaFuncSetAttribute(gemm_out_glu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_out_glu_kernel<<<(unsigned)grid, FG_THREADS, smem, st>>>(tA, tW1h, tW1l, tW2h, tW2l, p);
  EIGB_LAUNCH_CHECK("gemm_out_glu_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_out_glu_fused_supported(int D, int K1) { return out_glu_fused_supported(D, K1) ? 1 : 0; }

extern "C" int eigb200_out_glu_fused(void* stream, const float* d_y, int64_t ldy, const void* d_ws_out, const float* d_bias_out,
                                     const void* d_ws_glu, const float* d_bias_glu, float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                                     int64_t M, int D, int K1, const float* d_W_gate, float* d_partials) {
  EIGB_CHECK_ARG(d_y && d_ws_out && d_ws_glu && d_C && d_R, "out_glu_fused: null pointer");
  EIGB_CHECK_ARG(M > 0 && D > 0 && K1 > 0, "out_glu_fused: bad shape M=%lld D=%d K1=%d", (long long)M, D, K1);
  EIGB_CHECK_ARG(ldy >= K1 && ldc >= D && ldr >= D, "out_glu_fused: row stride smaller than the row");
  EIGB_CHECK_ARG((d_W_gate == nullptr) == (d_partials == nullptr), "out_glu_fused: the extractor needs both the gate weights and the partials buffer");
  EIGB_CHECK_ARG(tc_default_kind() == 1, "out_glu_fused: the prepared operands must be the fp16 split (EIGB200_GEMM_PRECISION=f16x3)");
  {
    // EIGB200_TAIL_FORM: "t" = the transposed form (k8_tail_fused_t.cu) where its shape conditions hold, "n" = this file's kernel.  Read per call (cheap next
    // to a launch) so that the tests can switch forms.
    const char* e = getenv("EIGB200_TAIL_FORM");
    const bool want_t = e ? (e[0] == 't') : (EIGB200_TAIL_DEFAULT_T != 0);
    if (want_t && out_glu_fused_t_supported(D, K1))
      return launch_out_glu_fused_t((cudaStream_t)stream, d_y, ldy, d_ws_out, d_bias_out, d_ws_glu, d_bias_glu, d_C, ldc, d_R, ldr, M, D, K1, d_W_gate, d_partials);
  }
  return launch_out_glu_fused((cudaStream_t)stream, d_y, ldy, d_ws_out, d_bias_out, d_ws_glu, d_bias_glu, d_C, ldc, d_R, ldr, M, D, K1, d_W_gate, d_partials);
}
