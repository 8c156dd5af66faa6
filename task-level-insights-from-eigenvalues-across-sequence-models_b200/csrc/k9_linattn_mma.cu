// k9_linattn_mma.cu -- causal linear / normalised attention layer forward in CHUNKED form on the warp-level tensor-core path
// (mma.sync.m16n8k8 tf32, error-compensated 3xTF32), with the depthwise causal conv + SiLU of q / k / v optionally fused into the tile loader.
//
// Reference operators: SelfLinAttention.forward, models/attention.py:63-83 (kv cumsum materialised as (B,T,H,d,dv)); SelfNormAttention.forward,
// models/norm_attention.py:61-89; the conv in front of them MHA.forward models/attention.py:153-156 / MHNA.forward norm_attention.py:236-239.
//
//   S_t = S_{t-1} + phi(k_t) (kscale v_t)^T,      out_t = scale_t * phi(q_t)^T S_t,      scale_t = 1 / (phi(q_t) . sum_{s<=t} phi(k_s))  or  gate[b,t,h]
//
// Per chunk of 64 tokens of one (b, h), with S0 the state at the chunk start:
//   P = mask_{s<=t}(Q K^T)      O = P V + Q S0      S0 += K^T V           four 64 x 64 x 64 products instead of 64 x 2 x 64 x 64 dependent FMAs.
// The recurrent column-owner kernel (k1_attn.cu) runs at 22 % of the FMA-pipe peak (8.4 ms per C5 layer: 1024 sequences x 1024 tokens x 8 heads); this
// form is bound by the legacy tensor path (277 TFLOP/s tf32 measured, tools/mma_rate.cu; 3 MMAs per product term) at about 2.4 ms.
//
// CTA = 4 warps = one (b, h), chunks in sequence, 3 CTAs per SM (one CTA's loads run under the others' MMAs).  Warp role r owns
//   * rows t = 16 r .. 16 r + 15 of P and O (P's accumulators ARE the A operand of P V: the accumulator's column pair (2j, 2j + 1) is read as K indices
//     (j, j + 4), and the rows of V are fetched in that order),
//   * rows dd = 16 r .. of the state S, which lives in accumulator registers for the whole sequence and is published to shared memory once per chunk
//     as the B operand of Q S0.
// Shared tiles are fp32 with row stride 68 (S: 72): every fragment load below is bank-conflict free, including the token-contracted ones (K^T V, P V), which
// read tokens (2j, 2j + 1) per K index pair.
// The phases are rolled loops (key-block pairs, head-dim steps, token blocks) with fully unrolled bodies of independent MMA chains: the first version unrolled
// everything (76 KB of straight-line code, instruction fetch was the top stall) and branched per tile (serialised chains).
#include "linattn.cuh"

namespace eigb200 {

constexpr int LM_C = 64;                 // tokens per chunk
constexpr int LM_LD = 68;                // row stride of the q / k / v tiles (floats): 4 mod 32
constexpr int LM_LDS = 72;               // row stride of the published state: 8 mod 32
constexpr int LM_THREADS = 128;
constexpr int LM_SMEM_FLOATS = 3 * LM_C * LM_LD + 64 * LM_LDS + 2 * 64;

__device__ __forceinline__ void lm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tf32 split.  The tensor core reads only the upper 19 bits of an operand register and truncates lo's own low bits the same way.
//   lm_split    hi rounded to nearest (integer add of half a tf32 ulp, then the mask): lo = v - hi is exact and SIGNED, so the hardware's truncation of lo
//               is unbiased.  3 instructions.  Used for the all-positive operands (phi(q), phi(k), P).
//   lm_split_tr hi = the fp32 word itself (no instruction), lo = v - trunc19(v): 2 instructions; on an all-positive operand the truncated lo is a one-sided
//               error, on the signed operands (v, the state) it has no preferred sign -- used there.
// The split is 40 % of this kernel's non-MMA instructions, and those ADD to its HMMA time rather than hide under it (ablations in profiles/).  Measured at the
// C5 layer shape / as activation deviation after 12 norm-attention blocks (tests/test_parity_fullshape_gpu.py; the fp32 recurrent kernel: 6.70 ms, 4.7e-6):
// nearest everywhere (-DLM_SPLIT_RN) 4.95 ms, 5.5e-6; this mix 4.79 ms, 6.3e-6; truncation everywhere (-DLM_SPLIT_TR) 4.57 ms, 7.5e-6.
__device__ __forceinline__ void lm_split(float v, uint32_t& hi, uint32_t& lo) {
#ifdef LM_SPLIT_TR
  hi = __float_as_uint(v);
  lo = __float_as_uint(v - __uint_as_float(hi & 0xffffe000u));
#else
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
#endif
}
__device__ __forceinline__ void lm_split_tr(float v, uint32_t& hi, uint32_t& lo) {
#ifdef LM_SPLIT_RN
  lm_split(v, hi, lo);
#else
  hi = __float_as_uint(v);
  lo = __float_as_uint(v - __uint_as_float(hi & 0xffffe000u));
#endif
}
// D[i] += A B_i for four n-tiles, A = ah + al, B_i = bh + bl: small terms first, the three dependent MMAs of a tile four instructions apart
__device__ __forceinline__ void lm_mma3x4(float (&d0)[4], float (&d1)[4], float (&d2)[4], float (&d3)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[8], const uint32_t (&bl)[8]) {
  lm_mma(d0, al, bh[0], bh[1]); lm_mma(d1, al, bh[2], bh[3]); lm_mma(d2, al, bh[4], bh[5]); lm_mma(d3, al, bh[6], bh[7]);
  lm_mma(d0, ah, bl[0], bl[1]); lm_mma(d1, ah, bl[2], bl[3]); lm_mma(d2, ah, bl[4], bl[5]); lm_mma(d3, ah, bl[6], bl[7]);
  lm_mma(d0, ah, bh[0], bh[1]); lm_mma(d1, ah, bh[2], bh[3]); lm_mma(d2, ah, bh[4], bh[5]); lm_mma(d3, ah, bh[6], bh[7]);
}
// A fragment (hi / lo) from four fp32 values
__device__ __forceinline__ void lm_afrag(float a0, float a1, float a2, float a3, uint32_t (&ah)[4], uint32_t (&al)[4]) {
  lm_split(a0, ah[0], al[0]); lm_split(a1, ah[1], al[1]); lm_split(a2, ah[2], al[2]); lm_split(a3, ah[3], al[3]);
}
// B fragments of four n-tiles of a SIGNED operand (v, the state): element pair (p0[8 i], p1[8 i]) for tile i
__device__ __forceinline__ void lm_bfrag4(const float* __restrict__ p0, const float* __restrict__ p1, uint32_t (&bh)[8], uint32_t (&bl)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { lm_split_tr(p0[8 * i], bh[2 * i], bl[2 * i]); lm_split_tr(p1[8 * i], bh[2 * i + 1], bl[2 * i + 1]); }
}
// phi = elu + 1 = x + 1 (x > 0), e^x (x <= 0): one expf instead of expm1f + 1 (each within an ulp of the exact value)
// (both sides are computed and selected: a per-element branch diverges on half the lanes)
__device__ __forceinline__ float lm_phi(float x) { const float e = expf(x), l = x + 1.f; return x > 0.f ? l : e; }   // (e may be +inf for large x: not selected)

// SiLU with ex2.approx / rcp.approx (as the Mamba conv, k2_ssd_scan.cu): ~3 ulp, no division slow path -- 96 activations per thread and chunk
__device__ __forceinline__ float lm_silu(float z) { return z * sigmoid_fast_f(z); }

// One 64 x 64 tile of q, k or v: global -> registers (lm_issue) ... -> (conv + SiLU) -> transform -> shared (lm_finish).  Thread = 4 columns (c4) x 8 rows
// (seg); rows >= tc are written as zeros (AFTER phi: phi(0) = 1 would otherwise enter the state).  is_v: multiply by kscale, else phi = elu + 1 (when
// phi_elu).  The two halves are separate so that the q / k rows of the NEXT chunk can be in flight under the state update of the current one.
struct LmRaw { float4 cur[8]; float4 hist[3]; };

template <bool CONV>
__device__ __forceinline__ void lm_issue(const LinAttnParams& p, const float* __restrict__ src, int64_t row0_global, int64_t seq_row0, int tc,
                                         int conv_ch, int tid, LmRaw& r) {
  const int c4 = tid & 15, seg = tid >> 4;
  const float* g = src + (row0_global + 8 * seg) * p.ld + 4 * c4;
  // branch-free loads: a row beyond the sequence end is clamped to the last valid row and zeroed in lm_finish (per-row guards cost a branch region per load)
  const int rlast = max(tc - 1 - 8 * seg, -8 * seg);                // last valid row of this segment, relative to its first row (may be < 0; >= row 0 of the chunk)
  if (tc == LM_C) {                                                 // full chunk (all but the last of a sequence): constant-stride addresses, no clamps
#pragma unroll
    for (int i = 0; i < 8; ++i) r.cur[i] = __ldg(reinterpret_cast<const float4*>(g + (int64_t)i * p.ld));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) r.cur[i] = __ldg(reinterpret_cast<const float4*>(g + (int64_t)min(i, rlast) * p.ld));
  }
  if (CONV && conv_ch >= 0) {
    const int64_t rfirst = -(seq_row0 + 8 * seg);                   // first row of the sequence relative to this segment's first row (<= 0)
#pragma unroll
    for (int i = 0; i < 3; ++i) {                                   // raw rows t - 3 + i of the sequence; zero padding before its start
      const int64_t rr = max((int64_t)min(i - 3, rlast), rfirst);
      const float4 hv = __ldg(reinterpret_cast<const float4*>(g + rr * p.ld));
      r.hist[i] = (i - 3 >= rfirst) ? hv : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <bool CONV>
__device__ __forceinline__ void lm_finish(const LinAttnParams& p, LmRaw& r, int tc, int conv_ch, bool is_v, float* __restrict__ dst, int tid) {
  const int c4 = tid & 15, seg = tid >> 4;
  const int rlast = tc - 1 - 8 * seg;
  if (CONV && conv_ch >= 0) {
    const int ch = conv_ch + 4 * c4;
    float w[4][4], bb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bb[e] = __ldg(p.conv_b + ch + e);
#pragma unroll
      for (int j = 0; j < 4; ++j) w[e][j] = (j >= 4 - p.kconv) ? __ldg(p.conv_w + (size_t)(ch + e) * p.kconv + j - (4 - p.kconv)) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x3 = r.cur[i];
      float4 o;                                                     // FMA order of conv_silu_kernel (oldest tap first, bias as the seed)
      o.x = lm_silu(fmaf(w[0][3], x3.x, fmaf(w[0][2], r.hist[2].x, fmaf(w[0][1], r.hist[1].x, fmaf(w[0][0], r.hist[0].x, bb[0])))));
      o.y = lm_silu(fmaf(w[1][3], x3.y, fmaf(w[1][2], r.hist[2].y, fmaf(w[1][1], r.hist[1].y, fmaf(w[1][0], r.hist[0].y, bb[1])))));
      o.z = lm_silu(fmaf(w[2][3], x3.z, fmaf(w[2][2], r.hist[2].z, fmaf(w[2][1], r.hist[1].z, fmaf(w[2][0], r.hist[0].z, bb[2])))));
      o.w = lm_silu(fmaf(w[3][3], x3.w, fmaf(w[3][2], r.hist[2].w, fmaf(w[3][1], r.hist[1].w, fmaf(w[3][0], r.hist[0].w, bb[3])))));
      r.hist[0] = r.hist[1]; r.hist[1] = r.hist[2]; r.hist[2] = x3;
      r.cur[i] = o;
    }
  }
  const bool phi = !is_v && p.phi_elu;
  const float mul = is_v ? p.kscale : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 o = r.cur[i];
    if (phi) { o.x = lm_phi(o.x); o.y = lm_phi(o.y); o.z = lm_phi(o.z); o.w = lm_phi(o.w); }
    o.x *= mul; o.y *= mul; o.z *= mul; o.w *= mul;
    if (i > rlast) o = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(dst + (8 * seg + i) * LM_LD + 4 * c4) = o;
  }
}

template <bool CONV>
__global__ void __launch_bounds__(LM_THREADS, 3) linattn_mma_kernel(const LinAttnParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                                                   // tiles: [Q | K | V], LM_C x LM_LD each
  float* Ks = Qs + LM_C * LM_LD;
  float* Vs = Ks + LM_C * LM_LD;
  float* Ss = Vs + LM_C * LM_LD;                                    // [dd][j], stride 72: the state at the chunk start
  float* ksum = Ss + 64 * LM_LDS;                                   // [2][64] running sum of phi(k), ping-pong over chunks (normalise)
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  // role rotation: the later row blocks do more causal work; hardware warp w sits on scheduler w % 4 in every CTA
  const int role = ((tid >> 5) + h + b) & 3;
  const int64_t rowbase = (int64_t)b * p.T;

  for (int i = tid; i < 64 * LM_LDS; i += LM_THREADS) Ss[i] = 0.f;
  ksum[tid] = 0.f;

  const float* qrow0 = Qs + (16 * role + g) * LM_LD;
  const float* qrow1 = qrow0 + 8 * LM_LD;
  float* srow0 = Ss + (16 * role + g) * LM_LDS + 2 * t4;            // this thread's state elements: rows dd = 16 role + g (+ 8), columns 8 nt + 2 t4 (+ 1)
  float* srow1 = srow0 + 8 * LM_LDS;
  const int tl0 = 16 * role + g, tl1 = tl0 + 8;                     // this thread's two token rows of the chunk
  float S[8][4];                                                    // live from the state update of a chunk to its write-back after the barrier
  int cbuf = 0;
  const float* qbase = p.q + (size_t)h * 64;
  const float* kbase = p.k + (size_t)h * 64;
  const float* vbase = p.v + (size_t)h * 64;
  const int chq = p.conv_ch_q >= 0 ? p.conv_ch_q + h * 64 : -1, chk = p.conv_ch_k >= 0 ? p.conv_ch_k + h * 64 : -1,
            chv = p.conv_ch_v >= 0 ? p.conv_ch_v + h * 64 : -1;

  for (int64_t t0 = 0; t0 < p.T; t0 += LM_C, cbuf ^= 1) {
    const int tc = (int)min((int64_t)LM_C, p.T - t0);
    // ---- state write-back, then stage the chunk, tile by tile: its rows were prefetched to L2 under the previous chunk's state update.  (Measured: rows
    // prefetched to REGISTERS are spilled by ptxas, which waits for them; all three tiles requested at once cost the conv variant 0.8 ms in spills.) ----
    if (t0 > 0) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<float2*>(srow0 + 8 * nt) = make_float2(S[nt][0], S[nt][1]);
        *reinterpret_cast<float2*>(srow1 + 8 * nt) = make_float2(S[nt][2], S[nt][3]);
      }
    }
#pragma unroll 1
    for (int m = 0; m < 3; ++m) {                                   // one copy of the loader code for the three tiles
      LmRaw r;
      const int ch = m == 0 ? chq : (m == 1 ? chk : chv);
      lm_issue<CONV>(p, m == 0 ? qbase : (m == 1 ? kbase : vbase), rowbase + t0, t0, tc, ch, tid, r);
      lm_finish<CONV>(p, r, tc, ch, m == 2, Qs + m * (LM_C * LM_LD), tid);
    }
    __syncthreads();

    float O[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { O[nt][0] = O[nt][1] = O[nt][2] = O[nt][3] = 0.f; }
    float rs0 = 0.f, rs1 = 0.f;                                     // row sums of the masked P (normalise)
    {
      // ---- Q fragments of this role's 16 rows, resident for the key-block loop ----
      uint32_t qh[8][4], ql[8][4];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) lm_afrag(qrow0[8 * kk + t4], qrow1[8 * kk + t4], qrow0[8 * kk + t4 + 4], qrow1[8 * kk + t4 + 4], qh[kk], ql[kk]);
      // ---- per pair of key blocks (16 keys): P = mask(Q K^T) (two accumulator tiles), O += P V.  Exactly the causal blocks: pairs 0 .. role ----
#ifdef LM_ABL_P
      if (t0 < 0)
#endif
#pragma unroll 1
      for (int kp = 0; kp <= role; ++kp) {
        float Pa[2][4], Pb[2][4], Pc[2][4];                         // the lo x hi, hi x lo, hi x hi chains of the two tiles (independent accumulators)
#pragma unroll
        for (int i = 0; i < 2; ++i) { Pa[i][0] = Pa[i][1] = Pa[i][2] = Pa[i][3] = 0.f; Pb[i][0] = Pb[i][1] = Pb[i][2] = Pb[i][3] = 0.f; Pc[i][0] = Pc[i][1] = Pc[i][2] = Pc[i][3] = 0.f; }
        const float* kp0 = Ks + (16 * kp + g) * LM_LD + t4;         // key rows 16 kp + g (tile 0) and + 8 (tile 1)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t bh0, bl0, bh1, bl1;
            lm_split(kp0[8 * i * LM_LD + 8 * kk], bh0, bl0); lm_split(kp0[8 * i * LM_LD + 8 * kk + 4], bh1, bl1);
            lm_mma(Pa[i], ql[kk], bh0, bh1); lm_mma(Pb[i], qh[kk], bl0, bl1); lm_mma(Pc[i], qh[kk], bh0, bh1);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float P[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) P[e] = (Pa[i][e] + Pb[i][e]) + Pc[i][e];
          const int s0 = 16 * kp + 8 * i + 2 * t4;                  // causal mask: keep s <= t (only the pair kp == role straddles the diagonal)
          if (s0 > tl0) P[0] = 0.f;
          if (s0 + 1 > tl0) P[1] = 0.f;
          if (s0 > tl1) P[2] = 0.f;
          if (s0 + 1 > tl1) P[3] = 0.f;
          rs0 += P[0] + P[1]; rs1 += P[2] + P[3];
          uint32_t ph[4], pl[4];
          lm_afrag(P[0], P[2], P[1], P[3], ph, pl);                 // accumulator columns (2 t4, 2 t4 + 1) = K indices (t4, t4 + 4)
          const float* v0 = Vs + (16 * kp + 8 * i + 2 * t4) * LM_LD + g;   // ... so the V rows are tokens 2 t4 and 2 t4 + 1 of the block
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t bh[8], bl[8];
            lm_bfrag4(v0 + 32 * hf, v0 + LM_LD + 32 * hf, bh, bl);
            lm_mma3x4(O[4 * hf], O[4 * hf + 1], O[4 * hf + 2], O[4 * hf + 3], ph, pl, bh, bl);
          }
        }
      }
    }
    float den0 = 1.f, den1 = 1.f;
    if (p.normalise) {                                              // den_t = q_t . ksum_t = rowsum of the masked P + q_t . ksum at the chunk start
      float r0 = rs0, r1 = rs1;
      const float* ks_cur = ksum + 64 * cbuf;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float k0 = ks_cur[8 * kk + t4], k1 = ks_cur[8 * kk + t4 + 4];
        r0 = fmaf(qrow0[8 * kk + t4], k0, r0); r0 = fmaf(qrow0[8 * kk + t4 + 4], k1, r0);
        r1 = fmaf(qrow1[8 * kk + t4], k0, r1); r1 = fmaf(qrow1[8 * kk + t4 + 4], k1, r1);
      }
      r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
      den0 = r0; den1 = r1;
      if (tid < 64) {                                               // ksum for the next chunk (the other buffer: nobody reads it in this phase)
        float a = ks_cur[tid];
        for (int r = 0; r < LM_C; ++r) a += Ks[r * LM_LD + tid];
        ksum[64 * (cbuf ^ 1) + tid] = a;
      }
    }
    // ---- O += Q S0 (the state at the chunk start, from shared memory) ----
#ifdef LM_ABL_QS
    if (t0 < 0) {
#else
    if (t0 > 0) {
#endif
#pragma unroll 2
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t ah[4], al[4];
        lm_afrag(qrow0[8 * kk + t4], qrow1[8 * kk + t4], qrow0[8 * kk + t4 + 4], qrow1[8 * kk + t4 + 4], ah, al);
        const float* s0 = Ss + (8 * kk + t4) * LM_LDS + g;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t bh[8], bl[8];
          lm_bfrag4(s0 + 32 * hf, s0 + 4 * LM_LDS + 32 * hf, bh, bl);
          lm_mma3x4(O[4 * hf], O[4 * hf + 1], O[4 * hf + 2], O[4 * hf + 3], ah, al, bh, bl);
        }
      }
    }
    // ---- scale and store ----
    {
      const int64_t r0 = rowbase + t0 + tl0, r1 = r0 + 8;
      float sc0 = 1.f, sc1 = 1.f;
      if (p.normalise) { sc0 = 1.f / den0; sc1 = 1.f / den1; }      // n.pow(-1) (models/attention.py:79)
      else if (p.gate) {
        if (tl0 < tc) sc0 = __ldg(p.gate + r0 * p.H + h);
        if (tl1 < tc) sc1 = __ldg(p.gate + r1 * p.H + h);
      }
      float* o0 = p.out + r0 * p.ldo + (size_t)h * 64 + 2 * t4;
      float* o1 = p.out + r1 * p.ldo + (size_t)h * 64 + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (tl0 < tc) *reinterpret_cast<float2*>(o0 + 8 * nt) = make_float2(sc0 * O[nt][0], sc0 * O[nt][1]);
        if (tl1 < tc) *reinterpret_cast<float2*>(o1 + 8 * nt) = make_float2(sc1 * O[nt][2], sc1 * O[nt][3]);
      }
    }
    // ---- S = S0 + K^T V for the next chunk (rows dd of this role), in registers until the barrier has passed ----
#ifdef LM_ABL_SU
    if (t0 < 0) {
#else
    if (t0 + LM_C < p.T) {
#endif
      {                                                             // next chunk's rows -> L2, in flight under the state update
        const int tcn = (int)min((int64_t)LM_C, p.T - t0 - LM_C);
        const int c4 = tid & 15, seg = tid >> 4;
        // one prefetch per 128-byte half row and tile: the threads with c4 = 0, 8 of each row segment
        if ((c4 & 7) == 0) {
          const int64_t row0 = rowbase + t0 + LM_C + 8 * seg;
          const int nrow = min(8, tcn - 8 * seg);                   // rows of this segment that exist
          const float* pq = qbase + row0 * p.ld + 4 * c4;
          const int64_t dk = kbase - qbase, dv = vbase - qbase;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < nrow) {
              asm volatile("prefetch.global.L2 [%0];" :: "l"(pq));
              asm volatile("prefetch.global.L2 [%0];" :: "l"(pq + dk));
              asm volatile("prefetch.global.L2 [%0];" :: "l"(pq + dv));
            }
            pq += p.ld;
          }
        }
      }
      // the chunk's K^T V goes into a ZEROED accumulator and is added to the state with round-to-nearest FADDs: the tensor core's own fp32 accumulation
      // truncates, and 16 chunks x 24 MMAs into one accumulator left a one-sided ~2e-5 on the state (activation deviation after 12 blocks 1.4e-5 -> 5.5e-6)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f; }
#pragma unroll 1
      for (int kb8 = 0; kb8 < 8; ++kb8) {
        const float* v0 = Vs + (8 * kb8 + 2 * t4) * LM_LD + g;      // token 2 t4 (K index t4) and 2 t4 + 1 (K index t4 + 4) of the block
        const float* k0 = Ks + (8 * kb8 + 2 * t4) * LM_LD + 16 * role + g;
        uint32_t kh[4], kl[4];
        lm_afrag(k0[0], k0[8], k0[LM_LD], k0[LM_LD + 8], kh, kl);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t bh[8], bl[8];
          lm_bfrag4(v0 + 32 * hf, v0 + LM_LD + 32 * hf, bh, bl);
          lm_mma3x4(S[4 * hf], S[4 * hf + 1], S[4 * hf + 2], S[4 * hf + 3], kh, kl, bh, bl);
        }
      }
    }
    if (t0 + LM_C < p.T) {                                          // S = S0 + K^T V (the state rows of this role are read by the other warps until the barrier)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 a = *reinterpret_cast<const float2*>(srow0 + 8 * nt), c = *reinterpret_cast<const float2*>(srow1 + 8 * nt);
        S[nt][0] += a.x; S[nt][1] += a.y; S[nt][2] += c.x; S[nt][3] += c.y;
      }
    }
    __syncthreads();                                                // every warp is done with the tiles and the state at the chunk start
  }
}

bool linattn_mma_supported(const LinAttnParams& p) {
  const bool al16 = p.ld % 4 == 0 && (((uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v) & 15) == 0;
  const bool al8 = p.ldo % 2 == 0 && ((uintptr_t)p.out & 7) == 0;
  return p.d == 64 && p.dv == 64 && al16 && al8 && (p.conv_w == nullptr || (p.kconv >= 1 && p.kconv <= 4 && p.conv_b != nullptr));
}

cudaError_t linattn_mma_launch(const LinAttnParams& p, int64_t B, cudaStream_t stream) {
  const size_t smem = sizeof(float) * LM_SMEM_FLOATS;
  dim3 grid(p.H, (unsigned)B);
  cudaError_t e;
  if (p.conv_w) {
    e = cudaFuncSetAttribute(linattn_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    linattn_mma_kernel<true><<<grid, LM_THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(linattn_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    linattn_mma_kernel<false><<<grid, LM_THREADS, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace eigb200
