// k8_tail_fused_t.cu -- the tail of a Mamba block in the TRANSPOSED orientation of k7_front_fused.cu:
//   x_out = GLU(GELU(y W_out^T + b_out) W_glu^T + b_glu) + x_in   (+ extractor partials),   computed as   D^T = W (.)^T   on 32-token chunks.
//
// Reference operators: MambaBlock.forward, models/mamba.py:333-337 (`self.activation(self.mamba(...))` = GELU(out_proj(y)), then `self.glu(x) + skip`),
// GLU, models/common.py:50-58; the extractor partials feed get_eig_mamba2 of the block OUTPUT (analysis/eval_eig.py:176-190, :512-520).
// Same C-ABI call, operands and results as gemm_out_glu_kernel (k4_gemm_fused.cu); eigb200_out_glu_fused picks one of the two (EIGB200_TAIL_FORM).
//
// Why a second form: gemm_out_glu_kernel works on ONE 128-row tile per CTA at a time -- its accumulators, the TMEM operand of the second GEMM and the
// operand stages fill the 512 TMEM columns -- so every hand-off between its roles is exposed: 40 % of its stall samples are mbarrier waits and it issues at
// 0.52 IPC per scheduler although no pipe is saturated (0.83 ms per C2 layer for 3.2 GB).  With the weights as the A operand (M = output channel = TMEM
// lane) and a 32-token chunk as the B operand (N = token = TMEM column) an accumulator is 32 columns wide, so THREE chunks ("slots") are in flight per
// CTA, each owned by four warps that do all the SIMT work of their chunk (k7's recipe); the tensor core and TMA run under them.
//
// Per CTA (one per SM, 448 threads = 12 slot warps + TMA warp + MMA warp); thread of a slot = output channel c (TMEM lane), its 32 tokens along the columns:
//   start   : weights: the hi halves (fp16, W_out 128 rows, W_glu value rows 128, W_glu gate rows 128) go global -> shared (staging) -> TMEM as A operands;
//             the lo halves stay in shared memory (96 KB) and are the A operand of the lo . hi correction MMAs
//   TMA     : y chunk (32 tokens x 128 fp32, four SWIZZLE_128B boxes) into the slot's stage of the ring, as soon as GEMM 1 of the previous chunk has retired
//   slot, chunk i:  (3) convert chunk i + 1 in place to [fp16 hi | lo] (as k7);  (1) epilogue 2 of chunk i - 1: D2 value / gate -> value * sigmoid(gate) +
//             residual -> coalesced stores (a warp = 128 contiguous bytes per token) and the extractor partials (transposed through the slot's O buffer);
//             (2) epilogue 1 of chunk i: D1 -> + b_out -> S_a GELU -> fp16 hi / lo -> the slot's O buffer [token][channel] = the K-major B operand of GEMM 2
//   MMA     : one thread polls the slots' barriers and issues whatever is ready: GEMM 1 (24 MMAs: hi hi, lo hi, hi lo per K step), GEMM 2 (2 x 24)
// TMEM (512 columns): 3 slots x (D1 32 | D2 value 32 | D2 gate 32) = 288 | weights hi 3 x 64 from column 320.
// Shared memory: 96 KB weights lo + 48 KB ring + 48 KB O buffers + 2 KB biases = 194 KB.
//
// Status: parity-green (tests/test_blocks_gpu.py runs both forms), measured SLOWER than gemm_out_glu_kernel at BASELINE C2 -- 1.00 against 0.83 ms per layer
// (profiles/r2_tail_t_ncu.txt) -- and therefore opt-in (EIGB200_TAIL_FORM=t).  In this orientation a thread owns ONE channel of 32 tokens, so the fp16 hi / lo
// split works on single values (cvt + 2-byte shared stores instead of packed pairs), the residual and the output are 32 strided accesses per thread and the
// extractor partials need a transpose through shared memory: 266 issued warp instructions per token against 234, at the same 0.5 IPC (12 slot warps; 15 % of the
// stall samples are instruction fetch: 98 KB of unrolled code shared by warps in three different phases).
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>

namespace eigb200 {

constexpr int TT_Q = 32;                         // tokens per chunk = N of the MMAs
constexpr int TT_D = 128;                        // d_model = output channels = K of GEMM 2
constexpr int TT_K1 = 128;                       // d_inner = K of GEMM 1
constexpr int TT_SLOTS = 3;
constexpr int TT_SLOT_WARPS = 4 * TT_SLOTS;
constexpr int TT_TMA_WARP = TT_SLOT_WARPS, TT_MMA_WARP = TT_SLOT_WARPS + 1;
constexpr int TT_THREADS = (TT_SLOT_WARPS + 2) * 32;
constexpr int TT_NST = TT_SLOTS;                 // one ring stage per slot: a stage's barriers carry ONE parity bit, so its users must come in order -- its own slot's chunks
constexpr int TT_STAGE_BYTES = TT_Q * TT_K1 * 4; // 16 KB
constexpr uint32_t TT_WLO = 0;                   // W_out lo | W_glu value lo | W_glu gate lo: each [K chunk of 64][128 rows][128 B] = 32 KB
constexpr uint32_t TT_RING = 98304;
constexpr uint32_t TT_OBUF = TT_RING + TT_NST * TT_STAGE_BYTES;        // per slot: [hi: K chunk of 64][32 tokens][128 B] 8 KB | lo 8 KB
constexpr uint32_t TT_BIAS = TT_OBUF + TT_SLOTS * 16384;               // b_out[128] | b_glu value[128] | b_glu gate[128] x -log2 e | w_gate[128]
constexpr uint32_t TT_BARS = TT_BIAS + 2048;
constexpr uint32_t TT_SMEM = TT_BARS + 512;
constexpr uint32_t TT_COL_W = 320;               // TMEM: weights hi: W_out +0, W_glu value +64, W_glu gate +128
constexpr float TT_SA = 16.f;                    // activation pre-scale of both GEMMs (no LayerNorm in front of either), as k4_gemm_fused.cu

struct TtParams {
  const float* bias1; const float* bias2; const float* osc1; const float* osc2;
  float* C; int64_t ldc; const float* R; int64_t ldr;
  const float* eig_w; float* eig_part;
  int64_t M; int nchunks; int* ovf_flag;
};

__device__ __forceinline__ void tt_umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool tt_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tt_tmem_st_32x32u(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
         "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ float4 tt_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 tt_lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tt_sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tt_sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tt_sts_u16(uint32_t addr, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ void tt_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint16_t tt_f16_bits(float x) { uint16_t h; asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x)); return h; }
__device__ __forceinline__ float tt_f16_to_f32(uint16_t h) { float r; asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h)); return r; }

// chunks of slot s of this CTA: q = (TT_SLOTS * k + s) * gridDim.x + blockIdx.x, k = 0, 1, ...
__device__ __forceinline__ int tt_steps(int nchunks, int s) {
  const int64_t first = (int64_t)gridDim.x * s + blockIdx.x;
  if (first >= nchunks) return 0;
  const int64_t stride = (int64_t)gridDim.x * TT_SLOTS;
  return (int)((nchunks - first + stride - 1) / stride);
}

__global__ void __launch_bounds__(TT_THREADS, 1)
tail_t_kernel(const __grid_constant__ CUtensorMap tmapY, const __grid_constant__ CUtensorMap tmapW1h, const __grid_constant__ CUtensorMap tmapW1l,
              const __grid_constant__ CUtensorMap tmapW2h, const __grid_constant__ CUtensorMap tmapW2l, const TtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + TT_BARS;
  const uint32_t bar_w = bars;
  auto bar_raw = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_op = [&](int s) { return bars + 8u * (1 + TT_NST + s); };
  auto bar_free = [&](int s) { return bars + 8u * (1 + 2 * TT_NST + s); };
  constexpr int B0 = 1 + 3 * TT_NST;
  auto bar_d1full = [&](int sl) { return bars + 8u * (B0 + sl); };
  auto bar_d1empty = [&](int sl) { return bars + 8u * (B0 + 3 + sl); };
  auto bar_ofull = [&](int sl) { return bars + 8u * (B0 + 6 + sl); };
  auto bar_d2full = [&](int sl) { return bars + 8u * (B0 + 9 + sl); };
  auto bar_d2empty = [&](int sl) { return bars + 8u * (B0 + 12 + sl); };
  const uint32_t tmem_slot = bars + 8u * (B0 + 15);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* const bias_s = reinterpret_cast<float*>(sm + TT_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < TT_NST; ++s) { mbar_init(bar_raw(s), 1); mbar_init(bar_op(s), 128); mbar_init(bar_free(s), 1); }
    for (int sl = 0; sl < TT_SLOTS; ++sl) {
      mbar_init(bar_d1full(sl), 1); mbar_init(bar_d1empty(sl), 128); mbar_init(bar_ofull(sl), 128);
      mbar_init(bar_d2full(sl), 1); mbar_init(bar_d2empty(sl), 128);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    bias_s[c] = p.bias1 ? p.bias1[c] : 0.f;
    bias_s[128 + c] = p.bias2 ? p.bias2[c] : 0.f;
    bias_s[256 + c] = (p.bias2 ? p.bias2[TT_D + c] : 0.f) * -1.4426950408889634f;
    bias_s[384 + c] = p.eig_w ? p.eig_w[c] : 0.f;
  }
  if (warp == TT_MMA_WARP) tmem_alloc(tmem_slot, 512);
  if (warp == TT_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmapY); tma_prefetch_desc(&tmapW1h); tma_prefetch_desc(&tmapW1l); tma_prefetch_desc(&tmapW2h); tma_prefetch_desc(&tmapW2l);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // ---- weights -------------------------------------------------------------------------------------------------------------------------------------------
  // lo halves -> TT_WLO (resident); hi halves -> staging in the ring / O buffers -> TMEM.  Every matrix: [K chunk of 64][128 rows][128 B].  The prepared
  // W_glu rows come in the order of gemm_tc_ts_kernel's GLU plan (glu_weight_row: per 32 rows 16 value rows then 16 gate rows of 16 output channels):
  // 16-row boxes sort them into a value matrix and a gate matrix whose row = output channel.
  if (warp == TT_TMA_WARP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, 6u * 32768u);
      for (int kch = 0; kch < 2; ++kch) {
        for (int h = 0; h < 2; ++h) {
          tma_load_2d(&tmapW1h, bar_w, base + TT_RING + kch * 16384 + h * 8192, kch * 64, h * 64);
          tma_load_2d(&tmapW1l, bar_w, base + TT_WLO + kch * 16384 + h * 8192, kch * 64, h * 64);
        }
        for (int j = 0; j < 8; ++j) {                                // output channels [16 j, 16 j + 16)
          const int prow = (j >> 2) * 128 + (j & 3) * 32;
          tma_load_2d(&tmapW2h, bar_w, base + TT_RING + 32768 + kch * 16384 + j * 2048, kch * 64, prow);
          tma_load_2d(&tmapW2h, bar_w, base + TT_RING + 65536 + kch * 16384 + j * 2048, kch * 64, prow + 16);
          tma_load_2d(&tmapW2l, bar_w, base + TT_WLO + 32768 + kch * 16384 + j * 2048, kch * 64, prow);
          tma_load_2d(&tmapW2l, bar_w, base + TT_WLO + 65536 + kch * 16384 + j * 2048, kch * 64, prow + 16);
        }
      }
    }
    __syncwarp();
  }
  if (warp < TT_SLOT_WARPS) {
    const int job = warp >> 2, quarter = warp & 3;                                  // job: 0 W_out, 1 W_glu value, 2 W_glu gate
    const int r = quarter * 32 + lane;
    mbar_wait(bar_w, 0);
    const uint32_t src0 = base + TT_RING + (uint32_t)job * 32768u + (uint32_t)r * 128u;
#pragma unroll
    for (int kch = 0; kch < 2; ++kch) {
      uint32_t w[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                                                 // logical 16-byte slot q sits at physical slot q ^ (row & 7)
        const uint4 v = tt_lds_u4(src0 + (uint32_t)kch * 16384u + (uint32_t)((q ^ (r & 7)) * 16));
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      tt_tmem_st_32x32u(tmem_base + ((uint32_t)(quarter * 32) << 16) + TT_COL_W + (uint32_t)(job * 64 + kch * 32), w);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();                                                                  // weights hi are in TMEM: ring and O buffers are free
  tc_fence_after();

  int steps[TT_SLOTS];
#pragma unroll
  for (int s = 0; s < TT_SLOTS; ++s) steps[s] = tt_steps(p.nchunks, s);

  if (warp < TT_SLOT_WARPS) {
    // ===================================== slot warps: slot = warp / 4, quarter = warp % 4, thread = output channel ======================================
    const int slot = warp >> 2, quarter = warp & 3;
    const int ch = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t d1 = tmem_base + lane_sel + (uint32_t)(slot * 96), d2v = d1 + 32u, d2g = d1 + 64u;
    const uint32_t obuf = base + TT_OBUF + (uint32_t)slot * 16384u;
    const float osc1 = __ldg(p.osc1), osc2 = __ldg(p.osc2);
    const float osc2_gate = -1.4426950408889634f * osc2;
    const float b1c = bias_s[ch], b2v = bias_s[128 + ch], b2g = bias_s[256 + ch];
    const int my_steps = steps[slot];
    const int64_t chunk_stride = (int64_t)gridDim.x * TT_SLOTS;
    const int64_t chunk0 = (int64_t)gridDim.x * slot + blockIdx.x;
    float amax = 0.f;
    // O buffer address of (token t, channel ch): K chunk ch / 64, row t (128 B), 16-byte slot ((ch % 64) / 8) ^ (t & 7), byte 2 (ch % 8)
    const uint32_t o_base = obuf + (uint32_t)(ch >> 6) * 4096u + (uint32_t)(ch & 7) * 2u;
    const uint32_t o_slot = (uint32_t)((ch & 63) >> 3);

    auto convert = [&](int i) {                                      // y chunk of step i -> [fp16 hi | lo] in place (thread = (token = lane, box = quarter))
      const int st = slot;
      const uint32_t ph = (uint32_t)i & 1u;
      const uint32_t row = base + TT_RING + st * TT_STAGE_BYTES + (uint32_t)quarter * 4096u + (uint32_t)lane * 128u;
      const int sw = lane & 7;
      mbar_wait(bar_raw(st), ph);
      float a[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = tt_lds_f4(row + (uint32_t)((q ^ sw) * 16));
        a[4 * q] = v.x * TT_SA; a[4 * q + 1] = v.y * TT_SA; a[4 * q + 2] = v.z * TT_SA; a[4 * q + 3] = v.w * TT_SA;
      }
      uint32_t hi2[16], lo2[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        hi2[e] = pack_f16x2(a[2 * e], a[2 * e + 1]);
        amax = fmaxf(amax, fmaxf(fabsf(a[2 * e]), fabsf(a[2 * e + 1])));
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) lo2[e] = pack_f16x2(a[2 * e] - f16_lo_to_f32(hi2[e]), a[2 * e + 1] - f16_hi_to_f32(hi2[e]));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tt_sts_u4(row + (uint32_t)((j ^ sw) * 16), hi2[4 * j], hi2[4 * j + 1], hi2[4 * j + 2], hi2[4 * j + 3]);
        tt_sts_u4(row + (uint32_t)(((4 + j) ^ sw) * 16), lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
      }
      fence_proxy_async();
      mbar_arrive(bar_op(st));
    };

    auto epi1 = [&](int i) {                                         // D1 -> + b_out -> S_a GELU -> fp16 hi / lo -> O buffer (B operand of GEMM 2)
      mbar_wait(bar_d1full(slot), (uint32_t)i & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(d1, v);
      tc_fence_before();
      mbar_arrive(bar_d1empty(slot));
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float z = fmaf(v[t], osc1, b1c);
        const float o = gelu_fast_scaled_f(z, z * TT_SA, TT_SA);
        amax = fmaxf(amax, fabsf(o));
        const uint16_t h = tt_f16_bits(o);
        const uint16_t l = tt_f16_bits(o - tt_f16_to_f32(h));
        const uint32_t addr = o_base + (uint32_t)t * 128u + ((o_slot ^ (uint32_t)(t & 7)) << 4);
        tt_sts_u16(addr, h);
        tt_sts_u16(addr + 8192u, l);
      }
      fence_proxy_async();
      mbar_arrive(bar_ofull(slot));
    };

    auto epi2 = [&](int j) {                                         // D2 -> value * sigmoid(gate) + residual -> stores, extractor partials
      const int64_t row0 = (chunk0 + (int64_t)j * chunk_stride) * TT_Q;
      const int nvalid = (int)min((int64_t)TT_Q, p.M - row0);
      float out[32];
      {
        const float* rp = p.R + row0 * p.ldr + ch;
        if (nvalid == TT_Q) {
#pragma unroll
          for (int t = 0; t < 32; ++t) out[t] = ldg_stream_f1(rp + (int64_t)t * p.ldr);
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) out[t] = t < nvalid ? ldg_stream_f1(rp + (int64_t)t * p.ldr) : 0.f;
        }
      }
      mbar_wait(bar_d2full(slot), (uint32_t)j & 1u);
      tc_fence_after();
      {
        float a[32], g[32];
        tmem_ld_32x32(d2v, a);
        tmem_ld_32x32(d2g, g);
        tc_fence_before();
        mbar_arrive(bar_d2empty(slot));
#pragma unroll
        for (int t = 0; t < 32; ++t) {                               // sigmoid(g + b) = 1 / (1 + 2^((g + b) * -log2 e)), operand scales folded into the FMAs
          float e2;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(g[t], osc2_gate, b2g)));
          out[t] = fmaf(fmaf(a[t], osc2, b2v), fast_rcp_f(1.f + e2), out[t]);
        }
      }
      {
        float* cp = p.C + row0 * p.ldc + ch;
        if (nvalid == TT_Q) {
#pragma unroll
          for (int t = 0; t < 32; ++t) cp[(int64_t)t * p.ldc] = out[t];
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) if (t < nvalid) cp[(int64_t)t * p.ldc] = out[t];
        }
      }
      if (p.eig_part) {
        // per 16 output columns of every row: dot with the gate weights, mean, centred second moment (the contract of eigb200_mamba2_eig_partials).  The
        // thread has 32 tokens of ONE column: transpose through the slot's O buffer (GEMM 2 of this chunk has retired, epilogue 1 of the next chunk has
        // not started): row = channel, float4 slot q (tokens 4 q .. 4 q + 3) at physical slot q ^ (channel & 7); then thread = token sums its warp's columns
        const uint32_t srow = obuf + (uint32_t)ch * 128u;
#pragma unroll
        for (int q = 0; q < 8; ++q) tt_sts_f4(srow + (uint32_t)((q ^ (ch & 7)) * 16), out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        tt_bar_sync(1 + slot, 128);
        const float* sc = reinterpret_cast<const float*>(sm + TT_OBUF + (uint32_t)slot * 16384u);
        const float* wg = bias_s + 384;
#pragma unroll
        for (int hg = 0; hg < 2; ++hg) {                             // the warp's two 16-column groups
          float x[16];
          float dot = 0.f, sum = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int c = quarter * 32 + hg * 16 + k;
            x[k] = sc[c * 32 + (((lane >> 2) ^ (c & 7)) << 2) + (lane & 3)];
            dot = fmaf(x[k], wg[c], dot); sum += x[k];
          }
          const float mean = sum * 0.0625f;
          float m2 = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) { const float d = x[k] - mean; m2 = fmaf(d, d, m2); }
          if (lane < nvalid) {
            float* pp = p.eig_part + (size_t)(quarter * 2 + hg) * 3 * p.M + (row0 + lane);
            pp[0] = dot; pp[p.M] = mean; pp[2 * p.M] = m2;
          }
        }
        tt_bar_sync(1 + slot, 128);                                  // the O buffer may be overwritten by epilogue 1 of the next chunk
      }
    };

    if (my_steps > 0) convert(0);
    for (int i = 0; i < my_steps; ++i) {
      if (i > 0) epi2(i - 1);
      epi1(i);
      if (i + 1 < my_steps) convert(i + 1);                          // its TMA load was issued when GEMM 1 of chunk i retired: a whole iteration ago
    }
    if (my_steps > 0) epi2(my_steps - 1);
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  } else if (warp == TT_TMA_WARP) {
    // ===================================== TMA producer: one stage per slot, refilled as soon as it is free ======================================
    if (elect_one()) {
      int it[TT_SLOTS] = {0, 0, 0};
      int remaining = steps[0] + steps[1] + steps[2];
      while (remaining > 0) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < TT_SLOTS; ++s) {
          if (it[s] < steps[s] && tt_mbar_test(bar_free(s), ((uint32_t)it[s] & 1u) ^ 1u)) {
            const int64_t q = ((int64_t)TT_SLOTS * it[s] + s) * gridDim.x + blockIdx.x;
            mbar_arrive_expect_tx(bar_raw(s), TT_STAGE_BYTES);
            const uint32_t dst = base + TT_RING + s * TT_STAGE_BYTES;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) tma_load_2d(&tmapY, bar_raw(s), dst + kc * 4096, kc * 32, (int)(q * TT_Q));
            ++it[s]; --remaining; any = true;
          }
        }
        if (!any) __nanosleep(100);
      }
    }
    __syncwarp();
  } else if (warp == TT_MMA_WARP) {
    // ===================================== MMA issuer: polls the slots, issues whatever is ready ======================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, TT_Q);
      // every barrier is tested for the phase right after the last one this thread saw complete (per slot, in order): a barrier carries one parity bit
      int i1[TT_SLOTS] = {0, 0, 0}, i2[TT_SLOTS] = {0, 0, 0};
      int remaining = 2 * (steps[0] + steps[1] + steps[2]);
      while (remaining > 0) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < TT_SLOTS; ++s) {
          if (i1[s] < steps[s] && tt_mbar_test(bar_op(s), (uint32_t)i1[s] & 1u) && tt_mbar_test(bar_d1empty(s), ((uint32_t)i1[s] & 1u) ^ 1u)) {
            tc_fence_after();
            const uint32_t stage = base + TT_RING + s * TT_STAGE_BYTES;
            const uint32_t d = tmem_base + (uint32_t)(s * 96);
            const uint32_t a_hi = tmem_base + TT_COL_W;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                         // K step of 16: activations [box of 32 columns][hi 64 B | lo 64 B]
              const uint64_t bh = umma_desc_k_sw128(stage + (uint32_t)(ks >> 1) * 4096u) + 2u * (ks & 1), bl = bh + 4u;
              const uint64_t al = umma_desc_k_sw128(base + TT_WLO + (uint32_t)(ks >> 2) * 16384u) + 2u * (ks & 3);
              umma_f16_ts(d, a_hi + 8u * ks, bh, idesc, ks > 0 ? 1u : 0u);
              tt_umma_f16_ss(d, al, bh, idesc, 1u);
              umma_f16_ts(d, a_hi + 8u * ks, bl, idesc, 1u);
            }
            umma_commit(bar_free(s));
            umma_commit(bar_d1full(s));
            ++i1[s]; --remaining; any = true;
          }
        }
#pragma unroll
        for (int s = 0; s < TT_SLOTS; ++s) {
          if (i2[s] < steps[s]) {
            if (tt_mbar_test(bar_ofull(s), (uint32_t)i2[s] & 1u) && tt_mbar_test(bar_d2empty(s), ((uint32_t)i2[s] & 1u) ^ 1u)) {
              tc_fence_after();
              const uint32_t ob = base + TT_OBUF + (uint32_t)s * 16384u;
#pragma unroll
              for (int tile = 0; tile < 2; ++tile) {                 // value rows, gate rows of W_glu
                const uint32_t d = tmem_base + (uint32_t)(s * 96 + 32 + tile * 32);
                const uint32_t a_hi = tmem_base + TT_COL_W + (uint32_t)(64 * (1 + tile));
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                  const uint64_t bh = umma_desc_k_sw128(ob + (uint32_t)(ks >> 2) * 4096u) + 2u * (ks & 3);
                  const uint64_t bl = umma_desc_k_sw128(ob + 8192u + (uint32_t)(ks >> 2) * 4096u) + 2u * (ks & 3);
                  const uint64_t al = umma_desc_k_sw128(base + TT_WLO + 32768u * (1 + tile) + (uint32_t)(ks >> 2) * 16384u) + 2u * (ks & 3);
                  umma_f16_ts(d, a_hi + 8u * ks, bh, idesc, ks > 0 ? 1u : 0u);
                  tt_umma_f16_ss(d, al, bh, idesc, 1u);
                  umma_f16_ts(d, a_hi + 8u * ks, bl, idesc, 1u);
                }
              }
              umma_commit(bar_d2full(s));
              ++i2[s]; --remaining; any = true;
            }
          }
        }
        if (!any) __nanosleep(40);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TT_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

bool out_glu_fused_t_supported(int D, int K1) { return D == TT_D && K1 == TT_K1; }

int launch_out_glu_fused_t(cudaStream_t st, const float* A, int64_t lda, const void* ws1, const float* bias1, const void* ws2, const float* bias2,
                           float* C, int64_t ldc, const float* R, int64_t ldr, int64_t M, int D, int K1, const float* eig_w, float* eig_part) {
  if (!out_glu_fused_t_supported(D, K1)) { set_error("out_glu_fused (transposed form): needs d_model = d_inner = 128 (D=%d K1=%d)", D, K1); return EIGB200_EUNSUPPORTED; }
  if (lda % 4 != 0 || ((uintptr_t)A & 15) || M >= (1LL << 31) - 64) {
    set_error("out_glu_fused (transposed form): y rows must be 16-byte aligned and M < 2^31"); return EIGB200_EUNSUPPORTED;
  }
  TcPrepared p1, p2;
  if (!tc_prepared_layout_f16(D, K1, EIGB200_EPI_GELU, ws1, &p1) || !tc_prepared_layout_f16(2 * D, D, EIGB200_EPI_GLU_RESIDUAL, ws2, &p2) ||
      p1.nsplit != 1 || p1.bn != 128 || p1.kp64 != 128 || p2.nsplit != 2 || p2.bn != 128 || p2.bg != 64 || p2.kp64 != 128) {
    set_error("out_glu_fused (transposed form): unexpected operand plan for D=%d K1=%d", D, K1); return EIGB200_EUNSUPPORTED;
  }
  CUtensorMap tY, tW1h, tW1l, tW2h, tW2l;
  int rc;
  if ((rc = tc_make_tmap_f32(&tY, A, (uint64_t)M, (uint64_t)K1, (uint64_t)lda, TT_Q))) return rc;
  if ((rc = tc_make_tmap_f16(&tW1h, p1.w_hi, (uint64_t)p1.wrows, (uint64_t)p1.kp64, 64))) return rc;
  if ((rc = tc_make_tmap_f16(&tW1l, p1.w_lo, (uint64_t)p1.wrows, (uint64_t)p1.kp64, 64))) return rc;
  if ((rc = tc_make_tmap_f16(&tW2h, p2.w_hi, (uint64_t)p2.wrows, (uint64_t)p2.kp64, 16))) return rc;
  if ((rc = tc_make_tmap_f16(&tW2l, p2.w_lo, (uint64_t)p2.wrows, (uint64_t)p2.kp64, 16))) return rc;
  TtParams p{};
  p.bias1 = bias1; p.bias2 = bias2; p.osc1 = p1.scal; p.osc2 = p2.scal;
  p.C = C; p.ldc = ldc; p.R = R; p.ldr = ldr; p.eig_w = eig_w; p.eig_part = eig_part;
  p.M = M; p.nchunks = (int)((M + TT_Q - 1) / TT_Q);
  p.ovf_flag = tc_overflow_flag();
  if (!p.ovf_flag) { set_error("out_glu_fused (transposed form): cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  const int64_t grid = p.nchunks < num_sms() ? p.nchunks : num_sms();
  const size_t smem = (size_t)TT_SMEM + 1024 /*alignment*/;
  EIGB_CUDA(cudaFuncSetAttribute(tail_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tail_t_kernel<<<(unsigned)grid, TT_THREADS, smem, st>>>(tY, tW1h, tW1l, tW2h, tW2l, p);
  EIGB_LAUNCH_CHECK("tail_t_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200
