// k7_front_fused.cu -- the FRONT of a Mamba-2 block as ONE kernel:   y = SSD(conv + SiLU(in_proj(LN(x))))   (the in_proj output never reaches HBM)
//
// Reference operators: MambaBlock.forward's prenorm + SSD.forward up to the scan, models/mamba.py:329-331 (`self.norm(x)`), :118-150 (`in_proj`, split into
// [x | B | C | dt], `conv1d` + SiLU on xBC, `softplus(dt + dt_bias)`, `A = -exp(A_log)`, `mamba_chunk_scan_combined(..., D=D)`).
//
// Why: as two kernels (eigb200_linear_ln + eigb200_mamba_conv_ssd) the projection z = [x | B | C | dt] (168 padded floats per token) is written to HBM and
// read back -- 2 x 1.35 GB per layer at BASELINE C2, a third of what the layer moves -- and the tensor-bound GEMM and the FMA-bound recurrence run back to
// back although they need different pipes.  Here the GEMM is computed TRANSPOSED, z^T = W_in LN(x)^T: the weights are the A operand (M = output channel =
// TMEM lane), a 32-token chunk of one sequence is the B operand (N = token = TMEM column).  The accumulator then already has the layout the recurrence
// wants -- thread = channel, its tokens along the columns -- so the scan threads read their channel's chunk straight out of TMEM with tcgen05.ld, run conv +
// SiLU along the registers and the selective-scan recurrence with the state row in registers, and only y is stored.  While the scan warps of one sequence
// work through a chunk, TMA, the converter warps and the tensor core prepare the next chunks of the other resident sequences.
//
// Shape (the C2 / MQAR family): d_model K = 128, d_inner P = 128, one head, one group, d_state N = 16, conv taps <= 4; fp16-split operands (kind::f16, prepared
// by eigb200_linear_prepare with the LayerNorm folded in, see k4_gemm_tc.cu).
//
// Per CTA (one per SM, 704 threads), FF_SLOTS = 4 sequences resident at a time, each walking its 32-token chunks in order:
//   TMA warp    : x chunk (32 tokens x 128 fp32 = 16 KB, four SWIZZLE_128B boxes) into a 6-stage ring, round-robin over the slots
//   converters  : 4 warps, thread = (token, 32-column box): (a - mu) rstd S_a -> fp16 hi / lo -> written IN PLACE over the raw chunk as the K-major
//                 SWIZZLE_128B B operand [K chunk of 64][32 tokens][128 B] (hi 8 KB | lo 8 KB), fence.proxy.async, mbarrier
//   MMA warp    : per chunk two M = 128, N = 32 accumulators: tile 1 = the 128 x channels, tile 2 = rows [B 16 | C 16 | dt 1 | zero padding] of W_in;
//                 3 kind::f16 MMAs per K step (hi hi, lo hi, hi lo), A (weights, resident, 96 KB) and B from shared memory
//   scan warps  : 4 per slot (warp % 4 = TMEM lane quarter), thread = channel.  Warp 0 of the slot first turns tile 2 lanes 0-31 into conv + SiLU'ed
//                 B_t / C_t rows in shared memory, warp 1 turns lane 32 (dt) into (dt, e^{dt A}, E_t, 1 / E_t); then all four pull their channel's tokens
//                 8 at a time from TMEM and run the recurrence exactly as ssd_scan_v3 does (rescaled-state form, direct form on decay underflow).
// TMEM: 4 slots x (32 + 32) columns.  Shared memory: 96 KB weights + 96 KB ring + 16 KB B/C rows + 3 KB = 212 KB.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace eigb200 {

constexpr int FF_Q = 32;                         // tokens per chunk = N of the MMAs
constexpr int FF_K = 128;                        // d_model
constexpr int FF_P = 128;                        // d_inner = x channels = lanes of tile 1
constexpr int FF_N = 16;                         // d_state
constexpr int FF_SLOTS = 4;
constexpr int FF_SCAN_WARPS = 4 * FF_SLOTS;
constexpr int FF_CONV_WARP0 = FF_SCAN_WARPS;
constexpr int FF_TMA_WARP = FF_SCAN_WARPS + 4, FF_MMA_WARP = FF_SCAN_WARPS + 5;
constexpr int FF_THREADS = (FF_SCAN_WARPS + 6) * 32;
constexpr int FF_NST = 6;
constexpr int FF_STAGE_BYTES = FF_Q * FF_K * 4;  // 16 KB: raw fp32 chunk = fp16 hi (8 KB) + lo (8 KB) operand
constexpr int FF_G = 8;                          // tokens per unrolled group of the recurrence
constexpr uint32_t FF_W_HI = 0, FF_W_LO = 49152, FF_RING = 98304;
constexpr uint32_t FF_BC = FF_RING + FF_NST * FF_STAGE_BYTES;          // [slot][token][B 16 | C 16] fp32
constexpr uint32_t FF_DD = FF_BC + FF_SLOTS * FF_Q * 2 * FF_N * 4;     // [slot][token] (dt, decay, E, 1 / E)
constexpr uint32_t FF_DTRAW = FF_DD + FF_SLOTS * FF_Q * 16;            // [slot][token] raw dt accumulator
constexpr uint32_t FF_FLAGS = FF_DTRAW + FF_SLOTS * FF_Q * 4;          // [slot] chunk takes the direct form
constexpr uint32_t FF_BARS = FF_FLAGS + 64;
constexpr uint32_t FF_SMEM = FF_BARS + 512;
constexpr float FF_SA = 1024.f;                  // activation pre-scale behind a LayerNorm (tc_prepare)

struct FfParams {
  const float2* ln_stats;                        // (M) (mean, rstd)
  const float* bias2; const float* osc;          // folded bias b + W beta (161), 1 / (S_a S_w)
  const float* conv_w; const float* conv_b; int kconv;
  const float* dt_bias; const float* A_log; const float* D;
  float* y; int64_t ldy;
  int64_t B, T, M; int nchunks; int zero; int* ovf_flag;
};

__device__ __forceinline__ void ff_umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ff_tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ff_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float4 ff_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ff_sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ff_sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void ff_sts_f1(uint32_t addr, float a) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(a) : "memory"); }
__device__ __forceinline__ float ff_lds_f1(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ int ff_lds_i1(uint32_t addr) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void ff_sts_i1(uint32_t addr, int a) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(a) : "memory"); }
__device__ __forceinline__ float ff_silu(float z) { return z * sigmoid_fast_f(z); }

// sequences of slot s of this CTA: b = blockIdx.x + gridDim.x * (FF_SLOTS * k + s), k = 0, 1, ...
__device__ __forceinline__ int ff_nseq(int64_t B, int s) {
  const int64_t first = (int64_t)blockIdx.x + (int64_t)gridDim.x * s;
  if (first >= B) return 0;
  const int64_t stride = (int64_t)gridDim.x * FF_SLOTS;
  return (int)((B - first + stride - 1) / stride);
}

__global__ void __launch_bounds__(FF_THREADS, 1)
mamba_front_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapWh, const __grid_constant__ CUtensorMap tmapWl,
                   const FfParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + FF_BARS;
  const uint32_t bar_w = bars;
  auto bar_raw = [&](int s) { return bars + 8u * (1 + s); };                        // TMA landed the raw chunk
  auto bar_op = [&](int s) { return bars + 8u * (1 + FF_NST + s); };                // converters wrote the operand
  auto bar_free = [&](int s) { return bars + 8u * (1 + 2 * FF_NST + s); };          // MMAs that read the stage retired
  auto bar_dfull = [&](int sl) { return bars + 8u * (1 + 3 * FF_NST + sl); };       // both accumulator tiles of the slot's chunk are complete
  auto bar_dempty = [&](int sl) { return bars + 8u * (1 + 3 * FF_NST + FF_SLOTS + sl); };   // the scan warps have pulled them into registers
  const uint32_t tmem_slot = bars + 8u * (1 + 3 * FF_NST + 2 * FF_SLOTS);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = p.nchunks;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < FF_NST; ++s) { mbar_init(bar_raw(s), 1); mbar_init(bar_op(s), 128); mbar_init(bar_free(s), 1); }
    for (int sl = 0; sl < FF_SLOTS; ++sl) { mbar_init(bar_dfull(sl), 1); mbar_init(bar_dempty(sl), 128); }
    fence_barrier_init();
  }
  if (warp == FF_MMA_WARP) tmem_alloc(tmem_slot, 256);
  if (warp == FF_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmapX); tma_prefetch_desc(&tmapWh); tma_prefetch_desc(&tmapWl); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int nseq[FF_SLOTS];
#pragma unroll
  for (int s = 0; s < FF_SLOTS; ++s) nseq[s] = ff_nseq(p.B, s);
  const int nsteps = nseq[0] * nchunks;                              // slot 0 never has fewer sequences than the others

  if (warp == FF_TMA_WARP) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, 2u * 49152u);
      for (int kch = 0; kch < 2; ++kch) {                            // tile 1: rows 0-127 (two 64-row boxes); tile 2: rows 128-191 (176+ are out of bounds: zero)
        tma_load_2d(&tmapWh, bar_w, base + FF_W_HI + kch * 16384, kch * 64, 0);
        tma_load_2d(&tmapWh, bar_w, base + FF_W_HI + kch * 16384 + 8192, kch * 64, 64);
        tma_load_2d(&tmapWh, bar_w, base + FF_W_HI + 32768 + kch * 8192, kch * 64, 128);
        tma_load_2d(&tmapWl, bar_w, base + FF_W_LO + kch * 16384, kch * 64, 0);
        tma_load_2d(&tmapWl, bar_w, base + FF_W_LO + kch * 16384 + 8192, kch * 64, 64);
        tma_load_2d(&tmapWl, bar_w, base + FF_W_LO + 32768 + kch * 8192, kch * 64, 128);
      }
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < nsteps; ++i) {
        const int k = i / nchunks, c = i - k * nchunks;
#pragma unroll
        for (int s = 0; s < FF_SLOTS; ++s) {
          if (k >= nseq[s]) continue;
          const int64_t b = (int64_t)blockIdx.x + (int64_t)gridDim.x * (FF_SLOTS * k + s);
          const int row0 = (int)(b * p.T + (int64_t)c * FF_Q);
          mbar_wait_one(bar_free(st), ph ^ 1);
          mbar_arrive_expect_tx(bar_raw(st), FF_STAGE_BYTES);
          const uint32_t dst = base + FF_RING + st * FF_STAGE_BYTES;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) tma_load_2d(&tmapX, bar_raw(st), dst + kc * 4096, kc * 32, row0);
          if (++st == FF_NST) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= FF_CONV_WARP0 && warp < FF_CONV_WARP0 + 4) {
    // ===================================== converters: thread = (token r, 32-column box kc) ======================================
    const int r = lane, kc = warp - FF_CONV_WARP0;
    const int sw = r & 7;
    const uint32_t src_off = (uint32_t)kc * 4096u + (uint32_t)r * 128u;
    const uint32_t dst_off = (uint32_t)(kc >> 1) * 4096u + (uint32_t)r * 128u;
    const int slot0 = (kc & 1) * 4;                                  // first 16-byte slot of this thread's 32 halfs inside the 128-byte operand row
    float amax = 0.f;
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < nsteps; ++i) {
      const int k = i / nchunks, c = i - k * nchunks;
#pragma unroll
      for (int s = 0; s < FF_SLOTS; ++s) {
        if (k >= nseq[s]) continue;
        const int64_t b = (int64_t)blockIdx.x + (int64_t)gridDim.x * (FF_SLOTS * k + s);
        const int64_t m = b * p.T + (int64_t)c * FF_Q + r;
        float2 stt = m < p.M ? __ldg(p.ln_stats + m) : make_float2(0.f, 0.f);
        stt.x = -stt.x * stt.y;                                      // (a - mu) rstd = fma(a, rstd, -mu rstd), as the converter of gemm_tc_ts_kernel
        stt.x *= FF_SA; stt.y *= FF_SA;
        const uint32_t stage = base + FF_RING + st * FF_STAGE_BYTES;
        mbar_wait(bar_raw(st), ph);
        float a[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {                                // logical 16-byte slot q sits at physical slot q ^ (row & 7)
          const float4 v = ff_lds_f4(stage + src_off + (uint32_t)((q ^ sw) * 16));
          a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) a[e] = fmaf(a[e], stt.y, stt.x);
        uint32_t hi2[16], lo2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          hi2[e] = pack_f16x2(a[2 * e], a[2 * e + 1]);
          amax = fmaxf(amax, fmaxf(fabsf(a[2 * e]), fabsf(a[2 * e + 1])));
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) lo2[e] = pack_f16x2(a[2 * e] - f16_lo_to_f32(hi2[e]), a[2 * e + 1] - f16_hi_to_f32(hi2[e]));
        ff_bar_sync(5, 128);                                         // every converter thread has its raw values in registers: the chunk may be overwritten
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t o = stage + dst_off + (uint32_t)(((slot0 + j) ^ sw) * 16);
          ff_sts_u4(o, hi2[4 * j], hi2[4 * j + 1], hi2[4 * j + 2], hi2[4 * j + 3]);
          ff_sts_u4(o + 8192u, lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
        }
        fence_proxy_async();                                         // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_op(st));
        if (++st == FF_NST) { st = 0; ph ^= 1; }
      }
    }
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  } else if (warp == FF_MMA_WARP) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, FF_Q);
      mbar_wait_one(bar_w, 0);
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < nsteps; ++i) {
        const int k = i / nchunks;
#pragma unroll
        for (int s = 0; s < FF_SLOTS; ++s) {
          if (k >= nseq[s]) continue;
          mbar_wait_one(bar_op(st), ph);
          mbar_wait_one(bar_dempty(s), (uint32_t)(i & 1) ^ 1u);       // chunk i - 1 of this slot is in the scan warps' registers
          tc_fence_after();
          const uint32_t stage = base + FF_RING + st * FF_STAGE_BYTES;
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const uint32_t d = tmem_base + (uint32_t)(s * 64 + tile * 32);
#pragma unroll
            for (int kch = 0; kch < 2; ++kch) {
              const uint32_t wofs = tile == 0 ? (uint32_t)kch * 16384u : 32768u + (uint32_t)kch * 8192u;
              const uint64_t ah = umma_desc_k_sw128(base + FF_W_HI + wofs), al = umma_desc_k_sw128(base + FF_W_LO + wofs);
              const uint64_t bh = umma_desc_k_sw128(stage + kch * 4096), bl = umma_desc_k_sw128(stage + 8192 + kch * 4096);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                ff_umma_f16_ss(d, ah + 2u * ks, bh + 2u * ks, idesc, (kch > 0 || ks > 0) ? 1u : 0u);
                ff_umma_f16_ss(d, al + 2u * ks, bh + 2u * ks, idesc, 1u);
                ff_umma_f16_ss(d, ah + 2u * ks, bl + 2u * ks, idesc, 1u);
              }
            }
          }
          umma_commit(bar_free(st));
          umma_commit(bar_dfull(s));
          if (++st == FF_NST) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================== scan warps: slot = warp / 4, thread = channel ======================================
    const int slot = warp >> 2, quarter = warp & 3;
    const int ch = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t d_x = tmem_base + lane_sel + (uint32_t)(slot * 64), d_bc = d_x + 32u;
    const uint32_t bc_s = base + FF_BC + (uint32_t)slot * (FF_Q * 2 * FF_N * 4);
    const uint32_t dd_s = base + FF_DD + (uint32_t)slot * (FF_Q * 16);
    const uint32_t dtraw_s = base + FF_DTRAW + (uint32_t)slot * (FF_Q * 4);
    const uint32_t flag_s = base + FF_FLAGS + (uint32_t)slot * 4;
    const float osc = __ldg(p.osc);
    const float Ah = -expf(__ldg(p.A_log));
    const float Dh = p.D ? __ldg(p.D) : 0.f;
    const float dtb = __ldg(p.dt_bias);
    const int kconv = p.kconv;
    float cw[4], cb;
    {
#pragma unroll
      for (int j = 0; j < 4; ++j) cw[j] = (j >= 4 - kconv) ? __ldg(p.conv_w + (size_t)ch * kconv + j - (4 - kconv)) : 0.f;
      cb = __ldg(p.conv_b + ch);
    }
    const float bias_x = __ldg(p.bias2 + ch);
    // warp 0 of the slot: conv channel 128 + lane (B_0..15, C_0..15); warp 1: the dt row
    float bw[4] = {0.f, 0.f, 0.f, 0.f}, bbias = 0.f, bias_bc = 0.f, bias_dt = 0.f;
    if (quarter == 0) {
      const int bch = FF_P + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) bw[j] = (j >= 4 - kconv) ? __ldg(p.conv_w + (size_t)bch * kconv + j - (4 - kconv)) : 0.f;
      bbias = __ldg(p.conv_b + bch);
      bias_bc = __ldg(p.bias2 + bch);
    } else if (quarter == 1) {
      bias_dt = __ldg(p.bias2 + FF_P + 2 * FF_N);
    }
    float s[FF_N];
    float h1 = 0.f, h2 = 0.f, h3 = 0.f;                              // raw x of the tokens t-1, t-2, t-3
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;                              // the same for this thread's B / C channel (warp 0 of the slot)
    const int my_steps = nseq[slot] * nchunks;
    for (int i = 0; i < my_steps; ++i) {
      const int k = i / nchunks, c = i - k * nchunks;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)gridDim.x * (FF_SLOTS * k + slot);
      const int64_t t0 = (int64_t)c * FF_Q;
      const int tc = (int)min((int64_t)FF_Q, p.T - t0);
      if (c == 0) {
#pragma unroll
        for (int n = 0; n < FF_N; ++n) s[n] = 0.f;
        h1 = h2 = h3 = 0.f; g1 = g2 = g3 = 0.f;
      }
      mbar_wait(bar_dfull(slot), (uint32_t)(i & 1));
      tc_fence_after();
      if (quarter == 0) {
        float v[32];
        tmem_ld_32x32(d_bc, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float raw = fmaf(v[j], osc, bias_bc);
          float o = fmaf(bw[3], raw, fmaf(bw[2], g1, fmaf(bw[1], g2, fmaf(bw[0], g3, bbias))));
          o = ff_silu(o);
          g3 = g2; g2 = g1; g1 = raw;
          ff_sts_f1(bc_s + (uint32_t)(j * 2 * FF_N + lane) * 4u, o);
        }
      } else if (quarter == 1) {
        float v[32];
        tmem_ld_32x32(d_bc, v);                                      // lane 0 of this warp = TMEM lane 32 = the dt row
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ff_sts_f4(dtraw_s + 16u * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
        const float z = fmaf(ff_lds_f1(dtraw_s + 4u * lane), osc, bias_dt);
        const float d = (lane < tc) ? softplus_f(z + dtb) : 0.f;
        const float dec = (lane < tc) ? expf(d * Ah) : 1.f;
        float E = dec;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, E, o); if (lane >= o) E *= u; }   // inclusive product scan
        const float Emin = __shfl_sync(0xffffffffu, E, 31);          // decays <= 1: the last product is the smallest
        if (lane == 0) ff_sts_i1(flag_s, (Emin < 0x1p-60f || !(Emin == Emin)) ? 1 : 0);
        ff_sts_f4(dd_s + 16u * lane, d, dec, E, 1.f / E);
        __syncwarp();
      }
      ff_bar_sync(1 + slot, 128);                                    // B / C rows and (dt, decay, E, 1 / E) of the chunk are in shared memory
      const bool direct = ff_lds_i1(flag_s) != 0;
      float* yp = p.y + (size_t)(b * p.T + t0) * p.ldy + ch;
#pragma unroll 1
      for (int g = 0; g < FF_Q / FF_G; ++g) {
        float xr[FF_G];
        ff_tmem_ld_32x8(d_x + (uint32_t)(g * FF_G), xr);
        if (g == FF_Q / FF_G - 1) { tc_fence_before(); mbar_arrive(bar_dempty(slot)); }     // the tensor core may overwrite both tiles of this slot
#pragma unroll
        for (int j = 0; j < FF_G; ++j) xr[j] = fmaf(xr[j], osc, bias_x);
        if (!direct) {
#pragma unroll
          for (int j = 0; j < FF_G; ++j) {
            const int tt = g * FF_G + j;
            const float4 dd = ff_lds_f4(dd_s + 16u * tt);
            const float xm1 = j >= 1 ? xr[j - 1] : h1, xm2 = j >= 2 ? xr[j - 2] : (j == 1 ? h1 : h2), xm3 = j >= 3 ? xr[j - 3] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
            float xv = fmaf(cw[3], xr[j], fmaf(cw[2], xm1, fmaf(cw[1], xm2, fmaf(cw[0], xm3, cb))));
            xv = ff_silu(xv);
            const float w0 = (xv * dd.x) * dd.w;                     // dt x / E_t
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int q = 0; q < FF_N / 4; ++q) {
              const float4 bv = ff_lds_f4(bc_s + (uint32_t)(tt * 2 * FF_N + 4 * q) * 4u);
              const float4 cv = ff_lds_f4(bc_s + (uint32_t)(tt * 2 * FF_N + FF_N + 4 * q) * 4u);
              s[4 * q + 0] = fmaf(w0, bv.x, s[4 * q + 0]); a0 = fmaf(cv.x, s[4 * q + 0], a0);
              s[4 * q + 1] = fmaf(w0, bv.y, s[4 * q + 1]); a1 = fmaf(cv.y, s[4 * q + 1], a1);
              s[4 * q + 2] = fmaf(w0, bv.z, s[4 * q + 2]); a2 = fmaf(cv.z, s[4 * q + 2], a2);
              s[4 * q + 3] = fmaf(w0, bv.w, s[4 * q + 3]); a3 = fmaf(cv.w, s[4 * q + 3], a3);
            }
            const float yv = fmaf(Dh, xv, dd.z * ((a0 + a1) + (a2 + a3)));
            if (tt < tc) yp[(size_t)tt * p.ldy] = yv;
          }
        } else {
#pragma unroll
          for (int j = 0; j < FF_G; ++j) {
            const int tt = g * FF_G + j;
            const float4 dd = ff_lds_f4(dd_s + 16u * tt);
            const float xm1 = j >= 1 ? xr[j - 1] : h1, xm2 = j >= 2 ? xr[j - 2] : (j == 1 ? h1 : h2), xm3 = j >= 3 ? xr[j - 3] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
            float xv = fmaf(cw[3], xr[j], fmaf(cw[2], xm1, fmaf(cw[1], xm2, fmaf(cw[0], xm3, cb))));
            xv = ff_silu(xv);
            const float uu = xv * dd.x;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int q = 0; q < FF_N / 4; ++q) {
              const float4 bv = ff_lds_f4(bc_s + (uint32_t)(tt * 2 * FF_N + 4 * q) * 4u);
              const float4 cv = ff_lds_f4(bc_s + (uint32_t)(tt * 2 * FF_N + FF_N + 4 * q) * 4u);
              s[4 * q + 0] = fmaf(dd.y, s[4 * q + 0], uu * bv.x); a0 = fmaf(cv.x, s[4 * q + 0], a0);
              s[4 * q + 1] = fmaf(dd.y, s[4 * q + 1], uu * bv.y); a1 = fmaf(cv.y, s[4 * q + 1], a1);
              s[4 * q + 2] = fmaf(dd.y, s[4 * q + 2], uu * bv.z); a2 = fmaf(cv.z, s[4 * q + 2], a2);
              s[4 * q + 3] = fmaf(dd.y, s[4 * q + 3], uu * bv.w); a3 = fmaf(cv.w, s[4 * q + 3], a3);
            }
            const float yv = fmaf(Dh, xv, (a0 + a1) + (a2 + a3));
            if (tt < tc) yp[(size_t)tt * p.ldy] = yv;
          }
        }
        h3 = xr[FF_G - 3]; h2 = xr[FF_G - 2]; h1 = xr[FF_G - 1];
      }
      if (!direct) {                                                 // back to the true state: S = r E
        const float Eend = ff_lds_f4(dd_s + 16u * (FF_Q - 1)).z;
#pragma unroll
        for (int n = 0; n < FF_N; ++n) s[n] *= Eend;
      }
      ff_bar_sync(1 + slot, 128);                                    // all four warps are done with the chunk's shared rows
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FF_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

bool mamba_front_fused_supported(int D, int d_inner, int H, int G, int N, int kconv) {
  return D == FF_K && d_inner == FF_P && H == 1 && G == 1 && N == FF_N && kconv >= 1 && kconv <= 4;
}

int launch_mamba_front_fused(cudaStream_t st, const float* x, int64_t ldx, const float* ln_stats, const void* ws_in,
                             const float* conv_w, const float* conv_b, int kconv, const float* dt_bias, const float* A_log, const float* D,
                             float* y, int64_t ldy, int64_t B, int64_t T) {
  const int64_t M = B * T;
  if (ldx % 4 != 0 || ((uintptr_t)x & 15) || M >= (1LL << 31) - 64) {
    set_error("mamba_front_fused: x rows must be 16-byte aligned and B*T < 2^31"); return EIGB200_EUNSUPPORTED;
  }
  const int n_in = FF_P + 2 * FF_N + 1;
  TcPrepared pw;
  if (!tc_prepared_layout_f16(n_in, FF_K, EIGB200_EPI_NONE, ws_in, &pw) || pw.nsplit != 1 || pw.kp64 != 128 || pw.wrows < n_in || pw.wrows > 192) {
    set_error("mamba_front_fused: unexpected operand plan for in_proj (N=%d K=%d)", n_in, FF_K); return EIGB200_EUNSUPPORTED;
  }
  CUtensorMap tX, tWh, tWl;
  int rc;
  if ((rc = tc_make_tmap_f32(&tX, x, (uint64_t)M, (uint64_t)FF_K, (uint64_t)ldx, FF_Q))) return rc;
  if ((rc = tc_make_tmap_f16(&tWh, pw.w_hi, (uint64_t)pw.wrows, (uint64_t)pw.kp64, 64))) return rc;
  if ((rc = tc_make_tmap_f16(&tWl, pw.w_lo, (uint64_t)pw.wrows, (uint64_t)pw.kp64, 64))) return rc;
  FfParams p{};
  p.ln_stats = reinterpret_cast<const float2*>(ln_stats);
  p.bias2 = pw.bias2; p.osc = pw.scal;
  p.conv_w = conv_w; p.conv_b = conv_b; p.kconv = kconv;
  p.dt_bias = dt_bias; p.A_log = A_log; p.D = D;
  p.y = y; p.ldy = ldy; p.B = B; p.T = T; p.M = M;
  p.nchunks = (int)((T + FF_Q - 1) / FF_Q); p.zero = 0;
  p.ovf_flag = tc_overflow_flag();
  if (!p.ovf_flag) { set_error("mamba_front_fused: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  const int64_t grid = B < (int64_t)num_sms() ? B : (int64_t)num_sms();
  const size_t smem = (size_t)FF_SMEM + 1024 /*alignment*/;
  EIGB_CUDA(cudaFuncSetAttribute(mamba_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mamba_front_kernel<<<(unsigned)grid, FF_THREADS, smem, st>>>(tX, tWh, tWl, p);
  EIGB_LAUNCH_CHECK("mamba_front_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_mamba_front_fused_supported(int D, int d_inner, int H, int G, int N, int kconv) {
  return mamba_front_fused_supported(D, d_inner, H, G, N, kconv) ? 1 : 0;
}

extern "C" int eigb200_mamba_front_fused(void* stream, const float* d_x, int64_t ldx, const float* d_ln_stats, const void* d_ws_in,
                                         const float* d_conv_w, const float* d_conv_b, int kconv, const float* d_dt_bias, const float* d_A_log,
                                         const float* d_D, float* d_y, int64_t ldy, int64_t B, int64_t T, int D, int d_inner, int N) {
  EIGB_CHECK_ARG(d_x && d_ln_stats && d_ws_in && d_conv_w && d_conv_b && d_dt_bias && d_A_log && d_y, "mamba_front_fused: null pointer");
  EIGB_CHECK_ARG(B > 0 && T > 0, "mamba_front_fused: bad shape B=%lld T=%lld", (long long)B, (long long)T);
  EIGB_CHECK_ARG(mamba_front_fused_supported(D, d_inner, 1, 1, N, kconv),
                 "mamba_front_fused: needs d_model = d_inner = 128, one head, one group, d_state = 16, 1..4 conv taps (D=%d d_inner=%d N=%d kconv=%d)", D, d_inner, N, kconv);
  EIGB_CHECK_ARG(ldx >= D && ldy >= d_inner, "mamba_front_fused: row stride smaller than the row");
  EIGB_CHECK_ARG(tc_default_kind() == 1, "mamba_front_fused: the prepared operands must be the fp16 split (EIGB200_GEMM_PRECISION=f16x3)");
  return launch_mamba_front_fused((cudaStream_t)stream, d_x, ldx, d_ln_stats, d_ws_in, d_conv_w, d_conv_b, kconv, d_dt_bias, d_A_log, d_D, d_y, ldy, B, T);
}
