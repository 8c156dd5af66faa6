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
// work through a chunk, TMA, the converter warps and the tensor core prepare the next chunks of the resident sequences.
//
// Shape (the C2 / MQAR family): d_model K = 128, d_inner P = 128, one head, one group, d_state N = 16, conv taps <= 4; fp16-split operands (kind::f16, prepared
// by eigb200_linear_prepare with the LayerNorm folded in, see k4_gemm_tc.cu).
//
// Per CTA (one per SM, 576 threads): FF_SLOTS = 4 sequences resident at a time, each owned by 4 warps (warp % 4 = TMEM lane quarter, thread = channel) that
// walk the sequence's 32-token chunks in order and do ALL the SIMT work of their sequence; two more warps issue TMA and MMA for the four slots.
//   start       : the prepared weights (hi / lo fp16, 161 rows padded to 192) go global -> shared (TMA, staged in the ring) -> TMEM (tcgen05.st, lane = weight
//                 row): the A operand of every MMA then comes from TMEM -- with A in shared memory each M 128 x N 32 x K 16 MMA fetched 5 KB of operands and the
//                 48 MMAs of a chunk took 240 KB of shared-memory bandwidth, more than the recurrence's own broadcast loads (measured: 45 cycles per MMA).
//                 Tile 1 = the 128 x channels.  Tile 2: lanes 0-31, 32-63 and 64-95 each hold the rows [B 16 | C 16] (three copies: the lanes would otherwise
//                 idle, and every lane quarter can then prepare a third of the chunk's tokens), lane 96 holds the dt row.
//   TMA warp    : x chunk (32 tokens x 128 fp32 = 16 KB, four SWIZZLE_128B boxes) + its 32 (mean, rstd) pairs (bulk copy) into an 8-stage ring
//   MMA warp    : per chunk two M = 128, N = 32 accumulators (tile 1, tile 2), 3 kind::f16 MMAs per K step (hi hi, lo hi, hi lo), A from TMEM, B from the ring
//   slot warps, per chunk i of their sequence:
//     convert   : chunk i + 1 (already landed): thread = (token, 32-column box): (a - mu) rstd S_a -> fp16 hi / lo, written IN PLACE over the 128 bytes the
//                 thread read as [hi 64 B | lo 64 B]: the box stays a SWIZZLE_128B K-major [32 tokens][128 B] tile, the hi / lo operands of a K step are
//                 32-byte slices of it; fence.proxy.async, mbarrier -> the tensor core projects chunk i + 1 under the recurrence of chunk i
//     prepare   : tile 2 of chunk i from TMEM: quarters 0-2 conv + SiLU of the B_t / C_t rows of their third of the tokens -> shared memory; quarter 3 the dt
//                 row: softplus, decay, running decay product E_t -> planes dt, e^{dt A}, E_t, dt / E_t; one named barrier of the slot's 128 threads
//     scan      : pull the channel's tokens 16 at a time from TMEM (tcgen05.ld; the accumulators go back to the tensor core after the second pull), conv +
//                 SiLU along the registers, then the recurrence as ssd_scan_v3 runs it (rescaled-state form r_t = S_t / E_t, 2 FMA-pipe operations per state
//                 element; direct form for a chunk whose decay product underflows), y stored straight from registers (a warp = 128 contiguous bytes per token)
// An earlier version gave conversion and B / C preparation to dedicated warps: each of them ran latency-bound at ~7 cycles per instruction next to four scan
// warps on its scheduler and the single B / C warp (3 900 cycles per chunk) paced the whole kernel (tools/front_trace.py time lines, profiles/).
// TMEM (512 columns): 4 slots x (32 + 32) accumulator columns | weights: tile 1 hi 64, lo 64, tile 2 hi 64, lo 64 (two fp16 per column).
// Shared memory: 128 KB ring + 32 KB B / C rows + 4 KB dt planes + 2 KB statistics = 166 KB.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include <type_traits>
#include <cstdio>
#include <cstdlib>

namespace eigb200 {

constexpr int FF_Q = 32;                         // tokens per chunk = N of the MMAs
constexpr int FF_K = 128;                        // d_model
constexpr int FF_P = 128;                        // d_inner = x channels = lanes of tile 1
constexpr int FF_N = 16;                         // d_state
constexpr int FF_SLOTS = 4;
constexpr int FF_SCAN_WARPS = 4 * FF_SLOTS;
constexpr int FF_TMA_WARP = FF_SCAN_WARPS, FF_MMA_WARP = FF_SCAN_WARPS + 1;
constexpr int FF_THREADS = (FF_SCAN_WARPS + 2) * 32;
#ifndef FF_NST_OVERRIDE
constexpr int FF_NST = 8;
#else
constexpr int FF_NST = FF_NST_OVERRIDE;
#endif
static_assert(FF_NST >= 6, "the ring doubles as the 96 KB staging area of the weights");
constexpr int FF_STAGE_BYTES = FF_Q * FF_K * 4;  // 16 KB: raw fp32 chunk = fp16 hi (8 KB) + lo (8 KB) operand
#ifndef FF_H_OVERRIDE
constexpr int FF_H = 8;                          // tokens pulled from TMEM and processed per trip of the recurrence loop (16: 1.02 ms, 8: 0.91 ms per C2 layer:
                                                 // the 8 registers it frees let ptxas issue the shared-memory operand loads further ahead)
#else
constexpr int FF_H = FF_H_OVERRIDE;
#endif
#ifndef FF_G_OVERRIDE
constexpr int FF_G = FF_H < 8 ? FF_H : 8;        // tokens per phase A / phase B group inside a trip
#else
constexpr int FF_G = FF_G_OVERRIDE;
#endif
constexpr uint32_t FF_RING = 0;
constexpr uint32_t FF_BC = FF_RING + FF_NST * FF_STAGE_BYTES;          // [slot][buffer][token][B 16 | C 16] fp32
constexpr uint32_t FF_BC_BYTES = FF_Q * 2 * FF_N * 4;                  // 4 KB per (slot, buffer)
constexpr uint32_t FF_DD = FF_BC + FF_SLOTS * 2 * FF_BC_BYTES;         // [slot][buffer][plane: dt, decay, E, dt / E][token]
constexpr uint32_t FF_DD_BYTES = 4 * FF_Q * 4;
constexpr uint32_t FF_DTRAW = FF_DD + FF_SLOTS * 2 * FF_DD_BYTES;      // [slot][token] raw dt accumulator (scratch of the slot's dt warp)
constexpr uint32_t FF_FLAGS = FF_DTRAW + FF_SLOTS * FF_Q * 4;                     // [slot][buffer] chunk takes the direct form
constexpr uint32_t FF_STATS = FF_FLAGS + 64;                           // [stage][token] (mean, rstd): bulk-copied next to the x chunk
constexpr uint32_t FF_BARS = FF_STATS + FF_NST * FF_Q * 8;
constexpr uint32_t FF_SMEM = FF_BARS + 512;
constexpr float FF_SA = 1024.f;                  // activation pre-scale behind a LayerNorm (tc_prepare)
// Registers: warps are allocated in fours, so 18 warps cost as much as 20 and the block gets 96 registers per thread.  Moving the issue warps' surplus to
// the slot warps with setmaxnreg (104 / 112 registers) was measured neutral to slower: ptxas does not turn the extra registers into deeper prefetch.
constexpr uint32_t FF_COL_W = 256;               // TMEM: weights behind the accumulators: + 64 job, job = 0 tile 1 hi, 1 tile 1 lo, 2 tile 2 hi, 3 tile 2 lo

struct FfParams {
  const float2* ln_stats;                        // (M) (mean, rstd)
  const float* bias2; const float* osc;          // folded bias b + W beta (161), 1 / (S_a S_w)
  const float* conv_w; const float* conv_b; int kconv;
  const float* dt_bias; const float* A_log; const float* D;
  float* y; int64_t ldy;
  int64_t B, T, M; int nchunks; int zero; int* ovf_flag;
  long long* trace;                              // FF_TRACE builds: per-role time stamps of CTA 0 (tools/front_trace.py)
  int stats_bulk;                                // the statistics of a chunk can be bulk-copied (16-byte aligned source: T even, aligned base)
};

__device__ __forceinline__ void ff_tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ff_tmem_ld_32x16(uint32_t taddr, float (&v)[8]) {      // FF_H = 8 builds
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ff_tmem_ld_32x16(uint32_t taddr, float (&v)[4]) {      // FF_H = 4 builds
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ff_tmem_st_32x32u(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
         "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void ff_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 ff_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ff_lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ff_sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ff_silu(float z) { return z * sigmoid_fast_f(z); }
__device__ __forceinline__ void ff_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

#ifdef FF_TRACE
#define FF_TR(role, item, k) do { if (blockIdx.x == 0 && (item) < 256 && p.trace) p.trace[((role) * 256 + (item)) * 4 + (k)] = clock64(); } while (0)
#else
#define FF_TR(role, item, k) do { } while (0)
#endif
// sequences of slot s of this CTA: b = blockIdx.x + gridDim.x * (FF_SLOTS * k + s), k = 0, 1, ...
__device__ __forceinline__ int ff_nseq(int64_t B, int s) {
  const int64_t first = (int64_t)blockIdx.x + (int64_t)gridDim.x * s;
  if (first >= B) return 0;
  const int64_t stride = (int64_t)gridDim.x * FF_SLOTS;
  return (int)((B - first + stride - 1) / stride);
}

// DENSE: y rows are FF_P floats apart (the pass's own y buffer): every store address is the chunk pointer plus an immediate
template <bool DENSE>
__global__ void __launch_bounds__(FF_THREADS, 1)
mamba_front_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapWh, const __grid_constant__ CUtensorMap tmapWl,
                   const FfParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (base - smem_u32(smem_raw));                       // generic pointer to the aligned base
  const uint32_t bars = base + FF_BARS;
  const uint32_t bar_w = bars;
  auto bar_raw = [&](int s) { return bars + 8u * (1 + s); };                        // TMA landed the raw chunk
  auto bar_op = [&](int s) { return bars + 8u * (1 + FF_NST + s); };                // the slot's threads wrote the operand
  auto bar_free = [&](int s) { return bars + 8u * (1 + 2 * FF_NST + s); };          // MMAs that read the stage retired
  constexpr int B0 = 1 + 3 * FF_NST;
  auto bar_dfull = [&](int sl) { return bars + 8u * (B0 + sl); };                   // both accumulator tiles of the slot's chunk are complete
  auto bar_dempty = [&](int sl) { return bars + 8u * (B0 + 4 + sl); };              // the slot's threads have pulled them into registers
  const uint32_t tmem_slot = bars + 8u * (B0 + 8);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = p.nchunks;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < FF_NST; ++s) { mbar_init(bar_raw(s), 1); mbar_init(bar_op(s), 128); mbar_init(bar_free(s), 1); }
    for (int sl = 0; sl < FF_SLOTS; ++sl) { mbar_init(bar_dfull(sl), 1); mbar_init(bar_dempty(sl), 128); }
    fence_barrier_init();
  }
  if (warp == FF_MMA_WARP) tmem_alloc(tmem_slot, 512);
  if (warp == FF_TMA_WARP && lane == 0) { tma_prefetch_desc(&tmapX); tma_prefetch_desc(&tmapWh); tma_prefetch_desc(&tmapWl); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // ---- weights: global -> ring (staging) -> TMEM, once per CTA -------------------------------------------------------------------------------------------
  // staging: tile 1 hi at 0, tile 1 lo at 32 KB: [K chunk of 64][128 rows][128 B]; tile 2 (W rows 128-191: B 16, C 16, dt, zero rows; 176+ are out of
  // bounds and arrive as zeros) at 64 KB: per K chunk (16 KB apart) [hi 64 rows][lo 64 rows]
  if (warp == FF_TMA_WARP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, 2u * 49152u);
      for (int kch = 0; kch < 2; ++kch) {
        tma_load_2d(&tmapWh, bar_w, base + FF_RING + kch * 16384, kch * 64, 0);
        tma_load_2d(&tmapWh, bar_w, base + FF_RING + kch * 16384 + 8192, kch * 64, 64);
        tma_load_2d(&tmapWl, bar_w, base + FF_RING + 32768 + kch * 16384, kch * 64, 0);
        tma_load_2d(&tmapWl, bar_w, base + FF_RING + 32768 + kch * 16384 + 8192, kch * 64, 64);
        tma_load_2d(&tmapWh, bar_w, base + FF_RING + 65536 + kch * 16384, kch * 64, 128);
        tma_load_2d(&tmapWl, bar_w, base + FF_RING + 65536 + kch * 16384 + 8192, kch * 64, 128);
      }
    }
    __syncwarp();
  }
  if (warp < FF_SCAN_WARPS) {
    const int job = warp >> 2, quarter = warp & 3;                                  // job: 0 tile 1 hi, 1 tile 1 lo, 2 tile 2 hi, 3 tile 2 lo
    mbar_wait(bar_w, 0);
    // tile 1: TMEM lane = weight row.  tile 2: quarters 0-2 each take the [B | C] rows (staging rows 0-31), quarter 3 the dt row and the zero rows (32-63)
    const int r = job < 2 ? quarter * 32 + lane : (quarter < 3 ? lane : 32 + lane);
    const uint32_t src0 = base + FF_RING + (job < 2 ? (uint32_t)job * 32768u : 65536u + (job == 3 ? 8192u : 0u)) + (uint32_t)r * 128u;
#pragma unroll
    for (int kch = 0; kch < 2; ++kch) {
      uint32_t w[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                                                 // logical 16-byte slot q sits at physical slot q ^ (row & 7)
        const uint4 v = ff_lds_u4(src0 + (uint32_t)kch * 16384u + (uint32_t)((q ^ (r & 7)) * 16));
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      ff_tmem_st_32x32u(tmem_base + ((uint32_t)(quarter * 32) << 16) + FF_COL_W + (uint32_t)(job * 64 + kch * 32), w);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();                                                                  // weights are in TMEM: the ring is free for the x chunks
  tc_fence_after();

  int nseq[FF_SLOTS];
#pragma unroll
  for (int s = 0; s < FF_SLOTS; ++s) nseq[s] = ff_nseq(p.B, s);
  const int nsteps = nseq[0] * nchunks;                              // slot 0 never has fewer sequences than the others: the active slots of a step are a prefix

  if (warp < FF_SCAN_WARPS) {
    // ===================================== slot warps: slot = warp / 4, quarter = warp % 4, thread = channel ======================================
    const int slot = warp >> 2, quarter = warp & 3;
    const int ch = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t d_x = tmem_base + lane_sel + (uint32_t)(slot * 64), d_bc = d_x + 32u;
    const float osc = __ldg(p.osc);
    const float Dh = p.D ? __ldg(p.D) : 0.f;
    const int kconv = p.kconv;
    // conv over z = osc acc + bias_x (the projection's scale and folded bias): sum_k w_k z_k + b = sum_k (w_k osc) acc_k + (b + bias_x sum_k w_k),
    // so the taps carry osc, the bias carries bias_x, and the recurrence works on the raw accumulator values (one FFMA per token less)
    float cw[4], cb, hpad;
    {
      const float bias_x = __ldg(p.bias2 + ch);
      hpad = -bias_x / osc;                                          // accumulator value whose z is the conv's zero padding: osc hpad + bias_x = 0
      float wsum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float w = (j >= 4 - kconv) ? __ldg(p.conv_w + (size_t)ch * kconv + j - (4 - kconv)) : 0.f; wsum += w; cw[j] = w * osc; }
      cb = fmaf(bias_x, wsum, __ldg(p.conv_b + ch));
    }
    // prepare role: quarters 0-2: conv channel 128 + lane (B_0..15, C_0..15); quarter 3: the dt row
    float bw[4], bbias, bias_p, Ah = 0.f, dtb = 0.f;
    {
      const int bch = FF_P + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) bw[j] = (j >= 4 - kconv) ? __ldg(p.conv_w + (size_t)bch * kconv + j - (4 - kconv)) : 0.f;
      bbias = __ldg(p.conv_b + bch);
      bias_p = quarter < 3 ? __ldg(p.bias2 + bch) : __ldg(p.bias2 + FF_P + 2 * FF_N);
      if (quarter == 3) { Ah = -expf(__ldg(p.A_log)); dtb = __ldg(p.dt_bias); }
    }
    const int64_t ldy = DENSE ? (int64_t)FF_P : p.ldy;
    // -DFF_FMA2: the state row as fp32 pairs, the recurrence on fma.rn.f32x2 (FFMA2: 16 instead of 32 issue slots per channel-token, bit-identical results).
    // Measured 8 % SLOWER (0.965 against 0.893 ms per C2 layer, profiles/r2_front_ffma2_ab.txt): an FFMA2 holds the FMA pipe for two passes, so the pipe -- not
    // the issue slot -- is what the recurrence saturates, and the 64-bit register pairs cost ptxas its scheduling freedom.  Kept as an opt-in build.
#ifndef FF_FMA2
    float s[FF_N];
#else
    uint64_t s2[FF_N / 2];
#endif
    float h1 = hpad, h2 = hpad, h3 = hpad;                           // raw x accumulators of the tokens t-1, t-2, t-3
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;                              // quarter 0: raw B / C values of the previous chunk's last three tokens

    // ---- convert one landed x chunk in place: thread = (token r = lane, 32-column box kc = quarter) --------------------------------------------------------
    // Each thread rewrites the 128 bytes it read -- 32 fp32 of row r of box kc -- as [fp16 hi of those 32 values (64 B) | fp16 lo (64 B)] in the same row:
    // the box stays a [32 tokens][128 B] SWIZZLE_128B K-major tile whose first two 32-byte K steps are the hi operand and whose last two are the lo operand
    // of K columns [32 kc, 32 kc + 32).  No thread touches another thread's bytes, so no barrier is needed between the read and the write.
    float amax = 0.f;
    auto convert = [&](int item, int64_t row0) {
      const int st = item % FF_NST;
      const uint32_t ph = (uint32_t)(item / FF_NST) & 1u;
      const int64_t m = row0 + lane;
      float2 stt = make_float2(0.f, 0.f);
      if (!p.stats_bulk && m < p.M) stt = __ldg(p.ln_stats + m);
      const uint32_t row = base + FF_RING + st * FF_STAGE_BYTES + (uint32_t)quarter * 4096u + (uint32_t)lane * 128u;
      const int sw = lane & 7;
      mbar_wait(bar_raw(st), ph);
      if (p.stats_bulk && m < p.M) stt = reinterpret_cast<const float2*>(sm + FF_STATS + st * (FF_Q * 8))[lane];
      stt.x = -stt.x * stt.y;                                        // (a - mu) rstd = fma(a, rstd, -mu rstd), as the converter of gemm_tc_ts_kernel
      stt.x *= FF_SA; stt.y *= FF_SA;
#ifndef FF_ABL_CONV
      float a[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                                  // logical 16-byte slot q sits at physical slot q ^ (row & 7)
        const float4 v = ff_lds_f4(row + (uint32_t)((q ^ sw) * 16));
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
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ff_sts_u4(row + (uint32_t)((j ^ sw) * 16), hi2[4 * j], hi2[4 * j + 1], hi2[4 * j + 2], hi2[4 * j + 3]);
        ff_sts_u4(row + (uint32_t)(((4 + j) ^ sw) * 16), lo2[4 * j], lo2[4 * j + 1], lo2[4 * j + 2], lo2[4 * j + 3]);
      }
#endif
      fence_proxy_async();                                           // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_op(st));
    };

    // ---- B / C rows of tokens [T0, T1) of the chunk: conv + SiLU along this lane's tile 2 columns -> bc[token][lane] -----------------------------------------
    auto prep_bc = [&](auto t0_tag, auto t1_tag, const float (&v)[32], float* __restrict__ dst) {
      constexpr int T0 = decltype(t0_tag)::value, T1 = decltype(t1_tag)::value;
      float raw[T1 - T0 + 3];
#pragma unroll
      for (int j = 0; j < T1 - T0 + 3; ++j) {
        const int t = T0 - 3 + j;
        raw[j] = t >= 0 ? fmaf(v[t >= 0 ? t : 0], osc, bias_p) : (t == -1 ? g1 : (t == -2 ? g2 : g3));
      }
#pragma unroll
      for (int j = 0; j < T1 - T0; ++j) {
        const float o = fmaf(bw[3], raw[j + 3], fmaf(bw[2], raw[j + 2], fmaf(bw[1], raw[j + 1], fmaf(bw[0], raw[j], bbias))));
        dst[(T0 + j) * 2 * FF_N] = ff_silu(o);
      }
    };

    // one half chunk of FF_H tokens.  RESC: recurrence on r_t = S_t / E_t (w = x dt / E_t; y = E_t (C . r) + D x); else the direct form.
    // Per group of 8 tokens: phase A (independent of the state) conv + SiLU and the per-token scalars, phase B the state recurrence.
    auto half = [&](auto resc_tag, auto guard_tag, const float (&xr)[FF_H], const float4* __restrict__ bc4, const float4* __restrict__ pl, float* yp, int nvalid) {
      constexpr bool RESC = decltype(resc_tag)::value, GUARD = decltype(guard_tag)::value;
#pragma unroll
      for (int g = 0; g < FF_H / FF_G; ++g) {
        float xv[FF_G], w0[FF_G], ez[FF_G];
#pragma unroll
        for (int q = 0; q < FF_G / 4; ++q) {
          const float4 wq = pl[(RESC ? 3 : 0) * (FF_Q / 4) + (FF_G / 4) * g + q];   // RESC: dt / E_t; direct: dt
          const float4 eq = pl[(RESC ? 2 : 1) * (FF_Q / 4) + (FF_G / 4) * g + q];   // RESC: E_t;      direct: decay
          w0[4 * q] = wq.x; w0[4 * q + 1] = wq.y; w0[4 * q + 2] = wq.z; w0[4 * q + 3] = wq.w;
          ez[4 * q] = eq.x; ez[4 * q + 1] = eq.y; ez[4 * q + 2] = eq.z; ez[4 * q + 3] = eq.w;
        }
#pragma unroll
        for (int jj = 0; jj < FF_G; ++jj) {
          const int j = FF_G * g + jj;
          const float xm1 = j >= 1 ? xr[j >= 1 ? j - 1 : 0] : h1, xm2 = j >= 2 ? xr[j >= 2 ? j - 2 : 0] : (j == 1 ? h1 : h2),
                      xm3 = j >= 3 ? xr[j >= 3 ? j - 3 : 0] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
          const float c = fmaf(cw[3], xr[j], fmaf(cw[2], xm1, fmaf(cw[1], xm2, fmaf(cw[0], xm3, cb))));
          xv[jj] = ff_silu(c);
          w0[jj] *= xv[jj];
        }
#pragma unroll
        for (int jj = 0; jj < FF_G; ++jj) {
          const int j = FF_G * g + jj;
#ifndef FF_FMA2
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int q = 0; q < FF_N / 4; ++q) {
            const float4 bv = bc4[j * 8 + q], cv = bc4[j * 8 + 4 + q];
            if (RESC) {
              s[4 * q + 0] = fmaf(w0[jj], bv.x, s[4 * q + 0]); s[4 * q + 1] = fmaf(w0[jj], bv.y, s[4 * q + 1]);
              s[4 * q + 2] = fmaf(w0[jj], bv.z, s[4 * q + 2]); s[4 * q + 3] = fmaf(w0[jj], bv.w, s[4 * q + 3]);
            } else {
              s[4 * q + 0] = fmaf(ez[jj], s[4 * q + 0], w0[jj] * bv.x); s[4 * q + 1] = fmaf(ez[jj], s[4 * q + 1], w0[jj] * bv.y);
              s[4 * q + 2] = fmaf(ez[jj], s[4 * q + 2], w0[jj] * bv.z); s[4 * q + 3] = fmaf(ez[jj], s[4 * q + 3], w0[jj] * bv.w);
            }
            a0 = fmaf(cv.x, s[4 * q + 0], a0); a1 = fmaf(cv.y, s[4 * q + 1], a1);       // two chains: 4 measured 1.6 % slower (2 registers, 2 FADDs more)
            a0 = fmaf(cv.z, s[4 * q + 2], a0); a1 = fmaf(cv.w, s[4 * q + 3], a1);
          }
#else
          // packed form: (s_2i, s_2i+1) pairs, w0 / ez as broadcast operands; the pair accumulator is the scalar form's two chains a0 (even n), a1 (odd n)
          const uint64_t w2 = pk2(w0[jj], w0[jj]), ez2 = pk2(ez[jj], ez[jj]);
          const ulonglong2* bc2 = reinterpret_cast<const ulonglong2*>(bc4);
          uint64_t acc = 0ull;
#pragma unroll
          for (int q = 0; q < FF_N / 4; ++q) {
            const ulonglong2 bv = bc2[j * 8 + q], cv = bc2[j * 8 + 4 + q];
            if (RESC) {
              s2[2 * q] = fma2(w2, bv.x, s2[2 * q]); s2[2 * q + 1] = fma2(w2, bv.y, s2[2 * q + 1]);
            } else {
              s2[2 * q] = fma2(ez2, s2[2 * q], mul2(w2, bv.x)); s2[2 * q + 1] = fma2(ez2, s2[2 * q + 1], mul2(w2, bv.y));
            }
            acc = fma2(cv.x, s2[2 * q], acc); acc = fma2(cv.y, s2[2 * q + 1], acc);
          }
          float a0, a1;
          upk2(acc, a0, a1);
#endif
          const float dot = a0 + a1;
          const float yv = fmaf(Dh, xv[jj], RESC ? ez[jj] * dot : dot);
          if (!GUARD || j < nvalid) yp[(int64_t)j * ldy] = yv;
        }
      }
      h3 = xr[FF_H - 3]; h2 = xr[FF_H - 2]; h1 = xr[FF_H - 1];
    };

    const int my_steps = nseq[slot] * nchunks;
    int steps_of[FF_SLOTS];
#pragma unroll
    for (int s2 = 0; s2 < FF_SLOTS; ++s2) steps_of[s2] = nseq[s2] * nchunks;
    auto nact = [&](int step) { int n = 0;
#pragma unroll
      for (int s2 = 0; s2 < FF_SLOTS; ++s2) n += (step < steps_of[s2]) ? 1 : 0;
      return n; };
    const int64_t seq_stride = (int64_t)gridDim.x * FF_SLOTS * p.T;   // rows between consecutive sequences of this slot
    int64_t row_cur = ((int64_t)blockIdx.x + (int64_t)gridDim.x * slot) * p.T;        // first row of chunk i
    int c = 0;                                                       // chunk of the sequence
    auto advance = [&](int64_t row, int cc, int64_t& row_n, int& c_n) {                  // (row, chunk) of the next step
      if (cc + 1 < nchunks) { row_n = row + FF_Q; c_n = cc + 1; } else { row_n = row - (int64_t)cc * FF_Q + seq_stride; c_n = 0; }
    };
    int item_next = slot;                                            // ring item of this slot's chunk i + 1 (items of a step: its active slots in order)
    if (my_steps > 0) { convert(item_next, row_cur); item_next += nact(0); }
    for (int i = 0; i < my_steps; ++i) {
      int64_t row_nx; int c_nx;
      advance(row_cur, c, row_nx, c_nx);
      const int tc = (int)min((int64_t)FF_Q, p.T - (int64_t)c * FF_Q);
      const int sb = 2 * slot + (i & 1);
      if (quarter == 0 && lane == 0) FF_TR(4 + slot, i, 0);
      if (i + 1 < my_steps) { convert(item_next, row_nx); item_next += nact(i + 1); }
      if (c == 0) {
#ifndef FF_FMA2
#pragma unroll
        for (int n = 0; n < FF_N; ++n) s[n] = 0.f;
#else
#pragma unroll
        for (int n = 0; n < FF_N / 2; ++n) s2[n] = 0ull;
#endif
        h1 = h2 = h3 = hpad; g1 = g2 = g3 = 0.f;
      }
      if (quarter == 0 && lane == 0) FF_TR(4 + slot, i, 1);
      mbar_wait(bar_dfull(slot), (uint32_t)i & 1u);
      tc_fence_after();
      if (quarter == 0 && lane == 0) FF_TR(4 + slot, i, 2);
      // ---- prepare: tile 2 of this chunk -> B / C rows and dt planes in shared memory (buffer i & 1) ----
      {
        float v[32];
        tmem_ld_32x32(d_bc, v);
        float* bcw = reinterpret_cast<float*>(sm + FF_BC + (uint32_t)sb * FF_BC_BYTES) + lane;
#ifndef FF_ABL_PREP
        if (quarter == 0) {
          prep_bc(std::integral_constant<int, 0>{}, std::integral_constant<int, 11>{}, v, bcw);
          g1 = fmaf(v[31], osc, bias_p); g2 = fmaf(v[30], osc, bias_p); g3 = fmaf(v[29], osc, bias_p);
        } else if (quarter == 1) {
          prep_bc(std::integral_constant<int, 11>{}, std::integral_constant<int, 22>{}, v, bcw);
        } else if (quarter == 2) {
          prep_bc(std::integral_constant<int, 22>{}, std::integral_constant<int, 32>{}, v, bcw);
        } else {
          float* const dtraw = reinterpret_cast<float*>(sm + FF_DTRAW) + slot * FF_Q;
          if (lane == 0) {                                           // lane 0 of quarter 3 = TMEM lane 96 = the dt row
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(dtraw)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          __syncwarp();
          const float z = fmaf(dtraw[lane], osc, bias_p);
          __syncwarp();
          const float d = (lane < tc) ? softplus_f(z + dtb) : 0.f;
          const float dec = (lane < tc) ? expf(d * Ah) : 1.f;
          float E = dec;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, E, o); if (lane >= o) E *= u; }   // inclusive product scan
          const float Emin = __shfl_sync(0xffffffffu, E, 31);        // decays <= 1: the last product is the smallest
          if (lane == 0) reinterpret_cast<int*>(sm + FF_FLAGS)[sb] = (Emin < 0x1p-60f || !(Emin == Emin)) ? 1 : 0;
          float* pl = reinterpret_cast<float*>(sm + FF_DD + (uint32_t)sb * FF_DD_BYTES) + lane;
          pl[0] = d; pl[FF_Q] = dec; pl[2 * FF_Q] = E; pl[3 * FF_Q] = d * (1.f / E);
        }
#endif
      }
      ff_bar_sync(1 + slot, 128);                                    // the chunk's rows are complete; every warp of the slot has left chunk i - 1
      const bool direct = reinterpret_cast<const int*>(sm + FF_FLAGS)[sb] != 0;
      const float4* bc4 = reinterpret_cast<const float4*>(sm + FF_BC + (uint32_t)sb * FF_BC_BYTES);
      const float4* pl4 = reinterpret_cast<const float4*>(sm + FF_DD + (uint32_t)sb * FF_DD_BYTES);
      float* yp = p.y + row_cur * ldy + ch;
#pragma unroll 1
      for (int hf = 0; hf < FF_Q / FF_H; ++hf) {
        float xr[FF_H];
        ff_tmem_ld_32x16(d_x + (uint32_t)(hf * FF_H), xr);
        if (hf == FF_Q / FF_H - 1) { tc_fence_before(); mbar_arrive(bar_dempty(slot)); }    // the tensor core may overwrite the slot's accumulators
        const float4* bch = bc4 + hf * FF_H * 8;
        const float4* plh = pl4 + hf * (FF_H / 4);
        float* yph = yp + (int64_t)(hf * FF_H) * ldy;
#ifdef FF_ABL_SCAN
        if (xr[0] == 123.456f) yph[0] = xr[1];
        continue;
#endif
        if (tc == FF_Q) {
          if (!direct) half(std::true_type{}, std::false_type{}, xr, bch, plh, yph, FF_H);
          else half(std::false_type{}, std::false_type{}, xr, bch, plh, yph, FF_H);
        } else {
          if (!direct) half(std::true_type{}, std::true_type{}, xr, bch, plh, yph, tc - hf * FF_H);
          else half(std::false_type{}, std::true_type{}, xr, bch, plh, yph, tc - hf * FF_H);
        }
      }
      if (!direct) {                                                 // back to the true state: S = r E
        const float Eend = reinterpret_cast<const float*>(pl4)[2 * FF_Q + FF_Q - 1];
#ifndef FF_FMA2
#pragma unroll
        for (int n = 0; n < FF_N; ++n) s[n] *= Eend;
#else
        const uint64_t Eend2 = pk2(Eend, Eend);
#pragma unroll
        for (int n = 0; n < FF_N / 2; ++n) s2[n] = mul2(s2[n], Eend2);
#endif
      }
      if (quarter == 0 && lane == 0) FF_TR(4 + slot, i, 3);
      row_cur = row_nx; c = c_nx;
    }
    if (!(amax <= 65504.f)) atomicOr(p.ovf_flag, 1);
  } else if (warp == FF_TMA_WARP) {
    // ===================================== TMA producer ======================================
    if (elect_one()) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < nsteps; ++i) {
        const int k = i / nchunks, c = i - k * nchunks;
#pragma unroll
        for (int s = 0; s < FF_SLOTS; ++s) {
          if (k >= nseq[s]) continue;
          const int64_t b = (int64_t)blockIdx.x + (int64_t)gridDim.x * (FF_SLOTS * k + s);
          const int row0 = (int)(b * p.T + (int64_t)c * FF_Q);
          FF_TR(0, i * 4 + s, 0);
          mbar_wait_one(bar_free(st), ph ^ 1);
          FF_TR(0, i * 4 + s, 1);
          int nst_rows = 0;
          if (p.stats_bulk) { const int64_t left = p.M - row0; nst_rows = left < FF_Q ? (int)left : FF_Q; }      // even: T and FF_Q are
          mbar_arrive_expect_tx(bar_raw(st), FF_STAGE_BYTES + (uint32_t)nst_rows * 8u);
          if (nst_rows > 0) ff_bulk_load(base + FF_STATS + st * (FF_Q * 8), p.ln_stats + row0, (uint32_t)nst_rows * 8u, bar_raw(st));
          const uint32_t dst = base + FF_RING + st * FF_STAGE_BYTES;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) tma_load_2d(&tmapX, bar_raw(st), dst + kc * 4096, kc * 32, row0);
          if (++st == FF_NST) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == FF_MMA_WARP) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, FF_Q);
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < nsteps; ++i) {
        const int k = i / nchunks;
#pragma unroll
        for (int s = 0; s < FF_SLOTS; ++s) {
          if (k >= nseq[s]) continue;
          FF_TR(2, i * 4 + s, 0);
          mbar_wait_one(bar_op(st), ph);
          FF_TR(2, i * 4 + s, 1);
          mbar_wait_one(bar_dempty(s), ((uint32_t)i & 1u) ^ 1u);     // chunk i - 1 of this slot is in its threads' registers
          FF_TR(2, i * 4 + s, 2);
          tc_fence_after();
          const uint32_t stage = base + FF_RING + st * FF_STAGE_BYTES;
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const uint32_t d = tmem_base + (uint32_t)(s * 64 + tile * 32);
            const uint32_t a_hi = tmem_base + FF_COL_W + (uint32_t)(tile * 128), a_lo = a_hi + 64u;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                         // K step of 16: weights 8 TMEM columns; activations [box of 32 columns][hi 64 B | lo 64 B]
              const uint64_t bh = umma_desc_k_sw128(stage + (uint32_t)(ks >> 1) * 4096u) + 2u * (ks & 1), bl = bh + 4u;
#ifndef FF_ABL_MMA                                                    // ablation builds (tools/ablate_front.sh): timing only, wrong results
              umma_f16_ts(d, a_hi + 8u * ks, bh, idesc, ks > 0 ? 1u : 0u);
              umma_f16_ts(d, a_lo + 8u * ks, bh, idesc, 1u);
              umma_f16_ts(d, a_hi + 8u * ks, bl, idesc, 1u);
#endif
            }
          }
          umma_commit(bar_free(st));
          umma_commit(bar_dfull(s));
          FF_TR(2, i * 4 + s, 3);
          if (++st == FF_NST) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FF_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
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
#ifdef FF_TRACE
  { static long long* tr = nullptr; if (!tr) { cudaMalloc(&tr, 8 * 256 * 4 * 8); } cudaMemsetAsync(tr, 0, 8 * 256 * 4 * 8, st); p.trace = tr;
    if (const char* e = getenv("FF_TRACE_PTR_FILE")) { FILE* f = fopen(e, "w"); if (f) { fprintf(f, "%llu\n", (unsigned long long)(uintptr_t)tr); fclose(f); } } }
#endif
  p.stats_bulk = (T % 2 == 0 && ((uintptr_t)ln_stats & 15) == 0) ? 1 : 0;
  p.ovf_flag = tc_overflow_flag();
  if (!p.ovf_flag) { set_error("mamba_front_fused: cannot resolve the overflow flag"); return EIGB200_ECUDA; }
  const int64_t grid = B < (int64_t)num_sms() ? B : (int64_t)num_sms();
  const size_t smem = (size_t)FF_SMEM + 1024 /*alignment*/;
  if (ldy == FF_P) {
    EIGB_CUDA(cudaFuncSetAttribute(mamba_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mamba_front_kernel<true><<<(unsigned)grid, FF_THREADS, smem, st>>>(tX, tWh, tWl, p);
  } else {
    EIGB_CUDA(cudaFuncSetAttribute(mamba_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mamba_front_kernel<false><<<(unsigned)grid, FF_THREADS, smem, st>>>(tX, tWh, tWl, p);
  }
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
