// k2_ssd_tc.cu -- K2b on the 5th-generation tensor cores: the SSD selective scan in its chunked ("state-space dual") form, fused with the depthwise
// causal conv + SiLU and softplus(dt) that precede it in SSD.forward.
//
// Reference operators replaced (same as k2_ssd_scan.cu): mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=D), models/mamba.py:138-150, and
// conv1d + SiLU + softplus(dt + dt_bias) + A = -exp(A_log), models/mamba.py:119-133.  Recurrence per (batch b, head h):
//   S_t[p,n] = exp(dt_t A) S_{t-1}[p,n] + (dt_t x_t[p]) B_t[n],      y_t[p] = sum_n C_t[n] S_t[p,n] + D x_t[p].
//
// Why a second form: the recurrent kernel (ssd_scan_v3) spends 2 FMA-pipe operations per state element and token -- 4096 per token at P = 128, N = 16 --
// and is issue-bound at 0.47 of the HBM roofline.  Over a chunk of Q = 64 tokens the same recurrence is three small matrix products
//   G    = C B^T                                   (Q x Q,  K = N)        M_ij = G_ij exp(cum_i - cum_j) [j <= i],  cum_i = sum_{j<=i} dt_j A
//   Y^T  = Xd^T M^T + S_prev (C e^{cum})^T         (P x Q,  K = Q + N)    Xd_j = dt_j x_j
//   dS   = Xd^T (B e^{cum_Q - cum})                (P x N,  K = Q)        S_new = e^{cum_Q} S_prev + dS
// with every exponent <= 0 (the "segment sum" form: nothing can overflow), 3xTF32 operands (fp32-level accuracy, as the projections), and the channel axis
// P = 128 on the 128 TMEM lanes, so that the threads that produce an operand (conv + SiLU of a channel along time), the tensor core that consumes it
// (A operand from TMEM) and the threads that finish the result (y of a channel along time) all see "lane = channel".  SIMT work left per chunk: conv + SiLU,
// one exp per element of the Q x Q mask, the operand splits -- about a quarter of the recurrent form's instructions.
//
// One CTA walks whole (sequence, head) items; per chunk the roles hand each other operands through mbarriers:
//   TMA warp       raw [B | C] rows of the chunk (+3 history rows for the 4-tap conv) global -> shared, 3-D tensor map (zero fill outside the sequence)
//   T group (4 w)  thread = token: conv + SiLU of B_t, C_t, softplus(dt), cum (warp scan) -> G operands; after G: mask + decay + split -> M^T (B operand
//                  of the Y product, K-major SWIZZLE_128B), C e^{cum} (its last K chunk), (B e^{cum_Q - cum})^T
//   X group (8 w)  thread = channel: conv + SiLU of x along time (x read straight from global), D x -> y staging tile, dt x -> tf32 hi / lo -> TMEM (tcgen05.st)
//   MMA warp       one thread issues G, dS, Y (tcgen05.mma kind::tf32, 3 per product), tcgen05.commit -> mbarriers
//   Y group (4 w)  thread = channel: S = e^{cum_Q} S + dS in registers (fp32) -> hi / lo -> TMEM operand of the next chunk; Y^T from TMEM + staged D x ->
//                  staging tile -> ONE TMA store per chunk (cp.async.bulk.tensor, clipped at the sequence end by the 3-D tensor map)
// TMEM (512 columns): Xd^T hi|lo 2 x 128, G / Y^T 2 x 64 (G is dead once M is formed: Y accumulates over it), dS 16, S operand 2 x 32.
#include "ssd.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace eigb200 {

constexpr int ST_Q = 64, ST_P = 128, ST_N = 16, ST_HIST = 3;
constexpr int ST_RAWROWS = ST_Q + ST_HIST;                          // 67
constexpr int ST_WARP_X0 = 0, ST_WARP_Y0 = 8, ST_WARP_T0 = 12, ST_WARP_TMA = 16, ST_WARP_MMA = 17;
constexpr int ST_THREADS = 18 * 32;
constexpr uint32_t ST_COL_XD = 0, ST_COL_YG = 256, ST_COL_DS = 384, ST_COL_SOP = 400;
// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t ST_MC_HI = 0, ST_MC_LO = 24576;                  // M^T | C e^cum : 3 K-chunks x [64 rows][128 B]
constexpr uint32_t ST_GA_HI = 49152, ST_GA_LO = 65536;              // [C; C]: 128 rows x 128 B (first 64 B used)
constexpr uint32_t ST_GB_HI = 81920, ST_GB_LO = 90112;              // B: 64 rows x 128 B (first 64 B used)
constexpr uint32_t ST_BW_HI = 98304, ST_BW_LO = 102400;             // (B e^{cum_Q - cum})^T: 2 K-chunks x [16 rows][128 B]
constexpr uint32_t ST_YST = 106496;                                 // 3 x [64 tokens][128 channels] fp32
constexpr uint32_t ST_RAWB = ST_YST + 3 * 32768;                    // 2 x [67][16] fp32, padded to 4352
constexpr uint32_t ST_RAWC = ST_RAWB + 2 * 4352;
constexpr uint32_t ST_DTS = ST_RAWC + 2 * 4352;                     // dt_s[2][64], cum_s hi [2][64] (+512), cum_s lo [2][64] (+1024)
constexpr uint32_t ST_CWBC = ST_DTS + 2048;                         // conv weights of the B / C channels: [32][4] + bias [32]
constexpr uint32_t ST_BARS = ST_CWBC + 1024;
constexpr uint32_t ST_SMEM = ST_BARS + 512;
constexpr uint32_t ST_RAW_BYTES = ST_RAWROWS * ST_N * 4;            // 4288

struct SsdTcParams {
  const float* z; int64_t ldz;                                      // projection buffer (B*T, ldz): x at column h*P, B / C / dt at the offsets below
  int colB, colC, colDt;
  const float* A_log; const float* D; const float* dt_bias; const float* conv_w; const float* conv_b;
  int64_t T; int H, G, kconv; int64_t nitems; int nchunks; int zero;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :: "l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
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
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float a) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(a) : "memory"); }
__device__ __forceinline__ float lds_f1(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float st_silu(float z) { return z * sigmoid_fast_f(z); }
__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u); }

__global__ void __launch_bounds__(ST_THREADS, 1)
ssd_tc_kernel(const __grid_constant__ CUtensorMap tmapZ, const __grid_constant__ CUtensorMap tmapY, const SsdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + ST_BARS;
  auto bar_raw_full = [&](int b) { return bars + 8u * (0 + b); };
  auto bar_raw_free = [&](int b) { return bars + 8u * (2 + b); };
  auto bar_tprep = [&](int b) { return bars + 8u * (4 + b); };
  const uint32_t bar_gop_full = bars + 8u * 6, bar_gop_free = bars + 8u * 7;
  auto bar_g_full = [&](int b) { return bars + 8u * (8 + b); };
  const uint32_t bar_bw_full = bars + 8u * 10, bar_mc_full = bars + 8u * 11, bar_mc_free = bars + 8u * 12;
  auto bar_xd_full = [&](int b) { return bars + 8u * (13 + b); };
  auto bar_xd_free = [&](int b) { return bars + 8u * (15 + b); };
  auto bar_y_full = [&](int b) { return bars + 8u * (17 + b); };
  auto bar_yg_free = [&](int b) { return bars + 8u * (19 + b); };
  const uint32_t bar_ds_full = bars + 8u * 21, bar_sop_full = bars + 8u * 22;
  auto bar_ysfree = [&](int b) { return bars + 8u * (23 + b); };   // 3 staging tiles
  auto bar_ydone = [&](int b) { return bars + 8u * (26 + b); };
  const uint32_t tmem_slot = bars + 8u * 28;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const uint32_t dts = base + ST_DTS;                               // dt_s[b][j] at dts + (b*64 + j)*4, cum_s[b][j] at dts + 512 + (b*64 + j)*4
  const uint32_t cwbc = base + ST_CWBC;                             // w[ch][k] at cwbc + (ch*4 + k)*4, bias[ch] at cwbc + 512 + ch*4   (ch 0-15: B, 16-31: C)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t T = p.T;
  const int nch = p.nchunks;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_raw_full(b), 1); mbar_init(bar_raw_free(b), 128); mbar_init(bar_tprep(b), 128); mbar_init(bar_g_full(b), 1);
      mbar_init(bar_xd_full(b), 256); mbar_init(bar_xd_free(b), 1); mbar_init(bar_y_full(b), 1); mbar_init(bar_yg_free(b), 128); mbar_init(bar_ydone(b), 128);
    }
    for (int b = 0; b < 3; ++b) mbar_init(bar_ysfree(b), 1);
    mbar_init(bar_gop_full, 128); mbar_init(bar_gop_free, 1); mbar_init(bar_bw_full, 128); mbar_init(bar_mc_full, 128); mbar_init(bar_mc_free, 1);
    mbar_init(bar_ds_full, 1); mbar_init(bar_sop_full, 128);
    fence_barrier_init();
  }
  // the unused halves of the 128-byte operand rows (G operands, C e^cum chunk) are never read; zero them once so that no NaN pattern can sit there
  for (uint32_t i = threadIdx.x; i < (ST_YST) / 16; i += ST_THREADS) sts_f4(base + i * 16, 0.f, 0.f, 0.f, 0.f);
  if (warp == ST_WARP_MMA) tmem_alloc(tmem_slot, 512);
  if (warp == ST_WARP_TMA && lane == 0) { tma_prefetch_desc(&tmapZ); tma_prefetch_desc(&tmapY); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == ST_WARP_TMA) {
    // ===================================== TMA producer: raw [B | C] rows ======================================
    if (elect_one()) {
      int64_t cc = 0;
      for (int64_t item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int b = (int)(item / p.H), h = (int)(item - (int64_t)b * p.H);
        const int g = h / (p.H / p.G);
        for (int c = 0; c < nch; ++c, ++cc) {
          const int buf = (int)(cc & 1); const uint32_t u = (uint32_t)(cc >> 1);
          mbar_wait_one(bar_raw_free(buf), (u & 1) ^ 1);
          mbar_arrive_expect_tx(bar_raw_full(buf), 2 * ST_RAW_BYTES);
          tma_load_3d(&tmapZ, bar_raw_full(buf), base + ST_RAWB + buf * 4352, p.colB + g * ST_N, c * ST_Q - ST_HIST, b);
          tma_load_3d(&tmapZ, bar_raw_full(buf), base + ST_RAWC + buf * 4352, p.colC + g * ST_N, c * ST_Q - ST_HIST, b);
        }
      }
    }
    __syncwarp();
  } else if (warp >= ST_WARP_T0 && warp < ST_WARP_T0 + 4) {
    // ===================================== T group: thread = token ======================================
    // Software-pipelined over the CTA's chunk sequence w = 0, 1, ... (all chunks of its items, in order): iteration w first PREPARES chunk w + 1 (conv, dt, cum,
    // operands of G) so that the tensor core computes G(w + 1) while this group forms M(w) from G(w) -- the G round trip is off the critical path.
    const int wq = warp - ST_WARP_T0;                               // TMEM lane quarter of this warp
    const int tok = 32 * (wq & 1) + lane;                           // token of the chunk this thread owns
    const int hb = wq >> 1;                                         // 0: B channels / columns 0-31 of G, 1: C channels / columns 32-63
    const int tid_t = threadIdx.x - ST_WARP_T0 * 32;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const int sw = tok & 7;
    const int64_t my_items = (p.nitems - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x;
    const int64_t nwork = my_items * nch;
    float vc[16], vn[16];                                           // conv + SiLU of this token's B or C channels: current chunk, next chunk
    float cum_c = 0.f, cuml_c = 0.f, cumQ_c = 0.f, cumQl_c = 0.f, cum_n = 0.f, cuml_n = 0.f, cumQ_n = 0.f, cumQl_n = 0.f;
    float dtr0 = 0.f, dtr1 = 0.f, Ah = 0.f, dtb = 0.f;
    const float* dtp = p.z;
#pragma unroll
    for (int i = 0; i < 16; ++i) { vc[i] = 0.f; vn[i] = 0.f; }
    for (int64_t w = -1; w < nwork; ++w) {
      if (w + 1 < nwork) {
        // ================= prepare chunk wn = w + 1 =================
        const int64_t wn = w + 1;
        const int64_t k = wn / nch;
        const int c = (int)(wn - k * nch);
        const int64_t item = (int64_t)blockIdx.x + k * (int64_t)gridDim.x;
        const int b = (int)(item / p.H), h = (int)(item - (int64_t)b * p.H);
        const int buf = (int)(wn & 1); const uint32_t u = (uint32_t)(wn >> 1);
        const int64_t t0 = (int64_t)c * ST_Q;
        if (c == 0) {                                               // a new (sequence, head): scalars, conv weights of its B / C group, first dt rows
          const int g = h / (p.H / p.G);
          Ah = -expf(p.A_log[h]); dtb = p.dt_bias[h];
          dtp = p.z + ((size_t)b * T) * p.ldz + p.colDt + h;
          named_bar_sync(2, 128);                                   // every T thread is done with the previous item's conv weights
          if (tid_t < 32) {
            const int ch = (tid_t < 16 ? p.colB : p.colC) + g * ST_N + (tid_t & 15);   // column of the projection buffer == conv channel index
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) sts_f1(cwbc + (tid_t * 4 + kk) * 4, (kk >= 4 - p.kconv) ? p.conv_w[(size_t)ch * p.kconv + kk - (4 - p.kconv)] : 0.f);
            sts_f1(cwbc + 512 + tid_t * 4, p.conv_b[ch]);
          }
          named_bar_sync(2, 128);
          dtr0 = (lane < T) ? __ldg(dtp + (size_t)lane * p.ldz) : 0.f;
          dtr1 = (32 + lane < T) ? __ldg(dtp + (size_t)(32 + lane) * p.ldz) : 0.f;
        }
        const float r0 = dtr0, r1 = dtr1;
        if (c + 1 < nch) {                                          // prefetch the raw dt of the chunk after this one
          const int64_t tn = t0 + ST_Q;
          dtr0 = (tn + lane < T) ? __ldg(dtp + (size_t)(tn + lane) * p.ldz) : 0.f;
          dtr1 = (tn + 32 + lane < T) ? __ldg(dtp + (size_t)(tn + 32 + lane) * p.ldz) : 0.f;
        }
        // ---- dt, cum: every T warp scans all 64 tokens (lane: tokens lane and lane + 32) ----
        const float d0 = (t0 + lane < T) ? softplus_f(r0 + dtb) : 0.f;
        const float d1 = (t0 + 32 + lane < T) ? softplus_f(r1 + dtb) : 0.f;
        // The mask needs exp(cum_i - cum_j) for NEARBY i, j deep into the chunk: |cum| reaches 30-60 there, so a single float (ulp 2e-6 .. 4e-6) would put
        // that error on weights of order one.  The scan runs in double; cum is kept as a float pair (hi, lo) and differences are formed pairwise.
        double c0d = (double)(d0 * Ah), c1d = (double)(d1 * Ah);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double v0 = __shfl_up_sync(0xffffffffu, c0d, o), v1 = __shfl_up_sync(0xffffffffu, c1d, o);
          if (lane >= o) { c0d += v0; c1d += v1; }
        }
        c1d += __shfl_sync(0xffffffffu, c0d, 31);
        const double cumQd = __shfl_sync(0xffffffffu, c1d, 31);
        const float c0 = (float)c0d, c1 = (float)c1d, c0l = (float)(c0d - (double)c0), c1l = (float)(c1d - (double)c1);
        cumQ_n = (float)cumQd; cumQl_n = (float)(cumQd - (double)cumQ_n);
        cum_n = (wq & 1) ? c1 : c0; cuml_n = (wq & 1) ? c1l : c0l;
        if (wn >= 2) {                                              // dt_s / cum_s[buf] of chunk wn - 2 are no longer read (X group, Y group)
          mbar_wait(bar_xd_full(buf), (u - 1) & 1);
          mbar_wait(bar_ydone(buf), (u - 1) & 1);
        }
        if (wq == 0) {
          sts_f1(dts + (buf * 64 + lane) * 4, d0); sts_f1(dts + (buf * 64 + 32 + lane) * 4, d1);
          sts_f1(dts + 512 + (buf * 64 + lane) * 4, c0); sts_f1(dts + 512 + (buf * 64 + 32 + lane) * 4, c1);
          sts_f1(dts + 1024 + (buf * 64 + lane) * 4, c0l); sts_f1(dts + 1024 + (buf * 64 + 32 + lane) * 4, c1l);
        }
        // ---- conv + SiLU of this token's 16 B (hb = 0) or C (hb = 1) channels ----
        mbar_wait(bar_raw_full(buf), u & 1);
        const uint32_t raw = base + (hb ? ST_RAWC : ST_RAWB) + buf * 4352 + tok * (ST_N * 4);   // row tok = token t0 + tok - 3: the first tap
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 x0 = lds_f4(raw + q * 16), x1 = lds_f4(raw + 64 + q * 16), x2 = lds_f4(raw + 128 + q * 16), x3 = lds_f4(raw + 192 + q * 16);
          const float xs0[4] = {x0.x, x0.y, x0.z, x0.w}, xs1[4] = {x1.x, x1.y, x1.z, x1.w}, xs2[4] = {x2.x, x2.y, x2.z, x2.w}, xs3[4] = {x3.x, x3.y, x3.z, x3.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = hb * 16 + 4 * q + e;
            const float4 wv = lds_f4(cwbc + ch * 16);
            const float bias = lds_f1(cwbc + 512 + ch * 4);
            vn[4 * q + e] = st_silu(fmaf(wv.w, xs3[e], fmaf(wv.z, xs2[e], fmaf(wv.y, xs1[e], fmaf(wv.x, xs0[e], bias)))));
          }
        }
        {
          uint32_t dep = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) dep ^= __float_as_uint(vn[i]);
          mbar_arrive_after(bar_raw_free(buf), dep, (uint32_t)p.zero);    // the raw rows are in registers: TMA may refill the buffer
        }
        mbar_arrive(bar_tprep(buf));
        // ---- operands of G = C B^T:  [C; C] (128 rows) and B (64 rows), 16 K-values = logical 16-byte slots 0-3 of each 128-byte row ----
        if (wn >= 1) mbar_wait(bar_gop_free, (uint32_t)((wn - 1) & 1));   // G(wn - 1) has read the previous operands
        {
          float hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { hi[i] = tf32_hi(vn[i]); lo[i] = tf32_hi(vn[i] - hi[i]); }
          const uint32_t rowo = (uint32_t)tok * 128u;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const uint32_t so = (uint32_t)((s4 ^ sw) << 4);
            if (hb) {
              sts_f4(base + ST_GA_HI + rowo + so, hi[4 * s4], hi[4 * s4 + 1], hi[4 * s4 + 2], hi[4 * s4 + 3]);
              sts_f4(base + ST_GA_HI + 8192 + rowo + so, hi[4 * s4], hi[4 * s4 + 1], hi[4 * s4 + 2], hi[4 * s4 + 3]);
              sts_f4(base + ST_GA_LO + rowo + so, lo[4 * s4], lo[4 * s4 + 1], lo[4 * s4 + 2], lo[4 * s4 + 3]);
              sts_f4(base + ST_GA_LO + 8192 + rowo + so, lo[4 * s4], lo[4 * s4 + 1], lo[4 * s4 + 2], lo[4 * s4 + 3]);
            } else {
              sts_f4(base + ST_GB_HI + rowo + so, hi[4 * s4], hi[4 * s4 + 1], hi[4 * s4 + 2], hi[4 * s4 + 3]);
              sts_f4(base + ST_GB_LO + rowo + so, lo[4 * s4], lo[4 * s4 + 1], lo[4 * s4 + 2], lo[4 * s4 + 3]);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_gop_full);
      }
      if (w >= 0) {
        // ================= finish chunk w: dS / state operands, then G(w) -> M(w) =================
        const int buf = (int)(w & 1); const uint32_t u = (uint32_t)(w >> 1);
        // ---- (B e^{cum_Q - cum})^T for dS (hb = 0) and C e^{cum} for the state part of Y (hb = 1) ----
        mbar_wait(bar_mc_free, (uint32_t)(w & 1) ^ 1u);
        if (hb == 0) {
          const float wdec = expf((cumQ_c - cum_c) + (cumQl_c - cuml_c));
          const uint32_t kc = (uint32_t)(tok >> 5) * 2048u, kk = (uint32_t)(tok & 31);
#pragma unroll
          for (int n = 0; n < 16; ++n) {                            // element (row n, K index tok): slot (kk >> 2) ^ (n & 7), word kk & 3
            const float val = vc[n] * wdec, hi = tf32_hi(val);
            const uint32_t off = kc + (uint32_t)n * 128u + ((((kk >> 2) ^ (uint32_t)(n & 7)) << 4) | ((kk & 3u) << 2));
            sts_f1(base + ST_BW_HI + off, hi);
            sts_f1(base + ST_BW_LO + off, tf32_hi(val - hi));
          }
        } else {
          const float e = expf(cum_c);
          const uint32_t rowo = 2u * 8192u + (uint32_t)tok * 128u;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            float hi[4], lo[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { const float val = vc[4 * s4 + kk] * e; hi[kk] = tf32_hi(val); lo[kk] = tf32_hi(val - hi[kk]); }
            const uint32_t so = (uint32_t)((s4 ^ sw) << 4);
            sts_f4(base + ST_MC_HI + rowo + so, hi[0], hi[1], hi[2], hi[3]);
            sts_f4(base + ST_MC_LO + rowo + so, lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_bw_full);
        // ---- G -> M: row tok, columns [32 hb, 32 hb + 32) ----
        {
          float gv[32];
          const bool all_masked = (hb == 1) && ((wq & 1) == 0);     // tokens 0-31 against columns 32-63: j > i everywhere (warp-uniform)
          if (!all_masked) {
            mbar_wait(bar_g_full(buf), u & 1);
            tc_fence_after();
            tmem_ld_32x32(tmem_base + lane_sel + ST_COL_YG + (uint32_t)buf * 64u + 32u * hb, gv);
            tc_fence_before();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int j = 32 * hb + jj;
              const float cj = lds_f1(dts + 512 + (buf * 64 + j) * 4), cjl = lds_f1(dts + 1024 + (buf * 64 + j) * 4);
              const float m = gv[jj] * fast_exp_f((cum_c - cj) + (cuml_c - cjl));
              gv[jj] = (j <= tok) ? m : 0.f;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) gv[jj] = 0.f;
          }
          const uint32_t rowo = (uint32_t)hb * 8192u + (uint32_t)tok * 128u;
#pragma unroll
          for (int s8 = 0; s8 < 8; ++s8) {
            float hi[4], lo[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { hi[kk] = tf32_hi(gv[4 * s8 + kk]); lo[kk] = tf32_hi(gv[4 * s8 + kk] - hi[kk]); }
            const uint32_t so = (uint32_t)((s8 ^ sw) << 4);
            sts_f4(base + ST_MC_HI + rowo + so, hi[0], hi[1], hi[2], hi[3]);
            sts_f4(base + ST_MC_LO + rowo + so, lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_mc_full);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) vc[i] = vn[i];
      cum_c = cum_n; cuml_c = cuml_n; cumQ_c = cumQ_n; cumQl_c = cumQl_n;
    }
  } else if (warp < ST_WARP_Y0) {
    // ===================================== X group: thread = channel, warp half = tokens [32 hx, 32 hx + 32) ======================================
    const int wq = warp & 3, hx = warp >> 2;
    const int pch = 32 * wq + lane;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    int64_t cc = 0;
    for (int64_t item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int b = (int)(item / p.H), h = (int)(item - (int64_t)b * p.H);
      const int ch = h * ST_P + pch;
      float cw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cw[k] = (k >= 4 - p.kconv) ? p.conv_w[(size_t)ch * p.kconv + k - (4 - p.kconv)] : 0.f;
      const float cb = p.conv_b[ch];
      const float Dh = p.D ? p.D[h] : 0.f;
      const float* xp = p.z + ((size_t)b * T) * p.ldz + ch;
      for (int c = 0; c < nch; ++c, ++cc) {
        const int buf = (int)(cc & 1); const uint32_t u = (uint32_t)(cc >> 1);
        const int ys = (int)(cc % 3); const uint32_t uy = (uint32_t)(cc / 3);
        const int64_t t0 = (int64_t)c * ST_Q + 32 * hx;             // first token of this warp's half
        float xr[35];                                               // raw x of tokens t0 - 3 .. t0 + 31 (zero outside the sequence)
        {
          const float* xrow = xp + (t0 - ST_HIST) * p.ldz;          // never dereferenced outside [i_lo, i_hi)
          const int i_lo = t0 >= ST_HIST ? 0 : (int)(ST_HIST - t0);
          const int64_t left = T - t0 + ST_HIST;
          const int i_hi = left >= 35 ? 35 : (left > 0 ? (int)left : 0);
#pragma unroll
          for (int i = 0; i < 35; ++i) xr[i] = (i >= i_lo && i < i_hi) ? __ldg(xrow + (int64_t)i * p.ldz) : 0.f;
        }
        mbar_wait(bar_tprep(buf), u & 1);
        mbar_wait(bar_ysfree(ys), (uy & 1) ^ 1);
        mbar_wait(bar_xd_free(buf), (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t yst = base + ST_YST + (uint32_t)ys * 32768u + (uint32_t)pch * 4u;
        const uint32_t xd = tmem_base + lane_sel + ST_COL_XD + (uint32_t)buf * 128u;
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int i = 16 * bt + jj;                             // token t0 + i, window xr[i .. i + 3]
            const float xv = st_silu(fmaf(cw[3], xr[i + 3], fmaf(cw[2], xr[i + 2], fmaf(cw[1], xr[i + 1], fmaf(cw[0], xr[i], cb)))));
            const int j = 32 * hx + i;
            sts_f1(yst + (uint32_t)j * 512u, Dh * xv);
            const float xdv = lds_f1(dts + (buf * 64 + j) * 4) * xv;
            const float h_ = tf32_hi(xdv);
            hi[jj] = __float_as_uint(h_); lo[jj] = __float_as_uint(tf32_hi(xdv - h_));
          }
          tmem_st_32x16(xd + (uint32_t)(32 * hx + 16 * bt), hi);
          tmem_st_32x16(xd + 64u + (uint32_t)(32 * hx + 16 * bt), lo);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_xd_full(buf));
      }
    }
  } else if (warp >= ST_WARP_Y0 && warp < ST_WARP_Y0 + 4) {
    // ===================================== Y group: thread = channel: state update and output ======================================
    const int wq = warp - ST_WARP_Y0;
    const int pch = 32 * wq + lane;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const bool storer = (warp == ST_WARP_Y0) && elect_one();
    int64_t cc = 0;
    for (int64_t item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int b = (int)(item / p.H), h = (int)(item - (int64_t)b * p.H);
      float S[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) S[n] = 0.f;
      for (int c = 0; c < nch; ++c, ++cc) {
        const int buf = (int)(cc & 1); const uint32_t u = (uint32_t)(cc >> 1);
        const int ys = (int)(cc % 3);
        // ---- S = e^{cum_Q} S + dS -> tf32 hi / lo -> TMEM operand of the next chunk ----
        mbar_wait(bar_ds_full, (uint32_t)(cc & 1));
        tc_fence_after();
        float ds[16];
        tmem_ld_32x16(tmem_base + lane_sel + ST_COL_DS, ds);
        mbar_wait(bar_tprep(buf), u & 1);
        const float decay = expf(lds_f1(dts + 512 + (buf * 64 + 63) * 4));
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          S[n] = fmaf(decay, S[n], ds[n]);
          const float h_ = tf32_hi(S[n]);
          hi[n] = __float_as_uint(h_); lo[n] = __float_as_uint(tf32_hi(S[n] - h_));
        }
        const uint32_t sop = tmem_base + lane_sel + ST_COL_SOP + (uint32_t)(cc & 1) * 32u;
        tmem_st_32x16(sop, hi);
        tmem_st_32x16(sop + 16u, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_sop_full);
        mbar_arrive(bar_ydone(buf));
        // ---- y = Y^T + D x: accumulate into the staging tile, one TMA store per chunk ----
        mbar_wait(bar_y_full(buf), u & 1);
        tc_fence_after();
        float acc0[32], acc1[32];
        tmem_ld_32x32(tmem_base + lane_sel + ST_COL_YG + (uint32_t)buf * 64u, acc0);
        tmem_ld_32x32(tmem_base + lane_sel + ST_COL_YG + (uint32_t)buf * 64u + 32u, acc1);
        tc_fence_before();
        mbar_arrive(bar_yg_free(buf));
        mbar_wait(bar_xd_full(buf), u & 1);                          // the X group's D x of this chunk is in the staging tile
        const uint32_t yst = base + ST_YST + (uint32_t)ys * 32768u + (uint32_t)pch * 4u;
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_f1(yst + (uint32_t)j * 512u, lds_f1(yst + (uint32_t)j * 512u) + acc0[j]);
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_f1(yst + (uint32_t)(32 + j) * 512u, lds_f1(yst + (uint32_t)(32 + j) * 512u) + acc1[j]);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (storer) {
          tma_store_3d(&tmapY, base + ST_YST + (uint32_t)ys * 32768u, h * ST_P, c * ST_Q, b);
          bulk_commit();
          bulk_wait_read<1>();                                       // the PREVIOUS chunk's store has read its tile: hand that tile back to the X group
          if (cc > 0) mbar_arrive(bar_ysfree((int)((cc - 1) % 3)));
        }
      }
    }
    if (storer) { bulk_wait_read<0>(); bulk_wait<0>(); }
  } else if (warp == ST_WARP_MMA) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc_g = umma_idesc_tf32(128, ST_Q), idesc_s = umma_idesc_tf32(128, ST_N);
      const uint64_t d_gah = umma_desc_k_sw128(base + ST_GA_HI), d_gal = umma_desc_k_sw128(base + ST_GA_LO);
      const uint64_t d_gbh = umma_desc_k_sw128(base + ST_GB_HI), d_gbl = umma_desc_k_sw128(base + ST_GB_LO);
      const int64_t my_items = (p.nitems - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x;
      const int64_t nwork = my_items * nch;
      for (int64_t w = -1; w < nwork; ++w) {
        if (w + 1 < nwork) {
          // ---- G(w + 1) = C B^T (K = 16), one chunk ahead of the products that need M ----
          const int64_t wn = w + 1;
          const int bufn = (int)(wn & 1); const uint32_t un = (uint32_t)(wn >> 1);
          const uint32_t d_g = tmem_base + ST_COL_YG + (uint32_t)bufn * 64u;
          mbar_wait_one(bar_gop_full, (uint32_t)(wn & 1));
          mbar_wait_one(bar_yg_free(bufn), (un & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_tf32(d_g, d_gah + 2u * k, d_gbh + 2u * k, idesc_g, k > 0 ? 1u : 0u);
            umma_tf32(d_g, d_gah + 2u * k, d_gbl + 2u * k, idesc_g, 1u);
            umma_tf32(d_g, d_gal + 2u * k, d_gbh + 2u * k, idesc_g, 1u);
          }
          umma_commit(bar_g_full(bufn));
          umma_commit(bar_gop_free);
        }
        if (w < 0) continue;
        const int buf = (int)(w & 1); const uint32_t u = (uint32_t)(w >> 1);
        const int c = (int)(w % nch);                               // chunk index inside its sequence
        const uint32_t d_yg = tmem_base + ST_COL_YG + (uint32_t)buf * 64u;
        const uint32_t a_xh = tmem_base + ST_COL_XD + (uint32_t)buf * 128u, a_xl = a_xh + 64u;
        // ---- dS = Xd^T (B e^{cum_Q - cum})  (K = 64 tokens, N = 16) ----
        mbar_wait_one(bar_xd_full(buf), u & 1);
        mbar_wait_one(bar_bw_full, (uint32_t)(w & 1));
        if (w > 0) mbar_wait_one(bar_sop_full, (uint32_t)((w - 1) & 1));   // the Y group has read the previous dS (and written the state operand)
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dbh = umma_desc_k_sw128(base + ST_BW_HI + (uint32_t)(ks >> 2) * 2048u) + 2u * (ks & 3);
          const uint64_t dbl = umma_desc_k_sw128(base + ST_BW_LO + (uint32_t)(ks >> 2) * 2048u) + 2u * (ks & 3);
          umma_tf32_ts(tmem_base + ST_COL_DS, a_xh + 8u * ks, dbh, idesc_s, ks > 0 ? 1u : 0u);
          umma_tf32_ts(tmem_base + ST_COL_DS, a_xh + 8u * ks, dbl, idesc_s, 1u);
          umma_tf32_ts(tmem_base + ST_COL_DS, a_xl + 8u * ks, dbh, idesc_s, 1u);
        }
        umma_commit(bar_ds_full);
        // ---- Y^T = Xd^T M^T (K = 64 tokens) + S_prev (C e^{cum})^T (K = 16 states; not for the first chunk of a sequence: S = 0) ----
        mbar_wait_one(bar_mc_full, (uint32_t)(w & 1));
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dbh = umma_desc_k_sw128(base + ST_MC_HI + (uint32_t)(ks >> 2) * 8192u) + 2u * (ks & 3);
          const uint64_t dbl = umma_desc_k_sw128(base + ST_MC_LO + (uint32_t)(ks >> 2) * 8192u) + 2u * (ks & 3);
          umma_tf32_ts(d_yg, a_xh + 8u * ks, dbh, idesc_g, ks > 0 ? 1u : 0u);
          umma_tf32_ts(d_yg, a_xh + 8u * ks, dbl, idesc_g, 1u);
          umma_tf32_ts(d_yg, a_xl + 8u * ks, dbh, idesc_g, 1u);
        }
        if (c > 0) {
          const uint32_t a_sh = tmem_base + ST_COL_SOP + (uint32_t)((w - 1) & 1) * 32u, a_sl = a_sh + 16u;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t dbh = umma_desc_k_sw128(base + ST_MC_HI + 2u * 8192u) + 2u * k;
            const uint64_t dbl = umma_desc_k_sw128(base + ST_MC_LO + 2u * 8192u) + 2u * k;
            umma_tf32_ts(d_yg, a_sh + 8u * k, dbh, idesc_g, 1u);
            umma_tf32_ts(d_yg, a_sh + 8u * k, dbl, idesc_g, 1u);
            umma_tf32_ts(d_yg, a_sl + 8u * k, dbh, idesc_g, 1u);
          }
        }
        umma_commit(bar_y_full(buf));
        umma_commit(bar_xd_free(buf));
        umma_commit(bar_mc_free);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ST_WARP_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled3 st_encode_fn() {
  static PFN_encodeTiled3 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled3>(ptr);
  }
  return fn;
}
// (B, T, cols) fp32 viewed as a 3-D tensor {cols, T, B} with row stride ld: box {box_cols, box_rows, 1}, no swizzle
static int st_make_tmap3(CUtensorMap* map, const float* ptr, uint64_t cols, uint64_t T, uint64_t B, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled3 enc = st_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EIGB200_ECUDA; }
  cuuint64_t gdim[3] = {cols, T, B};
  cuuint64_t gstride[2] = {ld * sizeof(float), T * ld * sizeof(float)};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("ssd_tc: cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu T=%llu B=%llu ld=%llu)", (int)r,
                                     (unsigned long long)cols, (unsigned long long)T, (unsigned long long)B, (unsigned long long)ld); return EIGB200_ECUDA; }
  return EIGB200_OK;
}

bool ssd_tc_ok(const SsdParams& p) {
  if (!p.fused || p.P != ST_P || p.N != ST_N || p.kconv < 1 || p.kconv > 4 || !p.conv_w || !p.conv_b || p.final_state) return false;
  if (p.ldx != p.ldbc || p.ldx != p.lddt || p.ldx % 4 != 0 || p.ldy % 4 != 0) return false;
  if (((uintptr_t)p.x & 15) || ((uintptr_t)p.y & 15) || p.H % p.G != 0) return false;
  if (p.Bm < p.x || p.Cm < p.x || p.dt < p.x) return false;
  if ((p.Bm - p.x) + (int64_t)p.G * ST_N > p.ldx || (p.Cm - p.x) + (int64_t)p.G * ST_N > p.ldx || (p.dt - p.x) + p.H > p.ldx) return false;
  return p.T >= 1 && p.T < (1LL << 30);
}

int launch_ssd_tc(cudaStream_t st, const SsdParams& p, int64_t B) {
  CUtensorMap tZ, tY;
  int rc;
  if ((rc = st_make_tmap3(&tZ, p.x, (uint64_t)p.ldx, (uint64_t)p.T, (uint64_t)B, (uint64_t)p.ldx, ST_N, ST_RAWROWS))) return rc;
  if ((rc = st_make_tmap3(&tY, p.y, (uint64_t)p.H * ST_P, (uint64_t)p.T, (uint64_t)B, (uint64_t)p.ldy, ST_P, ST_Q))) return rc;
  SsdTcParams q{};
  q.z = p.x; q.ldz = p.ldx; q.colB = (int)(p.Bm - p.x); q.colC = (int)(p.Cm - p.x); q.colDt = (int)(p.dt - p.x);
  q.A_log = p.A; q.D = p.D; q.dt_bias = p.dt_bias; q.conv_w = p.conv_w; q.conv_b = p.conv_b;
  q.T = p.T; q.H = p.H; q.G = p.G; q.kconv = p.kconv; q.nitems = B * p.H; q.nchunks = (int)((p.T + ST_Q - 1) / ST_Q); q.zero = 0;
  const int64_t grid = q.nitems < (int64_t)num_sms() ? q.nitems : (int64_t)num_sms();
  const size_t smem = ST_SMEM + 1024;
  EIGB_CUDA(cudaFuncSetAttribute(ssd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ssd_tc_kernel<<<(unsigned)grid, ST_THREADS, smem, st>>>(tZ, tY, q);
  EIGB_LAUNCH_CHECK("ssd_tc_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200
