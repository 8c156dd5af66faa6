// tc_ptx.cuh -- inline-PTX wrappers shared by the tcgen05 kernels (k4_gemm_tc.cu, k4_gemm_fused.cu): mbarrier, TMA, TMEM allocation / load / store,
// tcgen05.mma issue and commit, shared-memory and instruction descriptors, operand splits.  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace eigb200 {

// ---------------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
// Arrive that is DATA-dependent on `dep`: the barrier address is bar + (dep & zero) with `zero` a kernel parameter that is 0 at run time, so
// neither nvcc nor ptxas can fold the dependency away and the arrive cannot issue before the registers that feed `dep` -- the values loaded
// from the buffer being released -- have landed.  (An asm operand that the template does not reference creates no dependency in the PTX:
// with it the arrive overtook the last LDS of the chunk and TMA refilled the slot under the reader.)
__device__ __forceinline__ void mbar_arrive_after(uint32_t bar, uint32_t dep, uint32_t zero) {
  mbar_arrive(bar + (dep & zero));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
// suspend-time hint of mbarrier.try_wait: the hardware parks the thread until the phase completes or the hint expires, instead of returning to
// the polling loop after the (short) default window -- polling (SYNCS / BRA / YIELD) was a third of all issued instructions without it
constexpr uint32_t MBAR_SUSPEND_HINT_NS = 0x989680u;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" :: "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
  // the polling loop lives inside the asm block, so the compiler inserts no reconvergence point after it: lanes may leave it on
  // different iterations.  Every caller goes on to .sync.aligned instructions (tcgen05.ld / st / wait) or lane-0 election, which need
  // the warp converged -- without this barrier single TMEM lanes (rows) were silently dropped by tcgen05.st.
  __syncwarp();
}
// spin on the barrier from a single elected thread (no warp re-convergence: the caller is the only active lane of its warp)
__device__ __forceinline__ void mbar_wait_one(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" :: "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
}
// elect.sync: true in exactly one lane of the (converged) warp.  Code guarded by it is a single-thread region for the compiler, so the
// operands of UTCHMMA / UTMALDG are trivially warp-uniform -- with `lane == 0` instead it wrapped every tcgen05.mma in an ELECT/BRA.U.ANY
// waterfall loop and the issue rate, not the tensor pipe, paced the kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp receives columns [col, col+32) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// store 32 columns of 32-bit into TMEM: thread i of the warp writes columns [col, col+32) of lane (lane_base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr),
         "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
         "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
         "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
         "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
         "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
         "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
         "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
         "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]^T: the A operand (128 lanes = rows, 8 columns = the K step of tf32 values) is read from TMEM
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// kind::f16 with the A operand in TMEM: 128 lanes x 8 columns hold the K step of 16 fp16 values (two per 32-bit column, lower K index in the low half)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with both operands in shared memory (K-major SWIZZLE_128B descriptors)
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// store 16 columns of 32-bit into TMEM: thread i of the warp writes columns [col, col+16) of lane (lane_base + i)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// {lo, hi} -> packed f16x2 (round to nearest even), and back
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_lo_to_f32(uint32_t h2) {
  float r;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(r) : "r"(h2));
  return r;
}
__device__ __forceinline__ float f16_hi_to_f32(uint32_t h2) {
  float r;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(r) : "r"(h2));
  return r;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Converter split a = hi + lo for the 3xTF32 product.  hi = a rounded to tf32 (nearest, ties away: integer add of half a tf32 ulp, then
// mask -- the same result as cvt.rna.tf32.f32 for finite a, in 2 ALU operations instead of the ~5 the cvt expands to); lo = a - hi is
// exact in fp32 and is handed to the tensor core unrounded: kind::tf32 reads the top 19 bits of the container, so lo is truncated to
// tf32 by the hardware (|error| <= 2^-21 |a|, sign of lo, i.e. unbiased), the same order as the dropped a_lo * w_lo term.
__device__ __forceinline__ void split_tf32(float a, float& h, float& l) {
  h = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
  l = a - h;
}
__device__ __forceinline__ void split_tf32_4(const float4 a, float4& h, float4& l) {
  split_tf32(a.x, h.x, l.x); split_tf32(a.y, h.y, l.y); split_tf32(a.z, h.z, l.z); split_tf32(a.w, h.w, l.w);
}

// 8x8 transpose of float4 items inside each group of 8 lanes.  In: lane (8g+i) holds, for its accumulator row 8g+i, the eight
// float4 column quads q = 0..7 of a 32-column group.  Out: the same lane holds quad q = i of rows 8g+j, j = 0..7 -- so that for
// every j the 8 lanes of a group cover one row's 128 contiguous bytes and a warp store instruction writes 4 full lines.
__device__ __forceinline__ void transpose8x8_f4(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
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

// 256-bit streaming load (sm_100): one full 32-byte sector per thread
__device__ __forceinline__ void ldg_stream_v8(const float* ptr, float* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(ptr));
}

// 256-bit store (sm_100): one full 32-byte sector per thread
__device__ __forceinline__ void stg_v8(float* ptr, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in bits [0,14); LBO (unused for swizzled K-major) = 1 in [16,30); SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46);
// descriptor version 1 in [46,48); layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A/B tf32 (bits 7-9, 10-12 = 2), both K-major
// (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16: A / B fp16 (format 0), D fp32
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// GLU column order inside a CTA slice (bn = 2*bg accumulator columns): 32-column accumulator group k holds the VALUE columns of output
// columns [16k, 16k+16) of the slice in its first half and their GATE columns in its second half, so ONE tcgen05.ld hands a warp both
// factors of 16 finished output columns and every accumulator group is independent work for an epilogue warp.
__host__ __device__ __forceinline__ int glu_weight_row(int local, int split, int bg, int nout) {
  const int c = split * bg + (local >> 5) * 16 + (local & 15);
  if (c >= nout) return -1;
  return (local & 16) ? nout + c : c;
}


}  // namespace eigb200
