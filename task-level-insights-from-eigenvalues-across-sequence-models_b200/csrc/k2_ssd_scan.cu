// k2_ssd_scan.cu -- K2b: SSD selective scan, optionally fused with the depthwise causal conv + SiLU and the
// softplus(dt) that precede it in SSD.forward.
//
// Reference operators replaced:
//   mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=D, z=None)    models/mamba.py:138-150  (third-party mamba-ssm 2.1.0)
//   conv1d + SiLU + truncate, softplus(dt + dt_bias), A = -exp(A_log)      models/mamba.py:119-133
// Recurrence (per batch b, head h, group g = h / (H/G)):
//   S_t[p,n] = exp(dt_t A) S_{t-1}[p,n] + (dt_t x_t[p]) B_t[n];    y_t[p] = sum_n C_t[n] S_t[p,n] + D x_t[p]
//
// Roofline: HBM at small d_state (N=16: (2HP+2GN)*4+4H bytes vs 6HPN flops per token ~ 7 flop/B, at the fp32 SIMT ridge);
// FMA-pipe bound for large N.  Layout: every operand row-major over (b,t) with the channel axis contiguous.
// Mapping: a thread owns one head-channel p and an NS-wide slice of the state row S[p, n0:n0+NS] in registers; N/NS
// adjacent lanes share p and combine their partial y with shuffles.  A CTA (128 threads) owns 128/(N/NS) channels of one
// (b,h) and walks time in chunks of TC tokens: the raw rows of the chunk (+3 history rows for the 4-tap conv) are staged
// in shared memory with coalesced 128-bit loads, a prep phase turns them into conv'd B_t, C_t and (dt_t, exp(dt_t A)),
// then the serial phase runs TC steps out of shared memory (broadcast LDS.128 for B_t/C_t).  Several CTAs per SM overlap
// one CTA's staging with another's recurrence.
#include "common.cuh"

namespace eigb200 {

constexpr int SSD_TC = 32;          // tokens per staged chunk
constexpr int SSD_HIST = 3;         // history rows kept for the conv (max 4 taps)
constexpr int SSD_THREADS = 128;

struct SsdParams {
  // x channels, B channels, C channels, dt: each (b,t) row-major with its own row stride (elements)
  const float* x; int64_t ldx;
  const float* Bm; const float* Cm; int64_t ldbc;
  const float* dt; int64_t lddt;
  const float* A;            // fused: A_log (A = -exp(A_log));  plain: A itself
  const float* D;
  const float* dt_bias;      // fused only
  const float* conv_w;       // fused only: (H*P + 2*G*N, kconv) row-major over channels [x | B | C]
  const float* conv_b;
  float* y; int64_t ldy;
  float* final_state;        // (B,H,P,N) or null
  int64_t T; int H, P, G, N, kconv, fused;
};

template <int NS>
__global__ void __launch_bounds__(SSD_THREADS) ssd_scan_kernel(const SsdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int N = p.N, P = p.P;
  const int lpp = N / NS;                              // lanes per channel
  const int pc = SSD_THREADS / lpp;                    // channels per CTA
  const int b = blockIdx.z, h = blockIdx.y, pblk = blockIdx.x;
  const int g = h / (p.H / p.G);
  const int tid = threadIdx.x;
  const int pl = tid / lpp, sub = tid - pl * lpp;
  const int pch = pblk * pc + pl;                      // channel within the head
  const bool pvalid = pch < P;
  const int n0 = sub * NS;

  // shared memory carve-up
  float* xs = sm;                                      // [(TC+HIST)][pc]   raw x channels of this CTA
  float* bs = xs + (SSD_TC + SSD_HIST) * pc;           // [(TC+HIST)][N]    raw B
  float* cs = bs + (SSD_TC + SSD_HIST) * N;            // [(TC+HIST)][N]    raw C
  float* Bc = cs + (SSD_TC + SSD_HIST) * N;            // [TC][N]  conv'd B
  float* Cc = Bc + SSD_TC * N;                         // [TC][N]  conv'd C
  float* dts = Cc + SSD_TC * N;                        // [TC] dt, [TC] decay
  float* decs = dts + SSD_TC;

  const float Ah = p.fused ? -expf(p.A[h]) : p.A[h];
  const float Dh = p.D ? p.D[h] : 0.f;
  const float dtb = p.fused ? p.dt_bias[h] : 0.f;
  const bool conv = p.fused && p.kconv > 0;

  // conv taps for this thread's x channel, left-padded to 4 taps
  float cw[4] = {0.f, 0.f, 0.f, 1.f}, cb = 0.f;
  if (conv && pvalid) {
    const int ch = h * P + pch;
#pragma unroll
    for (int j = 0; j < 4; ++j) cw[j] = (j >= 4 - p.kconv) ? p.conv_w[(size_t)ch * p.kconv + j - (4 - p.kconv)] : 0.f;
    cb = p.conv_b[ch];
  }

  float s[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) s[i] = 0.f;
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;

  const size_t rowbase = (size_t)b * p.T;
  const float* xg = p.x + (size_t)h * P + (size_t)pblk * pc;
  const float* bg = p.Bm + (size_t)g * N;
  const float* cg = p.Cm + (size_t)g * N;
  const int HP = p.H * P;
  const int GN = p.G * N;

  for (int64_t t0 = 0; t0 < p.T; t0 += SSD_TC) {
    const int tc = (int)min((int64_t)SSD_TC, p.T - t0);
    __syncthreads();                                   // previous chunk fully consumed
    // ---- stage raw rows [t0-HIST, t0+tc) -------------------------------------------------------------
    const int nrows = tc + SSD_HIST;
    for (int i = tid; i < nrows * pc; i += SSD_THREADS) {
      const int r = i / pc, c = i - r * pc;
      const int64_t t = t0 - SSD_HIST + r;
      xs[i] = (t >= 0 && pblk * pc + c < P) ? __ldg(xg + (rowbase + t) * p.ldx + c) : 0.f;
    }
    for (int i = tid; i < nrows * N; i += SSD_THREADS) {
      const int r = i / N, c = i - r * N;
      const int64_t t = t0 - SSD_HIST + r;
      const bool ok = t >= 0;
      bs[i] = ok ? __ldg(bg + (rowbase + t) * p.ldbc + c) : 0.f;
      cs[i] = ok ? __ldg(cg + (rowbase + t) * p.ldbc + c) : 0.f;
    }
    if (tid < tc) {
      const float raw = __ldg(p.dt + (rowbase + t0 + tid) * p.lddt + h);
      const float d = p.fused ? softplus_f(raw + dtb) : raw;
      dts[tid] = d; decs[tid] = expf(d * Ah);
    }
    __syncthreads();
    // ---- prep: conv + SiLU of the B and C channels ------------------------------------------------------
    for (int i = tid; i < tc * N; i += SSD_THREADS) {
      const int r = i / N, c = i - r * N;
      if (conv) {
        const int chB = HP + g * N + c, chC = HP + GN + g * N + c;
        float ab = p.conv_b[chB], ac = p.conv_b[chC];
        for (int j = 0; j < p.kconv; ++j) {
          const int rr = r + SSD_HIST - (p.kconv - 1) + j;
          ab = fmaf(p.conv_w[(size_t)chB * p.kconv + j], bs[rr * N + c], ab);
          ac = fmaf(p.conv_w[(size_t)chC * p.kconv + j], cs[rr * N + c], ac);
        }
        Bc[i] = silu_f(ab); Cc[i] = silu_f(ac);
      } else {
        Bc[i] = bs[(r + SSD_HIST) * N + c]; Cc[i] = cs[(r + SSD_HIST) * N + c];
      }
    }
    __syncthreads();
    // ---- serial phase -----------------------------------------------------------------------------------
    if (t0 == 0) { w0 = xs[0 * pc + pl]; w1 = xs[1 * pc + pl]; w2 = xs[2 * pc + pl]; }   // zeros (padding)
    for (int tt = 0; tt < tc; ++tt) {
      const float w3 = xs[(tt + SSD_HIST) * pc + pl];
      float xc;
      if (conv) {
        xc = fmaf(cw[3], w3, fmaf(cw[2], w2, fmaf(cw[1], w1, fmaf(cw[0], w0, cb))));
        xc = silu_f(xc);
      } else xc = w3;
      w0 = w1; w1 = w2; w2 = w3;
      const float dtv = dts[tt], dec = decs[tt];
      const float u = xc * dtv;
      const float4* B4 = reinterpret_cast<const float4*>(Bc + tt * N + n0);
      const float4* C4 = reinterpret_cast<const float4*>(Cc + tt * N + n0);
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < NS / 4; ++q) {
        const float4 bv = B4[q], cv = C4[q];
        s[4 * q + 0] = fmaf(dec, s[4 * q + 0], u * bv.x); acc = fmaf(cv.x, s[4 * q + 0], acc);
        s[4 * q + 1] = fmaf(dec, s[4 * q + 1], u * bv.y); acc = fmaf(cv.y, s[4 * q + 1], acc);
        s[4 * q + 2] = fmaf(dec, s[4 * q + 2], u * bv.z); acc = fmaf(cv.z, s[4 * q + 2], acc);
        s[4 * q + 3] = fmaf(dec, s[4 * q + 3], u * bv.w); acc = fmaf(cv.w, s[4 * q + 3], acc);
      }
      for (int off = lpp >> 1; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (sub == 0 && pvalid) p.y[(rowbase + t0 + tt) * p.ldy + (size_t)h * P + pch] = fmaf(Dh, xc, acc);
    }
  }
  if (p.final_state && pvalid) {
    float* fs = p.final_state + (((size_t)b * p.H + h) * P + pch) * N + n0;
#pragma unroll
    for (int i = 0; i < NS; ++i) fs[i] = s[i];
  }
}

static int launch_ssd(cudaStream_t st, const SsdParams& p, int64_t B) {
  EIGB_CHECK_ARG(B > 0 && p.T > 0 && p.H > 0 && p.P > 0 && p.G > 0 && p.N > 0, "ssd_scan: bad shape");
  EIGB_CHECK_ARG(p.H % p.G == 0, "ssd_scan: nheads %d not divisible by ngroups %d", p.H, p.G);
  EIGB_CHECK_ARG(p.N % 4 == 0, "ssd_scan: d_state %d must be a multiple of 4", p.N);
  EIGB_CHECK_ARG(p.kconv >= 0 && p.kconv <= 4, "ssd_scan: conv kernel size %d not in 0..4", p.kconv);
  EIGB_CHECK_ARG(B <= 65535 && p.H <= 65535, "ssd_scan: batch/heads exceed grid limits");
  // state slice per thread: the largest of 16/8/4 that divides N with N/NS a power of two <= 32
  int ns = 0;
  for (int c : {16, 8, 4}) {
    if (p.N % c == 0) { const int l = p.N / c; if (l <= 32 && (l & (l - 1)) == 0) { ns = c; break; } }
  }
  EIGB_CHECK_ARG(ns != 0, "ssd_scan: unsupported d_state %d (need N/16, N/8 or N/4 a power of two <= 32)", p.N);
  const int lpp = p.N / ns, pc = SSD_THREADS / lpp;
  dim3 grid((p.P + pc - 1) / pc, p.H, (unsigned)B);
  const size_t smem = sizeof(float) * ((size_t)(SSD_TC + SSD_HIST) * (pc + 2 * p.N) + 2 * (size_t)SSD_TC * p.N + 2 * SSD_TC);
  EIGB_CHECK_ARG(smem <= 200 * 1024, "ssd_scan: d_state %d needs %zu bytes of shared memory", p.N, smem);
#define SSD_CASE(NS_)                                                                                                  \
  case NS_:                                                                                                            \
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(ssd_scan_kernel<NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ssd_scan_kernel<NS_><<<grid, SSD_THREADS, smem, st>>>(p);                                                           \
    break;
  switch (ns) { SSD_CASE(16) SSD_CASE(8) SSD_CASE(4) }
#undef SSD_CASE
  EIGB_LAUNCH_CHECK("ssd_scan_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_ssd_scan(void* stream, const float* d_x, int64_t ldx, const float* d_dt, const float* d_A,
                                const float* d_Bm, const float* d_Cm, int64_t ldbc, const float* d_D,
                                float* d_y, int64_t ldy, float* d_final_state,
                                int64_t B, int64_t T, int H, int P, int G, int N) {
  EIGB_CHECK_ARG(d_x && d_dt && d_A && d_Bm && d_Cm && d_y, "ssd_scan: null pointer");
  SsdParams p{};
  p.x = d_x; p.ldx = ldx; p.Bm = d_Bm; p.Cm = d_Cm; p.ldbc = ldbc; p.dt = d_dt; p.lddt = H;
  p.A = d_A; p.D = d_D; p.dt_bias = nullptr; p.conv_w = nullptr; p.conv_b = nullptr;
  p.y = d_y; p.ldy = ldy; p.final_state = d_final_state;
  p.T = T; p.H = H; p.P = P; p.G = G; p.N = N; p.kconv = 0; p.fused = 0;
  return launch_ssd((cudaStream_t)stream, p, B);
}

extern "C" int eigb200_mamba_conv_ssd(void* stream, const float* d_xbcdt, int64_t ldz, const float* d_conv_w, const float* d_conv_b, int kconv,
                                      const float* d_dt_bias, const float* d_A_log, const float* d_D,
                                      float* d_y, int64_t ldy, int64_t B, int64_t T, int H, int P, int G, int N) {
  EIGB_CHECK_ARG(d_xbcdt && d_dt_bias && d_A_log && d_y, "mamba_conv_ssd: null pointer");
  EIGB_CHECK_ARG(kconv == 0 || (d_conv_w && d_conv_b), "mamba_conv_ssd: conv weights missing");
  EIGB_CHECK_ARG(ldz >= (int64_t)H * P + 2 * (int64_t)G * N + H, "mamba_conv_ssd: row stride %lld too small", (long long)ldz);
  SsdParams p{};
  p.x = d_xbcdt; p.ldx = ldz;
  p.Bm = d_xbcdt + (size_t)H * P; p.Cm = d_xbcdt + (size_t)H * P + (size_t)G * N; p.ldbc = ldz;
  p.dt = d_xbcdt + (size_t)H * P + 2 * (size_t)G * N; p.lddt = ldz;
  p.A = d_A_log; p.D = d_D; p.dt_bias = d_dt_bias; p.conv_w = d_conv_w; p.conv_b = d_conv_b;
  p.y = d_y; p.ldy = ldy; p.final_state = nullptr;
  p.T = T; p.H = H; p.P = P; p.G = G; p.N = N; p.kconv = kconv; p.fused = 1;
  return launch_ssd((cudaStream_t)stream, p, B);
}
