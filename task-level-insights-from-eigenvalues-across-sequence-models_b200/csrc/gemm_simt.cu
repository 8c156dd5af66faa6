// gemm_simt.cu -- nn.Linear on the fp32 FFMA pipe.  128x64 CTA tile, BK=16, 256 threads, 8x4 register micro-tile,
// register-staged double buffering.  Used for small problems, odd shapes, and as the bit-faithful fp32 checker of the
// tensor-core path (k4_gemm_tc.cu).  The tile's 64 columns are two 32-wide halves: for the GLU epilogue the halves are the
// value columns [n0, n0+32) and their gate columns [Nout+n0, Nout+n0+32), so value and gate meet in the same thread.
#include "gemm_simt.cuh"

namespace eigb200 {

constexpr int GS_BM = 128, GS_BN = 64, GS_BK = 16, GS_THREADS = 256;

__device__ __forceinline__ int glu_row(int local, int n0, int Nout, bool glu) {
  // map local tile column (0..63) to a row of W
  if (!glu) return n0 + local;
  return local < 32 ? n0 + local : Nout + n0 + (local - 32);
}

template <int EPI>
__global__ void __launch_bounds__(GS_THREADS) linear_simt_kernel(const LinearParams p) {
  __shared__ __align__(16) float As[2][GS_BK][GS_BM + 4];
  __shared__ __align__(16) float Ws[2][GS_BK][GS_BN + 4];
  constexpr bool GLU = EPI == EIGB200_EPI_GLU_RESIDUAL;
  const int Nout = GLU ? p.N / 2 : p.N;
  const int ncols_tile = GLU ? 32 : 64;                 // output columns produced per CTA
  const int64_t m0 = (int64_t)blockIdx.y * GS_BM;
  const int n0 = blockIdx.x * ncols_tile;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;               // 16 x 16 thread grid: 8 rows x (2+2) cols each
  const int K = p.K;

  // global -> register staging indices
  const int a_row = tid >> 1, a_k4 = (tid & 1) * 2;     // each thread: rows a_row, 2 float4 along K (k = a_k4*4 .. +7)
  const int w_row = tid >> 2, w_k4 = tid & 3;           // 64 rows x 4 float4
  const int64_t am = m0 + a_row;
  const int wn = glu_row(w_row, n0, Nout, GLU);
  const bool a_ok = am < p.M;
  const bool w_ok = GLU ? (n0 + (w_row & 31) < Nout) : (wn < p.N);
  const float* Ap = p.A + am * p.lda;
  const float* Wp = p.W + (size_t)wn * K;
  const bool vecA = (p.lda % 4 == 0) && (((uintptr_t)p.A & 15) == 0);
  const bool vecW = (K % 4 == 0) && (((uintptr_t)p.W & 15) == 0);

  float4 ra[2], rw;
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + (a_k4 + i) * 4;
      if (a_ok && vecA && k + 3 < K) ra[i] = __ldg(reinterpret_cast<const float4*>(Ap + k));
      else {
        ra[i].x = (a_ok && k + 0 < K) ? Ap[k + 0] : 0.f; ra[i].y = (a_ok && k + 1 < K) ? Ap[k + 1] : 0.f;
        ra[i].z = (a_ok && k + 2 < K) ? Ap[k + 2] : 0.f; ra[i].w = (a_ok && k + 3 < K) ? Ap[k + 3] : 0.f;
      }
    }
    const int k = k0 + w_k4 * 4;
    if (w_ok && vecW && k + 3 < K) rw = __ldg(reinterpret_cast<const float4*>(Wp + k));
    else {
      rw.x = (w_ok && k + 0 < K) ? Wp[k + 0] : 0.f; rw.y = (w_ok && k + 1 < K) ? Wp[k + 1] : 0.f;
      rw.z = (w_ok && k + 2 < K) ? Wp[k + 2] : 0.f; rw.w = (w_ok && k + 3 < K) ? Wp[k + 3] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = (a_k4 + i) * 4;
      As[buf][k + 0][a_row] = ra[i].x; As[buf][k + 1][a_row] = ra[i].y; As[buf][k + 2][a_row] = ra[i].z; As[buf][k + 3][a_row] = ra[i].w;
    }
    const int k = w_k4 * 4;
    Ws[buf][k + 0][w_row] = rw.x; Ws[buf][k + 1][w_row] = rw.y; Ws[buf][k + 2][w_row] = rw.z; Ws[buf][k + 3][w_row] = rw.w;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + GS_BK - 1) / GS_BK;
  gload(0); sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GS_BK);
#pragma unroll
    for (int k = 0; k < GS_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float2 b0 = *reinterpret_cast<const float2*>(&Ws[buf][k][tx * 2]);
      const float2 b1 = *reinterpret_cast<const float2*>(&Ws[buf][k][32 + tx * 2]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) { sstore(buf ^ 1); }
    __syncthreads();
  }

  // ---- epilogue -----------------------------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    if (GLU) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + tx * 2 + j;
        if (n >= Nout) continue;
        float v = acc[i][j], g = acc[i][2 + j];
        if (p.bias) { v += p.bias[n]; g += p.bias[Nout + n]; }
        float o = v * sigmoid_f(g);
        if (p.R) o += p.R[m * p.ldr + n];
        p.C[m * p.ldc + n] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + (j < 2 ? tx * 2 + j : 32 + tx * 2 + (j - 2));
        if (n >= p.N) continue;
        float v = acc[i][j];
        if (p.bias) v += p.bias[n];
        if (EPI == EIGB200_EPI_GELU) v = gelu_f(v);
        if (EPI == EIGB200_EPI_RESIDUAL && p.R) v += p.R[m * p.ldr + n];
        p.C[m * p.ldc + n] = v;
      }
    }
  }
}

int launch_linear_simt(cudaStream_t st, const LinearParams& p) {
  const bool glu = p.epilogue == EIGB200_EPI_GLU_RESIDUAL;
  const int nout = glu ? p.N / 2 : p.N;
  const int per = glu ? 32 : 64;
  dim3 grid((nout + per - 1) / per, (unsigned)((p.M + GS_BM - 1) / GS_BM));
  EIGB_CHECK_ARG(grid.y <= 65535u * 32768u, "linear: M too large");
  if (grid.y > 65535) { set_error("linear(simt): M=%lld exceeds 65535 row tiles; split the call", (long long)p.M); return EIGB200_EINVAL; }
  switch (p.epilogue) {
    case EIGB200_EPI_NONE: linear_simt_kernel<EIGB200_EPI_NONE><<<grid, GS_THREADS, 0, st>>>(p); break;
    case EIGB200_EPI_GELU: linear_simt_kernel<EIGB200_EPI_GELU><<<grid, GS_THREADS, 0, st>>>(p); break;
    case EIGB200_EPI_GLU_RESIDUAL: linear_simt_kernel<EIGB200_EPI_GLU_RESIDUAL><<<grid, GS_THREADS, 0, st>>>(p); break;
    case EIGB200_EPI_RESIDUAL: linear_simt_kernel<EIGB200_EPI_RESIDUAL><<<grid, GS_THREADS, 0, st>>>(p); break;
    default: set_error("linear: unknown epilogue %d", p.epilogue); return EIGB200_EINVAL;
  }
  EIGB_LAUNCH_CHECK("linear_simt_kernel");
  return EIGB200_OK;
}

}  // namespace eigb200
