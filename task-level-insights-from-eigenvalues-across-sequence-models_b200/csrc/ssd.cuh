// ssd.cuh -- parameter block shared by the SSD kernels (k2_ssd_scan.cu: recurrent SIMT forms; k2_ssd_tc.cu: chunked form on the 5th-gen tensor cores).
#pragma once
#include "common.cuh"
namespace eigb200 {
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

#ifndef EIGB200_SSD_DEFAULT_TC
#define EIGB200_SSD_DEFAULT_TC 0
#endif
// Chunked SSD on tcgen05 (k2_ssd_tc.cu): fused conv + SiLU + softplus form, head dim 128, d_state 16
bool ssd_tc_ok(const SsdParams& p);
int launch_ssd_tc(cudaStream_t st, const SsdParams& p, int64_t B);
}  // namespace eigb200
