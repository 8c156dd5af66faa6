// gemm_simt.cuh -- fp32 FFMA GEMM  C = epilogue(A W^T + bias)  (exact-fp32 path; also the checker for the tcgen05 path).
#pragma once
#include "common.cuh"
namespace eigb200 {
struct LinearParams {
  const float* A; int64_t lda; const float* W; const float* bias;
  float* C; int64_t ldc; const float* R; int64_t ldr;
  int64_t M; int N, K, epilogue;
  // optional LayerNorm fused into the A operand (tensor-core path only): (mean, rstd) per row, gamma / beta per column
  const float* ln_stats = nullptr; const float* ln_gamma = nullptr; const float* ln_beta = nullptr;
  // optional extractor fusion in the GLU epilogue (tensor-core path only): gate weights (N/2) and the partials buffer ((N/32) * 3 * M floats)
  const float* eig_w = nullptr; float* eig_part = nullptr;
};
int launch_linear_simt(cudaStream_t st, const LinearParams& p);
}  // namespace eigb200
