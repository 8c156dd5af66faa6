// gemm_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) GEMM, k4_gemm_tc.cu.
#pragma once
#include "gemm_simt.cuh"
namespace eigb200 {
size_t tc_workspace_bytes(int N, int K);                 // resident-weight kernels only (K <= 256)
size_t tc_workspace_bytes_m(int64_t M, int N, int K);    // any supported shape: adds the split copy of A the streamed-operand kernel needs
bool tc_supported(const LinearParams& p);
// nsplit = 3: 3xTF32 error-compensated (fp32-level accuracy); nsplit = 1: plain TF32.
// p.W == nullptr: the workspace already holds the prepared weights (tc_prepare / eigb200_linear_prepare) of this (N, K, epilogue, LayerNorm).
int launch_linear_tc(cudaStream_t st, const LinearParams& p, int nsplit, void* workspace);
int tc_prepare(cudaStream_t st, const LinearParams& p, void* workspace);
}  // namespace eigb200
