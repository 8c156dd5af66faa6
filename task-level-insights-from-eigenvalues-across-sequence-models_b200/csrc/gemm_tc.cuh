// gemm_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) GEMM, k4_gemm_tc.cu.
#pragma once
#include "gemm_simt.cuh"
namespace eigb200 {
size_t tc_workspace_bytes(int N, int K);
bool tc_supported(const LinearParams& p);
// nsplit = 3: 3xTF32 error-compensated (fp32-level accuracy); nsplit = 1: plain TF32.
int launch_linear_tc(cudaStream_t st, const LinearParams& p, int nsplit, void* workspace);
}  // namespace eigb200
