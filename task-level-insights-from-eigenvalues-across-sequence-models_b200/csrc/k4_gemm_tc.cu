// k4_gemm_tc.cu -- placeholder until the tcgen05 kernel lands: reports "unsupported" so AUTO falls back to FFMA.
#include "gemm_tc.cuh"
namespace eigb200 {
size_t tc_workspace_bytes(int, int) { return 0; }
bool tc_supported(const LinearParams&) { return false; }
int launch_linear_tc(cudaStream_t, const LinearParams&, int, void*) { set_error("tcgen05 GEMM not built"); return EIGB200_EUNSUPPORTED; }
}  // namespace eigb200
