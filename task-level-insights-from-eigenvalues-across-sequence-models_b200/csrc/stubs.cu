// stubs.cu -- entry points declared in include/eigb200.h whose kernels have not landed yet.
// Each returns EIGB200_EUNSUPPORTED with a message; they are removed one by one as the kernels are written.
#include "common.cuh"
using namespace eigb200;
#define EIGB_STUB(name) do { set_error(name ": not implemented in this build"); return EIGB200_EUNSUPPORTED; } while (0)

extern "C" int eigb200_linattn_nu(void*, const float*, const float*, int64_t, int64_t, int64_t, int, int, double*) { EIGB_STUB("eigb200_linattn_nu"); }
extern "C" int eigb200_softmax_nu(void*, const float*, const float*, int64_t, int64_t, int64_t, int, int, double*, float*) { EIGB_STUB("eigb200_softmax_nu"); }
extern "C" int eigb200_softmax_eta(void*, const double*, const float*, int64_t, int64_t, int, double*, int32_t*, const double*, int) { EIGB_STUB("eigb200_softmax_eta"); }
extern "C" int eigb200_diag_scan(void*, const float*, const float*, float*, int64_t, int64_t, int, int) { EIGB_STUB("eigb200_diag_scan"); }
extern "C" int eigb200_ssd_scan(void*, const float*, int64_t, const float*, const float*, const float*, const float*, int64_t, const float*, float*, int64_t, float*, int64_t, int64_t, int, int, int, int) { EIGB_STUB("eigb200_ssd_scan"); }
extern "C" int eigb200_mamba_conv_ssd(void*, const float*, int64_t, const float*, const float*, int, const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int, int, int, int) { EIGB_STUB("eigb200_mamba_conv_ssd"); }
extern "C" int eigb200_dplr_abar(void*, const float*, const float*, const float*, const float*, int64_t, int, float*) { EIGB_STUB("eigb200_dplr_abar"); }
extern "C" int eigb200_eigvals_c64(void*, float*, int64_t, int, float*, int32_t*) { EIGB_STUB("eigb200_eigvals_c64"); }
extern "C" size_t eigb200_linear_workspace_bytes(int, int) { return 0; }
extern "C" int eigb200_linear(void*, const float*, int64_t, const float*, const float*, float*, int64_t, const float*, int64_t, int64_t, int, int, int, int, void*, size_t) { EIGB_STUB("eigb200_linear"); }
extern "C" int eigb200_embedding(void*, const int64_t*, const float*, const float*, float*, int64_t, int64_t, int, int64_t) { EIGB_STUB("eigb200_embedding"); }
extern "C" int eigb200_layernorm(void*, const float*, const float*, const float*, float, float*, int64_t, int) { EIGB_STUB("eigb200_layernorm"); }
extern "C" int eigb200_conv_silu(void*, const float*, int64_t, const float*, const float*, int, float*, int64_t, int64_t, int64_t, int) { EIGB_STUB("eigb200_conv_silu"); }
extern "C" int eigb200_linattn_forward(void*, const float*, const float*, const float*, int64_t, const float*, int, int, float, float*, int64_t, int64_t, int64_t, int, int, int) { EIGB_STUB("eigb200_linattn_forward"); }
extern "C" int eigb200_add(void*, const float*, const float*, float*, int64_t) { EIGB_STUB("eigb200_add"); }
extern "C" int eigb200_mul_silu(void*, const float*, const float*, float*, int64_t) { EIGB_STUB("eigb200_mul_silu"); }
extern "C" int eigb200_gelu(void*, const float*, float*, int64_t) { EIGB_STUB("eigb200_gelu"); }
