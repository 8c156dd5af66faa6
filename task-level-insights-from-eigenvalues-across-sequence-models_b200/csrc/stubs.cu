// stubs.cu -- entry points declared in include/eigb200.h whose kernels have not landed yet.
// Each returns EIGB200_EUNSUPPORTED with a message; they are removed one by one as the kernels are written.
#include "common.cuh"
using namespace eigb200;
#define EIGB_STUB(name) do { set_error(name ": not implemented in this build"); return EIGB200_EUNSUPPORTED; } while (0)

extern "C" int eigb200_softmax_nu(void*, const float*, const float*, int64_t, int64_t, int64_t, int, int, double*, float*) { EIGB_STUB("eigb200_softmax_nu"); }
extern "C" int eigb200_softmax_eta(void*, const double*, const float*, int64_t, int64_t, int, double*, int32_t*, const double*, int) { EIGB_STUB("eigb200_softmax_eta"); }
