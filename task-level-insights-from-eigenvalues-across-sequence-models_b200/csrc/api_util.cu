// api_util.cu -- error reporting, device queries and host-side threshold-edge construction for libeigb200.
#include "common.cuh"
#include <stdarg.h>

namespace eigb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return EIGB200_ECUDA;
}
int num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) { cached = n; cached_dev = dev; }
  }
  return cached;
}

static float floor32(double t) { float f = (float)t; return ((double)f > t) ? nextafterf(f, -INFINITY) : f; }
static float ceil32(double t)  { float f = (float)t; return ((double)f < t) ? nextafterf(f, INFINITY) : f; }

int make_edges_f(const double* thr, int nthr, int compare_mode, EdgesF* e) {
  EIGB_CHECK_ARG(thr && nthr >= 1 && nthr <= 6, "thresholds: need 1..6 thresholds, got %d", nthr);
  EIGB_CHECK_ARG(compare_mode == EIGB200_CMP_F64 || compare_mode == EIGB200_CMP_F32, "bad compare_mode %d", compare_mode);
  e->nb = nthr + 1;
  for (int j = 0; j < 7; ++j) { e->lo[j] = INFINITY; e->hi[j] = -INFINITY; }
  for (int j = 0; j < nthr; ++j) {
    const double lo = j == 0 ? 0.0 : thr[j - 1], hi = thr[j];
    if (compare_mode == EIGB200_CMP_F64) { e->lo[j] = j == 0 ? 0.f : ceil32(lo); e->hi[j] = floor32(hi); }
    else { e->lo[j] = (float)lo; e->hi[j] = (float)hi; }
  }
  e->gt = compare_mode == EIGB200_CMP_F64 ? floor32(thr[nthr - 1]) : (float)thr[nthr - 1];
  return EIGB200_OK;
}
int make_edges_d(const double* thr, int nthr, EdgesD* e) {
  EIGB_CHECK_ARG(thr && nthr >= 1 && nthr <= 6, "thresholds: need 1..6 thresholds, got %d", nthr);
  e->nb = nthr + 1;
  for (int j = 0; j < 7; ++j) { e->lo[j] = INFINITY; e->hi[j] = -INFINITY; }
  for (int j = 0; j < nthr; ++j) { e->lo[j] = j == 0 ? 0.0 : thr[j - 1]; e->hi[j] = thr[j]; }
  e->gt = thr[nthr - 1];
  return EIGB200_OK;
}

}  // namespace eigb200

extern "C" int eigb200_version(void) { return EIGB200_VERSION; }
extern "C" const char* eigb200_last_error(void) { return eigb200::g_err; }
extern "C" int eigb200_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return eigb200::cuda_fail(e, "cudaGetDeviceProperties");
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return EIGB200_OK;
}
extern "C" int eigb200_set_device(int dev) {
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return eigb200::cuda_fail(e, "cudaSetDevice");
  return EIGB200_OK;
}
