// xla_ffi_shim.cc -- jax.ffi custom-call handlers over the C ABI (include/eigb200.h), one per entry point the JAX branch of the reference needs
// (analysis/eval_eig.py:303-329 get_eigvals_ssm, :254-301 discrete_DPLR + eigvals, models/lru.py:86-99 / models/s5.py:65-93 associative_scan) plus the
// Mamba-2 extractor and the threshold statistics, so that the Flax layers can call the sm_100a kernels as XLA custom calls:
//
//     jax.ffi.register_ffi_target("eigb200_diag_scan", jax.ffi.pycapsule(lib.eigb200_ffi_diag_scan), platform="CUDA")
//     h = jax.ffi.ffi_call("eigb200_diag_scan", jax.ShapeDtypeStruct(Bu.shape, Bu.dtype))(lam, Bu, reverse=np.int32(0))
//
// COMPILE-GUARDED: JAX / XLA are not part of this image, so this file only builds where xla/ffi/api/ffi.h is on the include path
// (e.g. `pip show jaxlib` -> site-packages/jaxlib/include):
//     g++ -O2 -std=c++17 -fPIC -shared -I$JAXLIB/include -I/usr/local/cuda/include -Iinclude xla_ffi_shim.cc -L. -leigb200 -o libeigb200_ffi.so
// Every handler is a 1:1 wrapper: XLA hands over device buffers and its stream, the handler forwards them to the C ABI, nothing is allocated or copied here.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define EIGB200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef EIGB200_HAVE_XLA_FFI
#include <cuda_runtime_api.h>
#include <complex>
#include <cstdint>
#include <string>

#include "eigb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {
ffi::Error Check(int rc, const char* what) {
  if (rc == EIGB200_OK) return ffi::Error::Success();
  return ffi::Error(rc == EIGB200_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + eigb200_last_error());
}
static const double kThrRadius[6] = {0.1, 0.5, 0.9, 1.0, 10.0, 100.0};      // analysis/eval_eig.py:603, :665, :724

// lambda (B,T,H) f32 and the per-sample radius bin counts (B,H,8) i32 of get_eig_mamba2 (eval_eig.py:176-190); counts must arrive zeroed (donated operand)
ffi::Error Mamba2Eig(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> w_dt, ffi::Buffer<ffi::F32> dt_bias, ffi::Buffer<ffi::F32> a_log,
                     ffi::ResultBuffer<ffi::F32> lam, ffi::ResultBuffer<ffi::S32> counts) {
  auto d = x.dimensions();
  if (d.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_mamba2_eig: x must be (B,T,D)");
  const int H = static_cast<int>(w_dt.dimensions()[0]);
  return Check(eigb200_mamba2_eig(stream, x.typed_data(), EIGB200_F32, d[0], d[1], static_cast<int>(d[2]), w_dt.typed_data(), dt_bias.typed_data(),
                                  a_log.typed_data(), H, lam->typed_data(), 1, counts->typed_data(), kThrRadius, 6, EIGB200_CMP_F64, nullptr, 1e-5f),
               "eigb200_mamba2_eig");
}

// parameter-only eigenvalues of LRU (kind 0: nu_log, theta_log) and S5 (kind 1 zoh / 2 bilinear: Lambda_re, Lambda_im, log_step)  (eval_eig.py:303-329)
ffi::Error SsmLambda(cudaStream_t stream, ffi::Buffer<ffi::F32> p0, ffi::Buffer<ffi::F32> p1, ffi::Buffer<ffi::F32> p2, ffi::ResultBuffer<ffi::C64> lam, int32_t kind) {
  const int P = static_cast<int>(p0.element_count());
  return Check(eigb200_ssm_lambda(stream, kind, p0.typed_data(), p1.typed_data(), kind == 0 ? nullptr : p2.typed_data(), P,
                                  reinterpret_cast<float*>(lam->typed_data())), "eigb200_ssm_lambda");
}

// h_t = lam h_{t-1} + Bu_t: what associative_scan(binary_operator_diag, ...) evaluates (models/lru.py:95, models/s5.py:82, :85 with reverse)
ffi::Error DiagScan(cudaStream_t stream, ffi::Buffer<ffi::C64> lam, ffi::Buffer<ffi::C64> bu, ffi::ResultBuffer<ffi::C64> h, int32_t reverse) {
  auto d = bu.dimensions();
  if (d.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_diag_scan: Bu must be (B,T,P)");
  return Check(eigb200_diag_scan(stream, reinterpret_cast<const float*>(lam.typed_data()), reinterpret_cast<const float*>(bu.typed_data()),
                                 reinterpret_cast<float*>(h->typed_data()), d[0], d[1], static_cast<int>(d[2]), reverse), "eigb200_diag_scan");
}

// discrete_DPLR's A-bar for a batch of (Lambda, P, Q, step)  (eval_eig.py:254-274)
ffi::Error DplrAbar(cudaStream_t stream, ffi::Buffer<ffi::C64> lambda, ffi::Buffer<ffi::C64> p, ffi::Buffer<ffi::C64> q, ffi::Buffer<ffi::F32> step,
                    ffi::ResultBuffer<ffi::C64> abar) {
  auto d = lambda.dimensions();
  if (d.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_dplr_abar: Lambda must be (nmat,N)");
  return Check(eigb200_dplr_abar(stream, reinterpret_cast<const float*>(lambda.typed_data()), reinterpret_cast<const float*>(p.typed_data()),
                                 reinterpret_cast<const float*>(q.typed_data()), step.typed_data(), d[0], static_cast<int>(d[1]),
                                 reinterpret_cast<float*>(abar->typed_data())), "eigb200_dplr_abar");
}

// np.linalg.eigvals for a batch of complex64 matrices, N <= 64 (eval_eig.py:296); `a` is consumed (declare it donated / aliased to a scratch result)
ffi::Error EigvalsC64(cudaStream_t stream, ffi::Buffer<ffi::C64> a, ffi::ResultBuffer<ffi::C64> scratch, ffi::ResultBuffer<ffi::C64> eig, ffi::ResultBuffer<ffi::S32> info) {
  auto d = a.dimensions();
  if (d.size() != 3 || d[1] != d[2]) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_eigvals_c64: A must be (nmat,N,N)");
  const size_t bytes = static_cast<size_t>(a.element_count()) * sizeof(std::complex<float>);
  if (cudaMemcpyAsync(scratch->typed_data(), a.typed_data(), bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "eigb200_eigvals_c64: device copy failed");
  return Check(eigb200_eigvals_c64(stream, reinterpret_cast<float*>(scratch->typed_data()), d[0], static_cast<int>(d[1]),
                                   reinterpret_cast<float*>(eig->typed_data()), info->typed_data()), "eigb200_eigvals_c64");
}

// threshold_analysis (eval_eig.py:335-362): closed-interval bin counts of a (B,N,inner) array; counts (B,inner,8) must arrive zeroed
ffi::Error ThresholdCounts(cudaStream_t stream, ffi::Buffer<ffi::F32> values, ffi::ResultBuffer<ffi::S32> counts) {
  auto d = values.dimensions();
  if (d.size() < 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_threshold_counts: values must be (B,N,...)");
  int64_t inner = 1;
  for (size_t i = 2; i < d.size(); ++i) inner *= d[i];
  return Check(eigb200_ratio_hist(stream, values.typed_data(), EIGB200_F32, EIGB200_RATIO_NONE, d[0], d[1], inner, nullptr, 1, counts->typed_data(),
                                  kThrRadius, 6, EIGB200_CMP_F64), "eigb200_ratio_hist");
}

// sum_b c, sum_b c^2 of a whole pass's counts (L,B,inner,8) -> (2,L,inner,8) int64: the buffer the single all-reduce carries (eval_eig.py:620-623)
ffi::Error CountMoments(cudaStream_t stream, ffi::Buffer<ffi::S32> counts, ffi::ResultBuffer<ffi::S64> moments) {
  auto d = counts.dimensions();
  if (d.size() < 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "eigb200_count_moments: counts must be (L,B,...,8)");
  int64_t inner = 1;
  for (size_t i = 2; i + 1 < d.size(); ++i) inner *= d[i];
  int64_t* m = moments->typed_data();
  return Check(eigb200_count_moments_layers(stream, counts.typed_data(), d[0], d[1], inner, m, m + d[0] * inner * EIGB200_NSLOT), "eigb200_count_moments_layers");
}
}  // namespace

#define EIGB_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_mamba2_eig, Mamba2Eig,
                              EIGB_STREAM.Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_ssm_lambda, SsmLambda,
                              EIGB_STREAM.Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::C64>>().Attr<int32_t>("kind"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_diag_scan, DiagScan,
                              EIGB_STREAM.Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::C64>>().Attr<int32_t>("reverse"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_dplr_abar, DplrAbar,
                              EIGB_STREAM.Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_eigvals_c64, EigvalsC64,
                              EIGB_STREAM.Arg<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::S32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_threshold_counts, ThresholdCounts, EIGB_STREAM.Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(eigb200_ffi_count_moments, CountMoments, EIGB_STREAM.Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S64>>());
#undef EIGB_STREAM

#else   // no XLA FFI headers on this machine: the translation unit is empty on purpose (see the header comment)
extern "C" int eigb200_ffi_unavailable(void) { return 1; }
#endif
