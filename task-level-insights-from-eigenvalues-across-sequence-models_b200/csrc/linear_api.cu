// linear_api.cu -- eigb200_linear: argument checking and dispatch between the FFMA path and the tcgen05 path.
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
using namespace eigb200;

extern "C" size_t eigb200_linear_workspace_bytes(int N, int K) { return tc_workspace_bytes(N, K); }
extern "C" size_t eigb200_linear_workspace_bytes_m(int64_t M, int N, int K) { return tc_workspace_bytes_m(M, N, K); }

extern "C" int eigb200_linear(void* stream, const float* d_A, int64_t lda, const float* d_W, const float* d_bias,
                              float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                              int64_t M, int N, int K, int epilogue, int mode, void* d_workspace, size_t workspace_bytes) {
  EIGB_CHECK_ARG(d_A && d_C, "linear: null pointer");
  EIGB_CHECK_ARG(d_W || d_workspace, "linear: null weight pointer without a prepared workspace");
  EIGB_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  EIGB_CHECK_ARG(epilogue >= EIGB200_EPI_NONE && epilogue <= EIGB200_EPI_RESIDUAL, "linear: unknown epilogue %d", epilogue);
  EIGB_CHECK_ARG(epilogue != EIGB200_EPI_GLU_RESIDUAL || N % 2 == 0, "linear: GLU epilogue needs an even N");
  const int nout = epilogue == EIGB200_EPI_GLU_RESIDUAL ? N / 2 : N;
  EIGB_CHECK_ARG(lda >= K && ldc >= nout && (!d_R || ldr >= nout), "linear: row stride smaller than the row");
  LinearParams p{d_A, lda, d_W, d_bias, d_C, ldc, d_R, ldr, M, N, K, epilogue};
  cudaStream_t st = (cudaStream_t)stream;
  const bool prepared = d_W == nullptr;                              // workspace filled by eigb200_linear_prepare
  if (mode == EIGB200_GEMM_SIMT_F32) {
    EIGB_CHECK_ARG(!prepared, "linear: a prepared workspace serves the tensor-core path only");
    return launch_linear_simt(st, p);
  }
  const bool tc_ok = tc_supported(p) && d_workspace && workspace_bytes >= tc_workspace_bytes_m(M, N, K);
  if (mode == EIGB200_GEMM_TC_3XTF32 || mode == EIGB200_GEMM_TC_TF32 || mode == EIGB200_GEMM_TC_F16X3) {
    if (!tc_ok) { set_error("linear: shape/workspace not supported by the tensor-core path (M=%lld N=%d K=%d lda=%lld)", (long long)M, N, K, (long long)lda); return EIGB200_EUNSUPPORTED; }
    const int kind = mode == EIGB200_GEMM_TC_F16X3 ? 1 : 0;
    EIGB_CHECK_ARG(!prepared || kind == tc_default_kind(), "linear: a prepared workspace holds the operands of the default precision (EIGB200_GEMM_PRECISION), not of mode %d", mode);
    return launch_linear_tc(st, p, mode == EIGB200_GEMM_TC_TF32 ? 1 : 3, d_workspace, kind);
  }
  EIGB_CHECK_ARG(mode == EIGB200_GEMM_AUTO, "linear: unknown mode %d", mode);
  if (tc_ok && (M >= 1024 || prepared)) return launch_linear_tc(st, p, 3, d_workspace, tc_default_kind());
  EIGB_CHECK_ARG(!prepared, "linear: a prepared workspace serves the tensor-core path only (M=%lld N=%d K=%d)", (long long)M, N, K);
  return launch_linear_simt(st, p);
}

extern "C" int eigb200_linear_ln(void* stream, const float* d_A, int64_t lda, const float* d_ln_stats, const float* d_ln_gamma, const float* d_ln_beta,
                                 const float* d_W, const float* d_bias, float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                                 int64_t M, int N, int K, int epilogue, void* d_workspace, size_t workspace_bytes) {
  EIGB_CHECK_ARG(d_A && d_C && d_ln_stats, "linear_ln: null pointer");
  EIGB_CHECK_ARG((d_W && d_ln_gamma && d_ln_beta) || (!d_W && d_workspace), "linear_ln: weights / gamma / beta missing without a prepared workspace");
  EIGB_CHECK_ARG(M > 0 && N > 0 && K > 0, "linear_ln: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  EIGB_CHECK_ARG(epilogue >= EIGB200_EPI_NONE && epilogue <= EIGB200_EPI_RESIDUAL, "linear_ln: unknown epilogue %d", epilogue);
  EIGB_CHECK_ARG(epilogue != EIGB200_EPI_GLU_RESIDUAL || N % 2 == 0, "linear_ln: GLU epilogue needs an even N");
  EIGB_CHECK_ARG(((uintptr_t)d_ln_stats & 7) == 0, "linear_ln: row statistics must be 8-byte aligned");
  LinearParams p{d_A, lda, d_W, d_bias, d_C, ldc, d_R, ldr, M, N, K, epilogue};
  p.ln_stats = d_ln_stats; p.ln_gamma = d_ln_gamma; p.ln_beta = d_ln_beta;
  if (!(tc_supported(p) && d_workspace && workspace_bytes >= tc_workspace_bytes_m(M, N, K))) {
    set_error("linear_ln: shape/workspace not supported by the tensor-core path (M=%lld N=%d K=%d lda=%lld)", (long long)M, N, K, (long long)lda);
    return EIGB200_EUNSUPPORTED;
  }
  return launch_linear_tc((cudaStream_t)stream, p, 3, d_workspace, tc_default_kind());
}

extern "C" int eigb200_linear_glu_extract(void* stream, const float* d_A, int64_t lda, const float* d_W, const float* d_bias,
                                          float* d_C, int64_t ldc, const float* d_R, int64_t ldr, int64_t M, int N, int K,
                                          const float* d_W_gate, float* d_partials, void* d_workspace, size_t workspace_bytes) {
  EIGB_CHECK_ARG(d_A && d_C && d_R && d_W_gate && d_partials && (d_W || d_workspace), "linear_glu_extract: null pointer");
  EIGB_CHECK_ARG(M > 0 && N > 0 && N % 2 == 0 && K > 0, "linear_glu_extract: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  EIGB_CHECK_ARG(lda >= K && ldc >= N / 2 && ldr >= N / 2, "linear_glu_extract: row stride smaller than the row");
  LinearParams p{d_A, lda, d_W, d_bias, d_C, ldc, d_R, ldr, M, N, K, EIGB200_EPI_GLU_RESIDUAL};
  p.eig_w = d_W_gate; p.eig_part = d_partials;
  if (!(tc_supported(p) && d_workspace && workspace_bytes >= tc_workspace_bytes_m(M, N, K))) {
    set_error("linear_glu_extract: shape/workspace not supported by the tensor-core path (M=%lld N=%d K=%d)", (long long)M, N, K);
    return EIGB200_EUNSUPPORTED;
  }
  return launch_linear_tc((cudaStream_t)stream, p, 3, d_workspace, tc_default_kind());
}

extern "C" int eigb200_linear_prepare(void* stream, const float* d_W, const float* d_bias, const float* d_ln_gamma, const float* d_ln_beta,
                                      int N, int K, int epilogue, void* d_workspace, size_t workspace_bytes) {
  EIGB_CHECK_ARG(d_W && d_workspace, "linear_prepare: null pointer");
  EIGB_CHECK_ARG(N > 0 && K > 0, "linear_prepare: bad shape N=%d K=%d", N, K);
  EIGB_CHECK_ARG(epilogue >= EIGB200_EPI_NONE && epilogue <= EIGB200_EPI_RESIDUAL, "linear_prepare: unknown epilogue %d", epilogue);
  EIGB_CHECK_ARG(epilogue != EIGB200_EPI_GLU_RESIDUAL || N % 2 == 0, "linear_prepare: GLU epilogue needs an even N");
  EIGB_CHECK_ARG((d_ln_gamma == nullptr) == (d_ln_beta == nullptr), "linear_prepare: LayerNorm needs both gamma and beta");
  const size_t need = tc_workspace_bytes(N, K);
  if (need == 0 || workspace_bytes < need) {
    set_error("linear_prepare: N=%d K=%d has no resident-weight plan or the workspace is smaller than eigb200_linear_workspace_bytes", N, K);
    return EIGB200_EUNSUPPORTED;
  }
  LinearParams p{nullptr, K, d_W, d_bias, nullptr, N, nullptr, 0, 0, N, K, epilogue};
  p.ln_gamma = d_ln_gamma; p.ln_beta = d_ln_beta;
  return tc_prepare((cudaStream_t)stream, p, d_workspace, tc_default_kind());
}

extern "C" int eigb200_gemm_precision(void) { return tc_default_kind(); }
extern "C" int eigb200_set_gemm_precision(int kind) { tc_set_default_kind(kind); return EIGB200_OK; }

extern "C" int eigb200_gemm_overflow(void* stream, int reset, int* h_flag) {
  EIGB_CHECK_ARG(h_flag, "gemm_overflow: null pointer");
  return tc_overflow_query((cudaStream_t)stream, reset, h_flag);
}
