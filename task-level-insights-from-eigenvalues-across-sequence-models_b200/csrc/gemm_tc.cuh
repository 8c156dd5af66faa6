// gemm_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) GEMM, k4_gemm_tc.cu.
#pragma once
#include "gemm_simt.cuh"
#include <cuda.h>
namespace eigb200 {
size_t tc_workspace_bytes(int N, int K);                 // resident-weight kernels only (K <= 256)
size_t tc_workspace_bytes_m(int64_t M, int N, int K);    // any supported shape: adds the split copy of A the streamed-operand kernel needs
bool tc_supported(const LinearParams& p);
// nsplit = 3: error-compensated split (fp32-level accuracy); nsplit = 1: one rounded operand (plain TF32 / fp16).
// kind = 0: kind::tf32 operands (3xTF32), 1: kind::f16 operands (scaled fp16 split, twice the tensor rate; resident-weight kernels only).
// p.W == nullptr: the workspace already holds the prepared weights (tc_prepare / eigb200_linear_prepare) of this (N, K, epilogue, LayerNorm, kind).
// Default: the fp16 split.  Same accuracy class as 3xTF32 in every test (tests parametrise both) and in tools/split_error_study.py; measured on B200 at
// BASELINE C2: GLU GEMM 0.807 -> 0.731 ms, pass 11.6 -> 11.0 ms, and it is what lets the out_proj -> GLU tail run as one kernel (10.2 ms).
#ifndef EIGB200_GEMM_DEFAULT_KIND
#define EIGB200_GEMM_DEFAULT_KIND 1
#endif
void tc_set_default_kind(int kind);                    // 0 / 1 overrides the environment, anything else restores it
int tc_default_kind();                                   // EIGB200_GEMM_PRECISION = tf32x3 | f16x3, else EIGB200_GEMM_DEFAULT_KIND
int launch_linear_tc(cudaStream_t st, const LinearParams& p, int nsplit, void* workspace, int kind);
int tc_prepare(cudaStream_t st, const LinearParams& p, void* workspace, int kind);
int tc_overflow_query(cudaStream_t st, int reset, int* h_flag);
// shared with k4_gemm_fused.cu: where eigb200_linear_prepare put the fp16-split operands of (N, K, epilogue), tensor maps, the overflow flag
struct TcPrepared { const void* w_hi; const void* w_lo; const float* bias2; const float* scal; int bn, bg, nsplit, kp64, kch_w, wrows; };
bool tc_prepared_layout_f16(int N, int K, int epilogue, const void* ws, TcPrepared* out);
int tc_make_tmap_f32(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int tc_make_tmap_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
int* tc_overflow_flag();
// GELU(A W1^T + b1) -> GLU(. W2^T + b2) + R with the extractor partials, one kernel (the intermediate never leaves the SM); fp16 split, prepared operands
int launch_out_glu_fused(cudaStream_t st, const float* A, int64_t lda, const void* ws1, const float* bias1, const void* ws2, const float* bias2,
                         float* C, int64_t ldc, const float* R, int64_t ldr, int64_t M, int D, int K1, const float* eig_w, float* eig_part);
bool out_glu_fused_supported(int D, int K1);
// the same tail in the transposed orientation (k8_tail_fused_t.cu: weights = A operand from TMEM, 32-token chunks = B operand, three chunks in flight per CTA)
int launch_out_glu_fused_t(cudaStream_t st, const float* A, int64_t lda, const void* ws1, const float* bias1, const void* ws2, const float* bias2,
                           float* C, int64_t ldc, const float* R, int64_t ldr, int64_t M, int D, int K1, const float* eig_w, float* eig_part);
bool out_glu_fused_t_supported(int D, int K1);
}  // namespace eigb200
