/*
 * eigb200.h -- C ABI of libeigb200.so: the B200 (sm_100a) implementation of the eigenvalue-analysis hot path of
 * "Task-Level Insights from Eigenvalues across Sequence Models" (reference: analysis/eval_eig.py + models/).
 *
 * The reference has no FFI of its own (it is pure Python); each entry point below replaces ONE operator
 * boundary of the reference and cites it.  Conventions, identical for every function:
 *
 *   - extern "C", plain pointers and sizes, no framework types.  `stream` is a cudaStream_t passed as void*.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller; nothing is allocated or freed here;
 *     scratch is passed in by the caller (sizes documented per function).
 *   - calls only ENQUEUE work on `stream` and return; no host synchronisation, no hidden streams.
 *   - return value: 0 on success, negative EIGB200_E* on failure; eigb200_last_error() gives a thread-local message.
 *   - tensors are row-major and dense unless a row stride (`ld*`, in elements) is given.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with EIGB200_ECUDA.
 *
 * Bin-count layout shared by all histogram outputs ("counts"):  int32 [..., EIGB200_NSLOT = 8]
 *   slots 0 .. nthr   : the nthr+1 radius bins of threshold_analysis (analysis/eval_eig.py:335-362):
 *                       bin 0 = 0 <= v <= thr[0]; bin j = thr[j-1] <= v <= thr[j] (closed on both ends, so a value equal
 *                       to a threshold is counted twice); bin nthr = v > thr[nthr-1].   nthr <= 6.
 *   slot 7            : number of values whose PHASE falls in the first phase bin [0, thr_phase[0]]: for the real
 *                       non-negative Mamba-2 eigenvalues that is "not NaN" (eval_eig.py:614-618), for attention eta it is
 *                       "finite" because the reference bins 0*eta (eval_eig.py:673-674).
 *   counts are ACCUMULATED with integer atomics: zero the buffer first (eigb200_zero_i32 or cudaMemsetAsync).
 *
 * compare_mode selects the NumPy promotion the threshold compare reproduces:
 *   EIGB200_CMP_F64 (NumPy >= 2: value widened to float64)  |  EIGB200_CMP_F32 (pinned numpy 1.24.1: threshold rounded
 *   to the value's float32).  They differ only for values within 1 ulp of an edge.
 */
#ifndef EIGB200_H_
#define EIGB200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EIGB200_VERSION 100
#define EIGB200_NSLOT 8

enum { EIGB200_OK = 0, EIGB200_EINVAL = -1, EIGB200_ECUDA = -2, EIGB200_EUNSUPPORTED = -3 };
enum { EIGB200_CMP_F64 = 0, EIGB200_CMP_F32 = 1 };
/* norm_fn of get_eig_att_norm (analysis/eval_eig.py:142-151) */
enum { EIGB200_NORM_EXP = 0, EIGB200_NORM_ELU = 1, EIGB200_NORM_SOFTPLUS = 2, EIGB200_NORM_SIGMOID = 3 };
/* ratio orientation for eigb200_ratio_hist */
enum { EIGB200_RATIO_NONE = 0, EIGB200_RATIO_NEXT_OVER_CUR = 1, EIGB200_RATIO_CUR_OVER_NEXT = 2 };
/* element types */
enum { EIGB200_F32 = 0, EIGB200_F64 = 1, EIGB200_BF16 = 2 };
/* epilogues of eigb200_linear */
enum { EIGB200_EPI_NONE = 0, EIGB200_EPI_GELU = 1, EIGB200_EPI_GLU_RESIDUAL = 2, EIGB200_EPI_RESIDUAL = 3 };
/* math mode of eigb200_linear */
enum { EIGB200_GEMM_AUTO = 0, EIGB200_GEMM_SIMT_F32 = 1, EIGB200_GEMM_TC_3XTF32 = 2, EIGB200_GEMM_TC_TF32 = 3, EIGB200_GEMM_TC_F16X3 = 4 };

int         eigb200_version(void);
const char* eigb200_last_error(void);
/* sm_count / cc of device `dev`; EIGB200_ECUDA if there is no CUDA device. */
int         eigb200_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor);
/* cudaSetDevice for this library's (statically linked) CUDA runtime; call it from the thread that launches. */
int         eigb200_set_device(int dev);
int         eigb200_zero_i32(void* stream, int32_t* d_buf, size_t n);

/* ---- K1: fused gate-projection -> discretisation -> eigenvalue -> bin counts --------------------------------------
 * Replaces get_eig_mamba2(x, layer) (analysis/eval_eig.py:176-190) + the radius/phase threshold_analysis of its
 * output (:605-618, :335-362).  lambda[b,t,h] = exp(softplus(x[b,t,:] . W_dt[h,:] + dt_bias[h]) * -exp(A_log[h])).
 *   d_x      (B,T,D) float32 (x_dtype EIGB200_F32) or bfloat16 (EIGB200_BF16), D % 4 == 0 (bf16: D % 8 == 0), 16-byte aligned
 *   d_W_dt   (H,D)   float32: rows [d_inner + 2*ngroups*d_state, +nheads) of in_proj.weight (models/mamba.py:62-64)
 *   d_lam    float32 out, may be NULL (statistics only); entry (b,t,h) is written at d_lam[((b*T + t)*H + h) * lam_stride], so
 *            lam_stride = 1 gives a dense (B,T,H) array and lam_stride = L with d_lam offset by the layer index writes straight
 *            into the reference's (B,T,H,L) layout (np.concatenate(..., axis=-1), eval_eig.py:524-526)
 *   d_counts (B,H,8) int32 accumulated, may be NULL
 *   d_rowstats (B,T,2) float32 out, may be NULL: (mean, 1/sqrt(var + ln_eps)) of every row of x -- the LayerNorm statistics of the
 *            NEXT block's prenorm (models/mamba.py:329-331), produced for free while x streams through the extractor and consumed by
 *            eigb200_linear_ln so that the normalised activations never exist in HBM */
int eigb200_mamba2_eig(void* stream, const void* d_x, int x_dtype, int64_t B, int64_t T, int D,
                       const float* d_W_dt, const float* d_dt_bias, const float* d_A_log, int H,
                       float* d_lam, int64_t lam_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode,
                       float* d_rowstats, float ln_eps);

/* The same extractor without re-reading x: eigb200_linear_glu_extract (the GLU + residual GEMM that PRODUCES x, models/mamba.py:335-337) leaves, per row
 * and per group of 16 output columns, (x . W_dt, mean, M2) in d_partials[(g * 3 + c) * B*T + row]; this call combines the D / 16 = ngroups16 groups in a
 * fixed order and finishes lambda, the bin counts and the next block's LayerNorm statistics.  One head (H = 1); d_dt_bias / d_A_log are device pointers. */
int eigb200_mamba2_eig_partials(void* stream, const float* d_partials, int ngroups16, int64_t B, int64_t T,
                                const float* d_dt_bias, const float* d_A_log, float* d_lam, int64_t lam_stride, int32_t* d_counts,
                                const double* thresholds, int nthr, int compare_mode, float* d_rowstats, float ln_eps);

/* get_eig_mamba2_LTI (analysis/eval_eig.py:192-205): lambda[h] = exp(beta[h] * -softplus(A[h])), broadcast over (B,T).
 * d_lam (B,T,H) may be NULL; d_counts (B,H,8) accumulated. */
int eigb200_mamba2_lti_eig(void* stream, const float* d_A, const float* d_beta, int64_t B, int64_t T, int H,
                           float* d_lam, int64_t lam_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode);

/* ---- K1': normalised-attention gate  n[b,t,h] = exp(-norm_fn(x . W_n[h] + b_n[h] (+ offset[h])))  in float32 --------
 * First half of get_eig_att_norm (analysis/eval_eig.py:154-163); W_n/b_n are rows [D + 2*d_qk, +H) of Wvqkn
 * (models/norm_attention.py:201-203, :233-235); d_offset may be NULL.  d_n (B,T,H) float32 out.
 * Follow with eigb200_ratio_hist(..., EIGB200_F32, EIGB200_RATIO_NEXT_OVER_CUR) for eta (:165-169). */
int eigb200_normattn_gate(void* stream, const void* d_x, int x_dtype, int64_t B, int64_t T, int D,
                          const float* d_W_n, const float* d_b_n, const float* d_offset, int H, int norm_fn,
                          float* d_n);

/* ---- K1'': linear-attention normaliser  nu[b,t,h] = phi(q_t) . sum_{s<=t} phi(k_s),  phi = elu + 1 -------------------
 * O(T) form of the (B,T,T,H) score tensor of get_eig_att_linear (analysis/eval_eig.py:109-126); prefix sum and dot product
 * in float64.  q/k: raw projections, element (b,t,h,i) at d_q[(b*T + t)*ld + h*d + i] (so they can live inside the
 * Wqkv output, q at column 0 and k at column d_qk).  d_nu (B,T,H) float64 out.
 * Follow with eigb200_ratio_hist(..., EIGB200_F64, EIGB200_RATIO_CUR_OVER_NEXT) (:127-130). */
int eigb200_linattn_nu(void* stream, const float* d_q, const float* d_k, int64_t ld, int64_t B, int64_t T, int H, int d,
                       double* d_nu);

/* Softmax-attention normaliser with the reference's multiplicative-mask quirk (analysis/eval_eig.py:57-90):
 * m_t = max(max_{s<=t} q_t.k_s, 0 if t<T-1);  nu_t = sum_{s<=t} exp(q_t.k_s - m_t) + (T-1-t).  Streaming, no (B,T,T,H) tensor.
 * d_nu (B,T,H) float64, d_m (B,T,H) float32.  eta_t = nu_t/nu_{t+1} * exp(m_t - m_{t+1}) is formed by eigb200_softmax_eta. */
int eigb200_softmax_nu(void* stream, const float* d_q, const float* d_k, int64_t ld, int64_t B, int64_t T, int H, int d,
                       double* d_nu, float* d_m);
int eigb200_softmax_eta(void* stream, const double* d_nu, const float* d_m, int64_t B, int64_t T, int H,
                        double* d_eta, int32_t* d_counts, const double* thresholds, int nthr);
/* SelfAttention.forward (models/attention.py:14-35), causal softmax attention in float32 without the (B,H,T,T) score tensor:
 * out[b,t,h,:] = sum_{s<=t} softmax_s(q_t . (k_s * scale)) v_s.  q/k/v live in one projection buffer with row stride ld (element (b,t,h,i) of q at
 * d_q[(b*T+t)*ld + h*d + i], v at d_v[(b*T+t)*ld + h*dv + i]); out row stride ldo.  The reference's additive -10000 mask equals an exact mask
 * unless a row's scores fall below about -9900.  dv in {16,32,64,128}. */
int eigb200_softmax_attn_forward(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld, float scale,
                                 float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv);

/* ---- ratios + threshold statistics ----------------------------------------------------------------------------------
 * a: (B,N,inner) of `dtype` (EIGB200_F32 / EIGB200_F64).  mode NONE: v = a[b,n,i] (N values per (b,i));
 * NEXT_OVER_CUR: v = z(a[b,n+1,i]) / z(a[b,n,i]); CUR_OVER_NEXT: v = z(a[b,n,i]) / z(a[b,n+1,i])  (N-1 values),
 * z(u) = 2e-23 if u == 0 else u (analysis/eval_eig.py:127, :167), computed in float64.
 * d_out: values written as float64 (ratio modes), may be NULL; entry (b,n,i) at d_out[((b*(N-1) + n)*inner + i) * out_stride]
 * (out_stride = 1: dense (B,N-1,inner)).   d_counts (B,inner,8) accumulated.
 * With mode NONE this is threshold_analysis itself (analysis/eval_eig.py:335-362) for a (B,N,H*L) array. */
int eigb200_ratio_hist(void* stream, const void* d_a, int dtype, int mode, int64_t B, int64_t N, int64_t inner,
                       double* d_out, int64_t out_stride, int32_t* d_counts, const double* thresholds, int nthr, int compare_mode);

/* Batch moments of the bin counts: sum_b c and sum_b c^2 (int64) -- the only cross-sample (and cross-GPU) quantities
 * behind np.mean / np.std over the batch axis (analysis/eval_eig.py:620-623, :677-680).  d_counts (B,inner,8);
 * d_sum, d_sumsq (inner,8) int64, overwritten. */
int eigb200_count_moments(void* stream, const int32_t* d_counts, int64_t B, int64_t inner, int64_t* d_sum, int64_t* d_sumsq);
/* The same for the (L,B,inner,8) count buffer of a whole pass in ONE launch: d_sum, d_sumsq (L,inner,8) int64, overwritten.  This 2 x L x inner x 8
 * int64 buffer is everything a rank contributes to the single all-reduce of the path (eigb200_stats_allreduce below). */
int eigb200_count_moments_layers(void* stream, const int32_t* d_counts, int64_t L, int64_t B, int64_t inner, int64_t* d_sum, int64_t* d_sumsq);

/* Log-spaced histogram / quantiles of a (B,N,inner) array per `inner` column (the finer, on-device view of the eigenvalue radii per layer / head / state that
 * the north star asks for; the reference itself only has the fixed threshold bins above).  d_hist (inner, nbins + 3) int64, ACCUMULATED (zero it first; sum it
 * over batches / GPUs with eigb200_stats_allreduce): slot 0 = v < lo (incl. v <= 0), slots 1..nbins = log-spaced bins of [lo, hi), slot nbins + 1 = v >= hi,
 * slot nbins + 2 = NaN.  eigb200_hist_quantiles: d_q (nq) doubles in [0,1] (device) -> d_out (inner, nq) float64, log-linear interpolation inside a bin (NaN
 * entries excluded; values outside [lo, hi) report the range edge). */
int eigb200_log_hist(void* stream, const void* d_values, int dtype, int64_t B, int64_t N, int64_t inner, double lo, double hi, int nbins, int64_t* d_hist);
int eigb200_hist_quantiles(void* stream, const int64_t* d_hist, int64_t inner, double lo, double hi, int nbins, const double* d_q, int nq, double* d_out);

/* The ONE exchange step of the path: every GPU holds the moments (eigb200_count_moments_layers: 2 x L x inner x 8 int64) of ITS slice of the analysis
 * batch; their sum over the GPUs is all that np.mean / np.std over the batch axis need (analysis/eval_eig.py:620-623).  Integer sums: order independent, the
 * combined statistics are bit-reproducible.  eigb200_stats_allreduce sums d_moments in place over the ranks of `comm` (an ncclComm_t: one process per GPU as
 * torch.distributed / the reference's launcher creates it, or from eigb200_stats_comm_init_all for ONE process that drives several GPUs; then bracket the per-GPU
 * calls with eigb200_stats_group_start / _end).  Stream-ordered on `stream` of the current device.  NCCL is dlopen'ed at run time (the copy already in the process,
 * else $EIGB200_NCCL_LIB, else libnccl.so.2); EIGB200_EUNSUPPORTED when there is none (eigb200_stats_available() == 0). */
int eigb200_stats_available(void);
int eigb200_stats_comm_init_all(int ndev, const int* devs, void** comms);
int eigb200_stats_comm_destroy(void* comm);
int eigb200_stats_group_start(void);
int eigb200_stats_group_end(void);
int eigb200_stats_allreduce(void* stream, void* comm, int64_t* d_moments, size_t count);

/* ---- K2a: diagonal complex recurrence  h_t = lam * h_{t-1} + Bu_t  ---------------------------------------------------
 * What jax.lax.associative_scan(binary_operator_diag, (Lambda_elements, Bu_elements)) evaluates (models/lru.py:14-19, :95;
 * models/s5.py:51-62, :82, :85 with reverse).  complex64 as interleaved (re,im) float pairs.
 * d_lam (P), d_Bu (B,T,P), d_h (B,T,P); reverse != 0 runs t = T-1..0. */
int eigb200_diag_scan(void* stream, const float* d_lam, const float* d_Bu, float* d_h, int64_t B, int64_t T, int P, int reverse);

/* Parameter-only eigenvalues of the diagonal SSMs, get_eigvals_ssm("lru"|"s5") (analysis/eval_eig.py:303-329) and the
 * discretisations of models/s5.py:16-47.  kind 0 LRU: (p0,p1) = (nu_log, theta_log); kind 1 S5 zero-order hold, kind 2 S5 bilinear:
 * (p0,p1,p2) = (Lambda_re, Lambda_im, log_step).  All (P) float32; d_lam (P) complex64. */
int eigb200_ssm_lambda(void* stream, int kind, const float* d_p0, const float* d_p1, const float* d_p2, int P, float* d_lam);

/* ---- K2b: SSD selective scan --------------------------------------------------------------------------------------------
 * mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=D, z=None) at its call site models/mamba.py:138-150:
 *   h_t[h,p,n] = exp(dt_t[h] A[h]) h_{t-1}[h,p,n] + dt_t[h] B_t[g,n] x_t[h,p];   y_t[h,p] = sum_n C_t[g,n] h_t[h,p,n] + D[h] x_t[h,p]
 * x (B,T,H*P) row stride ldx; dt (B,T,H) (already softplus'ed); Bm, Cm (B,T,G*N) row stride ldbc; y (B,T,H*P) row stride ldy.
 * d_final_state (B,H,P,N) may be NULL. */
int eigb200_ssd_scan(void* stream, const float* d_x, int64_t ldx, const float* d_dt, const float* d_A,
                     const float* d_Bm, const float* d_Cm, int64_t ldbc, const float* d_D,
                     float* d_y, int64_t ldy, float* d_final_state,
                     int64_t B, int64_t T, int H, int P, int G, int N);

/* Fused depthwise causal conv (k taps, padding k-1, truncate) + SiLU on xBC, softplus(dt + dt_bias), then the SSD scan, reading
 * the in_proj output once (models/mamba.py:123-150).  d_xbcdt (B,T,ldz) = [x (H*P) | B (G*N) | C (G*N) | dt (H)];
 * d_conv_w (H*P + 2*G*N, k), d_conv_b (H*P + 2*G*N); kconv <= 4 (0 = no conv).  y (B,T,H*P) row stride ldy. */
int eigb200_mamba_conv_ssd(void* stream, const float* d_xbcdt, int64_t ldz, const float* d_conv_w, const float* d_conv_b, int kconv,
                           const float* d_dt_bias, const float* d_A_log, const float* d_D,
                           float* d_y, int64_t ldy, int64_t B, int64_t T, int H, int P, int G, int N);

/* ---- K3: S4 DPLR discretisation + batched nonsymmetric eigenvalues ---------------------------------------------------
 * eigb200_dplr_abar: A-bar of discrete_DPLR (analysis/eval_eig.py:254-274; models/s4.py:16-36) for nmat parameter sets:
 * d_Lambda, d_P, d_Q (nmat,N) complex64, d_step (nmat) float32 -> d_Abar (nmat,N,N) complex64 row-major.  N <= 64.
 * eigb200_eigvals_c64: np.linalg.eigvals(Ad) (analysis/eval_eig.py:296) for nmat dense complex64 N x N matrices (N <= 64), one warp
 * per matrix: Hessenberg reduction + shifted QR.  d_A is overwritten.  d_eig (nmat,N) complex64 (unordered, like LAPACK);
 * d_info (nmat) int32: 0 = converged, >0 = number of eigenvalues not converged. */
int eigb200_dplr_abar(void* stream, const float* d_Lambda, const float* d_P, const float* d_Q, const float* d_step,
                      int64_t nmat, int N, float* d_Abar);
int eigb200_eigvals_c64(void* stream, float* d_A, int64_t nmat, int N, float* d_eig, int32_t* d_info);

/* ---- K4 + block glue: what propagates activations from one layer to the next ---------------------------------------
 * eigb200_linear: C = epilogue(A W^T + bias) -- nn.Linear (models/mamba.py:64,109; common.py:53; attention.py:120; ...).
 *   A (M,K) row stride lda, W (N,K) dense (torch layout), bias (N) or NULL, C (M,Nout) row stride ldc.
 *   EPI_NONE: Nout = N.  EPI_GELU: exact-erf GELU (models/mamba.py:318, :333).  EPI_RESIDUAL: C = A W^T + bias + R.
 *   EPI_GLU_RESIDUAL: N = 2*Nout, C = z[:, :Nout] * sigmoid(z[:, Nout:]) + R  (GLU, models/common.py:55-58, + skip, mamba.py:337),
 *   R (M,Nout) row stride ldr (NULL = no residual).
 *   mode: SIMT_F32 = fp32 FFMA; TC_3XTF32 = tcgen05 tensor cores with the 3xTF32 split (fp32-level accuracy); AUTO picks.
 *   d_workspace/workspace_bytes: eigb200_linear_workspace_bytes(N, K) bytes (split weights for the tensor-core path). */
size_t eigb200_linear_workspace_bytes(int N, int K);
/* workspace for ANY shape the tensor-core path takes: equals the above when the weight slice stays resident in shared memory (K <= 256); for larger
 * K the streamed-operand kernel additionally keeps a tf32 hi/lo copy of A (2 * M * round_up(K,32) floats) there. */
size_t eigb200_linear_workspace_bytes_m(int64_t M, int N, int K);
/* Weights are constants of an analysis run (nn.Module parameters in eval mode, analysis/eval_eig.py:505-520): eigb200_linear_prepare splits W into the
 * tensor-core operand layout ONCE (tf32 hi / lo, rows in the order the CTAs consume them; with d_ln_gamma / d_ln_beta the LayerNorm scale is folded into
 * the weights and bias + W beta is stored behind them) into a workspace of eigb200_linear_workspace_bytes(N, K) bytes that the caller keeps per layer.
 * eigb200_linear / eigb200_linear_ln / eigb200_linear_glu_extract then take d_W = NULL with that workspace ("prepared": tensor-core path only, same N, K,
 * epilogue and LayerNorm as prepared; d_bias is still passed except for the LayerNorm form, whose folded bias lives in the workspace) and launch no
 * preparation kernels.  Shapes without a resident-weight plan (K > 256) return EIGB200_EUNSUPPORTED here and keep preparing per call. */
int eigb200_linear_prepare(void* stream, const float* d_W, const float* d_bias, const float* d_ln_gamma, const float* d_ln_beta,
                           int N, int K, int epilogue, void* d_workspace, size_t workspace_bytes);
/* eigb200_linear with nn.LayerNorm fused into the A operand: C = epilogue(LN(A) W^T + bias), LN(A)[m,k] = (A[m,k] - mean_m) * rstd_m *
 * gamma[k] + beta[k] with d_ln_stats (M,2) = (mean, rstd) per row (from eigb200_mamba2_eig / eigb200_embedding / eigb200_rowstats).
 * Tensor-core path only (same shape limits as eigb200_linear mode TC_3XTF32); EIGB200_EUNSUPPORTED otherwise. */
int eigb200_linear_ln(void* stream, const float* d_A, int64_t lda, const float* d_ln_stats, const float* d_ln_gamma, const float* d_ln_beta,
                      const float* d_W, const float* d_bias, float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                      int64_t M, int N, int K, int epilogue, void* d_workspace, size_t workspace_bytes);
/* (mean, 1/sqrt(var + eps)) of every row of x (rows, D) -> d_stats (rows, 2). */
int eigb200_rowstats(void* stream, const float* d_x, int64_t rows, int D, float eps, float* d_stats);
int eigb200_linear(void* stream, const float* d_A, int64_t lda, const float* d_W, const float* d_bias,
                   float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                   int64_t M, int N, int K, int epilogue, int mode, void* d_workspace, size_t workspace_bytes);
/* Operand precision of the tensor-core GEMMs.  The reference runs every nn.Linear as a full-fp32 cuBLAS SGEMM (torch default allow_tf32 = False), so both
 * tensor-core forms are error-compensated splits with fp32 accumulation and fp32-level accuracy:
 *   TC_3XTF32  a = hi + lo in tf32, 3 kind::tf32 MMAs per product;
 *   TC_F16X3   a S_a = hi + lo in fp16 with power-of-two scales (S_w from max |w| at preparation, S_a fixed), 3 kind::f16 MMAs per product at twice the
 *              tensor rate and half the operand footprint (resident-weight shapes, K <= 256).  An activation beyond 65504 / S_a (S_a = 16, or 1024 behind a
 *              fused LayerNorm) cannot be represented: the result holds inf / NaN there AND the sticky flag below is raised -- rerun with TC_3XTF32.
 * eigb200_gemm_precision(): 0 / 1 = what EIGB200_GEMM_AUTO, eigb200_linear_ln, eigb200_linear_glu_extract and eigb200_linear_prepare use (environment
 * EIGB200_GEMM_PRECISION = tf32x3 | f16x3).  eigb200_gemm_overflow: copies the flag of the current device to *h_flag (synchronises `stream`), optionally clears it. */
int eigb200_gemm_precision(void);
/* kind 0 (3xTF32) / 1 (fp16 split) overrides the environment for this process, any other value restores it.  Workspaces filled by eigb200_linear_prepare hold
 * the operands of the precision that was current when they were prepared: prepare again after switching. */
int eigb200_set_gemm_precision(int kind);
int eigb200_gemm_overflow(void* stream, int reset, int* h_flag);
/* GLU + residual GEMM (epilogue GLU_RESIDUAL) whose epilogue also emits the extractor partials of its OUTPUT rows for eigb200_mamba2_eig_partials:
 * d_W_gate (N/2) = the dt row of the block's in_proj, d_partials ((N/32) * 3 * M floats).  Tensor-core path only (K <= 256, N/2 % 16 == 0, R 32-byte aligned). */
int eigb200_linear_glu_extract(void* stream, const float* d_A, int64_t lda, const float* d_W, const float* d_bias,
                               float* d_C, int64_t ldc, const float* d_R, int64_t ldr, int64_t M, int N, int K,
                               const float* d_W_gate, float* d_partials, void* d_workspace, size_t workspace_bytes);

/* The tail of MambaBlock.forward (models/mamba.py:333-337) as ONE kernel:  C = GLU(GELU(y W_out^T + b_out) W_glu^T + b_glu) + R, optionally with the extractor
 * partials of the output rows (as eigb200_linear_glu_extract).  The intermediate GELU(out_proj(y)) never reaches HBM: the GELU epilogue of the first GEMM
 * writes it, split into fp16 hi / lo, into the shared-memory operand of the second.  d_ws_out / d_ws_glu: workspaces filled by eigb200_linear_prepare for
 * (W_out, EPI_GELU) and (W_glu, EPI_GLU_RESIDUAL) under the fp16-split precision (eigb200_gemm_precision() == 1; EIGB200_EINVAL otherwise).
 * Shapes: d_model D = 128, d_inner K1 a multiple of 32 <= 128 (eigb200_out_glu_fused_supported); y rows 16-byte, R / C rows 32-byte aligned.
 * d_W_gate / d_partials both NULL: no extractor partials.  Overflow of the fp16 range raises the flag of eigb200_gemm_overflow. */
int eigb200_out_glu_fused_supported(int D, int K1);
int eigb200_out_glu_fused(void* stream, const float* d_y, int64_t ldy, const void* d_ws_out, const float* d_bias_out,
                          const void* d_ws_glu, const float* d_bias_glu, float* d_C, int64_t ldc, const float* d_R, int64_t ldr,
                          int64_t M, int D, int K1, const float* d_W_gate, float* d_partials);

/* The FRONT of MambaBlock.forward as ONE kernel: prenorm LayerNorm (models/mamba.py:329-331) -> in_proj (:118) -> split [x | B | C | dt] -> conv1d + SiLU
 * (:123-127) -> softplus(dt + dt_bias), A = -exp(A_log) (:129-133) -> mamba_chunk_scan_combined(..., D=D) (:138-150).  The projection is computed transposed
 * on the tensor cores (weights = A operand, a 32-token chunk = B operand), so the accumulator in TMEM already has the thread = channel layout of the
 * recurrence and the in_proj output never reaches HBM.
 *   d_x (B*T, D) row stride ldx, 16-byte aligned rows; d_ln_stats (B*T, 2) = (mean, rstd) per row (as eigb200_linear_ln);
 *   d_ws_in: workspace filled by eigb200_linear_prepare(W_in, NULL, gamma, beta, N = d_inner + 2 N + 1, K = D, EPI_NONE) under the fp16-split precision;
 *   d_conv_w (d_inner + 2 N, kconv), d_conv_b (d_inner + 2 N), d_dt_bias (1), d_A_log (1), d_D (1) or NULL; d_y (B*T, d_inner) row stride ldy.
 * Shapes (eigb200_mamba_front_fused_supported): D = d_inner = 128, one head, one group, d_state 16, 1 <= kconv <= 4.  Any T >= 1. */
int eigb200_mamba_front_fused_supported(int D, int d_inner, int H, int G, int N, int kconv);
int eigb200_mamba_front_fused(void* stream, const float* d_x, int64_t ldx, const float* d_ln_stats, const void* d_ws_in,
                              const float* d_conv_w, const float* d_conv_b, int kconv, const float* d_dt_bias, const float* d_A_log,
                              const float* d_D, float* d_y, int64_t ldy, int64_t B, int64_t T, int D, int d_inner, int N);

/* TokenEmbeddings.forward (models/common.py:160-176): out[b,t,:] = word[ids[b,t],:] (+ pos[t,:] if d_pos != NULL). ids int64. */
int eigb200_embedding(void* stream, const int64_t* d_ids, const float* d_word, const float* d_pos, float* d_out,
                      int64_t B, int64_t T, int D, int64_t vocab);
/* same, also emitting the LayerNorm row statistics (mean, rstd) of the embedded rows into d_rowstats (B,T,2). */
int eigb200_embedding_stats(void* stream, const int64_t* d_ids, const float* d_word, const float* d_pos, float* d_out,
                            int64_t B, int64_t T, int D, int64_t vocab, float* d_rowstats, float ln_eps);
/* nn.LayerNorm over the last axis, eps inside the sqrt, biased variance (models/mamba.py:321; transformer.py:84). */
int eigb200_layernorm(void* stream, const float* d_x, const float* d_w, const float* d_b, float eps, float* d_out, int64_t rows, int D);
/* Depthwise causal conv1d (k taps, padding k-1, truncated to T) + SiLU over x (B,T,C) row stride ldx -> out row stride ldo
 * (models/attention.py:153-156; norm_attention.py:236-239).  k <= 8. */
int eigb200_conv_silu(void* stream, const float* d_x, int64_t ldx, const float* d_w, const float* d_b, int k,
                      float* d_out, int64_t ldo, int64_t B, int64_t T, int C);
/* Causal linear attention in O(T) form -- SelfLinAttention / SelfNormAttention (models/attention.py:63-83; norm_attention.py:61-89).
 * q,k (B,T,H,d), v (B,T,H,dv) inside one projection buffer of row stride ld; phi_elu != 0 applies elu+1 to q,k.
 * normalise: 1 = divide by phi(q_t).sum_s phi(k_s) (linear attention), 0 = multiply by d_gate[b,t,h] (norm attention, may be NULL).
 * kscale multiplies k (scale_B).  out (B,T,H*dv) row stride ldo. */
int eigb200_linattn_forward(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld,
                            const float* d_gate, int phi_elu, int normalise, float kscale,
                            float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv);
/* The same layer with the depthwise causal conv + SiLU of MHA.forward / MHNA.forward (models/attention.py:153-156; norm_attention.py:236-239) fused
 * into the tile loader: d_q / d_k / d_v point at the RAW projection columns, d_conv_w (C, kconv) / d_conv_b (C) are conv1d's parameters and conv_ch_* the
 * conv channel of each matrix's first column (head 0), < 0 = that matrix is not convolved (conv_type != "full" leaves v as it is).  Chunked tensor-core
 * kernel only: d = dv = 64, kconv <= 4, ld % 4 == 0, 16-byte aligned q / k / v, 8-byte aligned out -- eigb200_linattn_conv_fusable returns 1 when the call
 * is possible, otherwise run eigb200_conv_silu + eigb200_linattn_forward (EIGB200_EINVAL here). */
int eigb200_linattn_conv_fusable(const float* d_q, const float* d_k, const float* d_v, int64_t ld, const float* d_out, int64_t ldo, int d, int dv, int kconv);
int eigb200_linattn_forward_conv(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld,
                                 const float* d_gate, int phi_elu, int normalise, float kscale,
                                 const float* d_conv_w, const float* d_conv_b, int kconv, int conv_ch_q, int conv_ch_k, int conv_ch_v,
                                 float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv);
/* elementwise helpers of the transformer block (models/transformer.py:90-111): out = a + b; out = y * silu(z) etc. */
int eigb200_add(void* stream, const float* d_a, const float* d_b, float* d_out, int64_t n);
int eigb200_mul_silu(void* stream, const float* d_y, const float* d_z, float* d_out, int64_t n);
int eigb200_gelu(void* stream, const float* d_x, float* d_out, int64_t n);
/* out[m,j] = a[m,j] * s[j]: the D * u feed-through of LRU / S5 (models/lru.py:97; s5.py:247-248). */
int eigb200_scale_cols(void* stream, const float* d_a, const float* d_s, float* d_out, int64_t rows, int cols);
/* SSD_LTI.forward (models/mamba.py:262-281), in place on the conv'd projection buffer (rows, ld): the single dt column (ngroups = 1) becomes
 * dt[m, j] = softplus(buf[m, col_dt] + dt_bias[j / khead]) for the d_state columns j (khead = d_state / nheads, :200-201, :277-278) and B <- dt * B.
 * The scan then runs with dt := beta (ones) and A := -softplus(A) (:274-275, :283-295) through eigb200_ssd_scan. */
int eigb200_lti_scale_b(void* stream, float* d_buf, int64_t ld, int col_b, int col_dt, const float* d_dt_bias, int64_t rows, int N, int khead);

/* ---- S4 layer call, CNN mode (models/s4.py:50-79, :169-173) ------------------------------------------------------------
 * eigb200_s4_kernel: kernel_DPLR for H features at once.  Parameters (H,N) complex64 interleaved (Lambda already clipped to Re <= -1e-4, :116), step (H) =
 * exp(log_step); d_Kt (L,H) float32 out: the convolution kernel, lag-major so that the convolution reads it coalesced over features.
 * Workspace eigb200_s4_kernel_workspace_bytes(H, L) bytes (the generating function at the L roots of unity, complex128).
 * eigb200_s4_causal_conv: y[b,t,h] = sum_{s<=t} Kt[t-s,h] u[b,s,h] + D[h] u[b,t,h]  (causal_convolution(u, K) + D * u); u, y (B,T,H) float32, d_D nullable. */
size_t eigb200_s4_kernel_workspace_bytes(int H, int L);
int eigb200_s4_kernel(void* stream, const float* d_Lambda, const float* d_P, const float* d_Q, const float* d_B, const float* d_C,
                      const float* d_step, int H, int N, int L, float* d_Kt, void* d_workspace, size_t workspace_bytes);
int eigb200_s4_causal_conv(void* stream, const float* d_u, const float* d_Kt, const float* d_D, float* d_y, int64_t B, int64_t T, int H);

#ifdef __cplusplus
}
#endif
#endif /* EIGB200_H_ */
