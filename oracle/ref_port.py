"""CPU oracle for the eigenvalue-analysis hot path.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Every function names the reference file:line it restates.  Two flavours exist where precision matters:
`dtype=np.float64` (the "truth" the CUDA kernels are toleranced against) and `dtype=np.float32`
(the reference's own working precision, used for the bin-count / underflow quirks).
Nothing here is copied from the reference; it is re-derived from the behaviour of the cited lines.
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    "softplus", "elu", "sigmoid", "norm_fn_apply",
    "mamba2_dt_rows", "mamba2_eig", "mamba2_lti_eig", "ssd_lti_mixer_forward", "normattn_rows", "normattn_eta",
    "linattn_qk", "linattn_eta_quadratic", "linattn_eta_prefix", "softmax_eta_quadratic", "softmax_eta_closed", "smattn_forward",
    "THRESHOLDS_RADIUS", "THRESHOLDS_PHASE", "threshold_counts", "threshold_analysis", "threshold_analysis_ssm",
    "radius_phase", "batch_mean_std", "batch_mean_std_from_counts",
    "lru_lambda", "s5_lambda", "diag_scan", "lru_forward", "s5_discretize", "s5_forward",
    "make_hippo", "make_nplr_hippo", "make_dplr_hippo", "discrete_dplr_abar", "s4_eigvals", "dplr_exact_spectrum",
    "s4_kernel_dplr", "s4_discrete_dplr", "s4_kernel_recurrent", "s4_forward",
    "ssd_scan_sequential", "ssd_scan_chunked", "layer_norm", "causal_depthwise_conv_silu", "gelu_erf", "glu",
    "ssd_mixer_forward", "mamba_block_forward", "token_embedding", "linattn_forward", "normattn_forward",
    "transformer_block_forward", "mamba_eval_pass", "transformer_eval_pass", "mamba_eval_pass_torch_cpu",
]

# --------------------------------------------------------------------------------------------------
# scalar activations with torch semantics
# --------------------------------------------------------------------------------------------------

def softplus(z):
    """torch.nn.functional.softplus(beta=1, threshold=20): identity above 20 (eval_eig.py:182, :147)."""
    z = np.asarray(z)
    return np.where(z > 20.0, z, np.log1p(np.exp(np.minimum(z, 20.0))))


def elu(z):
    """torch.nn.functional.elu(alpha=1) = expm1 for z<=0 (eval_eig.py:109-110, :145)."""
    z = np.asarray(z)
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def sigmoid(z):
    z = np.asarray(z)
    return 1.0 / (1.0 + np.exp(-z))


def norm_fn_apply(name: str, z):
    """norm_fn selection of get_eig_att_norm (eval_eig.py:142-151); unknown name -> RuntimeError."""
    if name == "exp":
        return np.exp(z)
    if name == "elu":
        return elu(z)
    if name == "softplus":
        return softplus(z)
    if name == "sigmoid":
        return sigmoid(z)
    raise RuntimeError("normalization function {0} not implemented!".format(name))


# --------------------------------------------------------------------------------------------------
# a1 / a2: Mamba-2 eigenvalues
# --------------------------------------------------------------------------------------------------

def mamba2_dt_rows(d_inner: int, ngroups: int, d_state: int, nheads: int):
    """Row slice of in_proj.weight that produces dt: split order [x | B | C | dt]
    (models/mamba.py:62-63, :123-125; eval_eig.py:179-181)."""
    lo = d_inner + 2 * ngroups * d_state
    return slice(lo, lo + nheads)


def mamba2_eig(x, in_proj_weight, dt_bias, A_log, d_inner, ngroups, d_state, nheads, dtype=np.float64):
    """get_eig_mamba2 (eval_eig.py:176-190): lambda = exp(softplus(x W_dt^T + dt_bias) * (-exp(A_log))).
    x (B,T,D) -> (B,T,H,1).  in_proj has no bias (models/mamba.py:40,64)."""
    x = np.asarray(x, dtype)
    W = np.asarray(in_proj_weight, dtype)[mamba2_dt_rows(d_inner, ngroups, d_state, nheads)]
    z = x @ W.T + np.asarray(dt_bias, dtype)
    dt = softplus(z).astype(dtype)
    A = -np.exp(np.asarray(A_log, dtype))
    lam = np.exp(dt * A).astype(dtype)
    return lam[..., None]


def mamba2_lti_eig(batch, seqlen, A_param, beta, dtype=np.float64):
    """get_eig_mamba2_LTI (eval_eig.py:192-205): lambda = exp(beta * (-softplus(A))) broadcast over (B,T)."""
    A = -softplus(np.asarray(A_param, dtype))
    lam = np.exp(np.asarray(beta, dtype) * A).astype(dtype)
    return np.broadcast_to(lam, (batch, seqlen, lam.shape[0])).copy()[..., None]


# --------------------------------------------------------------------------------------------------
# a4: normalised attention eta
# --------------------------------------------------------------------------------------------------

def normattn_rows(d_model: int, d_qk: int, num_heads: int):
    """Rows of Wvqkn producing n: split order [v | q | k | n] (norm_attention.py:233-235; eval_eig.py:156-158)."""
    lo = d_model + 2 * d_qk
    return slice(lo, lo + num_heads)


def normattn_eta(x, Wvqkn_weight, Wvqkn_bias, offset, norm_fn, d_model, d_qk, num_heads, dtype=np.float32):
    """get_eig_att_norm (eval_eig.py:137-174).  n is formed in the reference's working precision (fp32 by
    default here, because the `n == 0 -> 2e-23` patch at :167 acts on fp32 underflow), then cast to fp64,
    eta_t = n_{t+1} / n_t (:169).  offset may be None (:160-163).  Returns (B,T-1,H,1) float64."""
    x = np.asarray(x, dtype)
    rows = normattn_rows(d_model, d_qk, num_heads)
    W = np.asarray(Wvqkn_weight, dtype)[rows]
    b = np.asarray(Wvqkn_bias, dtype)[rows]
    raw = (x @ W.T + b).astype(dtype)
    if offset is not None:
        raw = (raw + np.asarray(offset, dtype)).astype(dtype)
    with np.errstate(over="ignore", under="ignore"):
        n = np.exp(-norm_fn_apply(norm_fn, raw).astype(dtype)).astype(dtype)
    n = n.astype(np.float64)
    n[n == 0.0] = 2e-23
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        eta = n[:, 1:, :] / n[:, :-1, :]
    return eta[..., None]


# --------------------------------------------------------------------------------------------------
# a3 / a5: linear- and softmax-attention eta
# --------------------------------------------------------------------------------------------------

def linattn_qk(x, Wqkv_weight, Wqkv_bias, d_qk, num_heads, dtype=np.float32):
    """q,k of get_eig_att_linear / _softmax (eval_eig.py:46-52, :99-108): rows [q (d_qk) | k (d_qk) | v],
    reshaped '(two h d)'.  Returns q,k of shape (B,T,H,d)."""
    x = np.asarray(x, dtype)
    W = np.asarray(Wqkv_weight, dtype)[: 2 * d_qk]
    b = np.asarray(Wqkv_bias, dtype)[: 2 * d_qk] if Wqkv_bias is not None else 0.0
    qk = (x @ W.T + b).astype(dtype)
    B, T, _ = qk.shape
    d = d_qk // num_heads
    qk = qk.reshape(B, T, 2, num_heads, d)
    return qk[:, :, 0], qk[:, :, 1]


def linattn_eta_quadratic(q, k, dtype=np.float32):
    """get_eig_att_linear (eval_eig.py:109-135), literal O(T^2) form: scores_{ts} = phi(q_t).phi(k_s) in the
    working precision, tril mask by multiplication, cast to fp64, nu_t = sum_s, nu==0 -> 2e-23,
    eta_t = nu_t / nu_{t+1}.  q,k (B,T,H,d) raw projections.  Returns (B,T-1,H,1) float64."""
    q = (elu(np.asarray(q, dtype)) + 1).astype(dtype)
    k = (elu(np.asarray(k, dtype)) + 1).astype(dtype)
    T = q.shape[1]
    scores = np.einsum("bthd,bshd->btsh", q, k).astype(dtype)
    mask = np.tril(np.ones((T, T), dtype))
    scores = scores * mask[None, :, :, None]
    scores = np.nan_to_num(scores.astype(np.float64))
    nu = scores.sum(axis=2)
    nu[nu == 0.0] = 2e-23
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        eta = nu[:, :-1, :] / nu[:, 1:, :]
    return eta[..., None]


def linattn_eta_prefix(q, k, dtype=np.float32):
    """O(T) restatement of the same quantity: nu_t = phi(q_t) . sum_{s<=t} phi(k_s) with the prefix sum and
    the dot product carried in fp64 (what the CUDA kernel does).  Equal to the quadratic form up to fp32
    rounding of the individual scores."""
    q = (elu(np.asarray(q, dtype)) + 1).astype(dtype).astype(np.float64)
    k = (elu(np.asarray(k, dtype)) + 1).astype(dtype).astype(np.float64)
    S = np.cumsum(k, axis=1)
    nu = np.einsum("bthd,bthd->bth", q, S)
    nu[nu == 0.0] = 2e-23
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        eta = nu[:, :-1, :] / nu[:, 1:, :]
    return eta[..., None]


def softmax_eta_quadratic(q, k, dtype=np.float32):
    """get_eig_att_softmax (eval_eig.py:43-95), literal form.  Masking is by multiplication, so masked logits
    are 0 (not -inf): the row max includes 0 for every row but the last, and each masked position contributes
    exp(0 - m_t)... after the second mask multiply the subtracted max is also zeroed there, i.e. exp(0)=1."""
    q = np.asarray(q, dtype)
    k = np.asarray(k, dtype)
    T = q.shape[1]
    scores = np.einsum("bthd,bshd->btsh", q, k).astype(dtype)
    mask = np.tril(np.ones((T, T), dtype))
    scores = (scores * mask[None, :, :, None]).astype(dtype)
    smax = scores.max(axis=2)                                   # (B,T,H) includes masked zeros
    smax_r = np.repeat(smax[:, :, None, :], T, axis=2) * mask[None, :, :, None]
    norm = (scores - smax_r.astype(dtype)).astype(dtype).astype(np.float64)
    with np.errstate(over="ignore"):
        e = np.nan_to_num(np.exp(norm))
    nu = e.sum(axis=2)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        eta = nu[:, :-1, :] / nu[:, 1:, :]
        diff = (-smax[:, 1:, :] + smax[:, :-1, :]).astype(dtype)
        eta = eta * np.exp(diff.astype(np.float64))
    return eta[..., None]


def softmax_eta_closed(q, k, dtype=np.float32):
    """Streaming closed form of the same quirk:
    m_t = max(max_{s<=t} s_ts, 0 if t<T-1), nu_t = sum_{s<=t} exp(s_ts - m_t) + (T-1-t),
    eta_t = nu_t/nu_{t+1} * exp(m_t - m_{t+1})."""
    q = np.asarray(q, dtype)
    k = np.asarray(k, dtype)
    B, T, H, _ = q.shape
    nu = np.empty((B, T, H))
    m = np.empty((B, T, H), dtype)
    for t in range(T):
        s = np.einsum("bhd,bshd->bsh", q[:, t], k[:, : t + 1]).astype(dtype)
        mt = s.max(axis=1)
        if t < T - 1:
            mt = np.maximum(mt, dtype(0))
        m[:, t] = mt
        d = (s - mt[:, None, :]).astype(dtype).astype(np.float64)
        nu[:, t] = np.exp(d).sum(axis=1) + (T - 1 - t)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        eta = nu[:, :-1] / nu[:, 1:] * np.exp((m[:, :-1] - m[:, 1:]).astype(dtype).astype(np.float64))
    return eta[..., None]


# --------------------------------------------------------------------------------------------------
# a9 / a10 / a11: threshold statistics
# --------------------------------------------------------------------------------------------------

THRESHOLDS_RADIUS = np.array([0.1, 0.5, 0.9, 1.0, 10, 100])     # eval_eig.py:603, :665, :724
THRESHOLDS_PHASE = np.array([1, 10, 45, 90, 180])               # eval_eig.py:612, :671, :732


def threshold_counts(values, thresholds, axis, compare="float64"):
    """Integer bin counts behind threshold_analysis (eval_eig.py:335-362).
    Bin 0: 0 <= v <= thr[0]; bin j (1..n-1): thr[j-1] <= v <= thr[j] (closed on BOTH ends, :350, :359);
    last bin: v > thr[-1] (:354).  NaN falls in no bin; negative values fall in no bin.
    compare="float64": NumPy >= 2 promotion (value widened to fp64); compare="float32": the pinned
    numpy 1.24.1 value-based casting, thresholds rounded to the value dtype."""
    v = np.asarray(values)
    thr = np.asarray(thresholds, np.float64).ravel()
    if compare == "float32" and v.dtype == np.float32:
        thr = thr.astype(np.float32)
    else:
        v = v.astype(np.float64)
    n = thr.shape[0]
    out = []
    with np.errstate(invalid="ignore"):
        out.append(((v >= 0) & (v <= thr[0])).sum(axis=axis))
        for j in range(n - 1):
            out.append(((v >= thr[j]) & (v <= thr[j + 1])).sum(axis=axis))
        out.append((v > thr[-1]).sum(axis=axis))
    return np.stack(out, axis=0).astype(np.int64)


def threshold_analysis(eig_val, thresholds, num_layers=None, num_heads=None, batch_size=None, compare="float64"):
    """threshold_analysis (eval_eig.py:335-362): eig_val (B,N,H,L) -> percentages (n_thr+1,B,H,L) float64,
    count / N * 100 (:351)."""
    eig_val = np.asarray(eig_val)
    counts = threshold_counts(eig_val, thresholds, axis=1, compare=compare)
    return counts / eig_val.shape[1] * 100


def threshold_analysis_ssm(eig_val, thresholds, num_layers=None, compare="float64"):
    """threshold_analysis_ssm (eval_eig.py:364-391): eig_val (P,L) -> (n_thr+1,L)."""
    eig_val = np.asarray(eig_val)
    counts = threshold_counts(eig_val, thresholds, axis=0, compare=compare)
    return counts / eig_val.shape[0] * 100


def radius_phase(eig):
    """|lambda| = sqrt(re^2 + im^2) in the array's own precision and arg in degrees
    (eval_eig.py:605-606, :614-615, :726-727, :734-735)."""
    eig = np.asarray(eig)
    rad = np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2))
    ph = np.arctan2(eig.imag, eig.real) * 180 / np.pi
    return rad, ph


def batch_mean_std(percentage):
    """np.mean / np.std over the batch axis, ddof 0 (eval_eig.py:620-623, :677-680)."""
    return np.mean(percentage, axis=1), np.std(percentage, axis=1)


def batch_mean_std_from_counts(sum_c, sum_c2, n_per_seq: int, batch: int):
    """The integer-moment form used across GPUs: mean = S1*100/(N*B),
    std = 100/N * sqrt(S2/B - (S1/B)^2) in float64."""
    s1 = np.asarray(sum_c, np.float64)
    s2 = np.asarray(sum_c2, np.float64)
    mean = s1 * 100.0 / (n_per_seq * batch)
    var = np.maximum(s2 / batch - (s1 / batch) ** 2, 0.0)
    return mean, 100.0 / n_per_seq * np.sqrt(var)


# --------------------------------------------------------------------------------------------------
# a6 / a7: LRU and S5 eigenvalues, diagonal scan, layer forward
# --------------------------------------------------------------------------------------------------

def lru_lambda(nu_log, theta_log, dtype=np.complex128):
    """lambda = exp(-exp(nu_log) + i exp(theta_log)) (eval_eig.py:321-324; models/lru.py:88)."""
    ft = np.float64 if dtype == np.complex128 else np.float32
    nu = np.exp(np.asarray(nu_log, ft))
    th = np.exp(np.asarray(theta_log, ft))
    return np.exp(-nu + 1j * th).astype(dtype)


def s5_lambda(Lambda_re, Lambda_im, log_step, dtype=np.complex128):
    """lambda_bar = exp((Lambda_re + i Lambda_im) * exp(log_step)) -- the analysis always uses ZOH and never
    clips (eval_eig.py:306-311)."""
    ft = np.float64 if dtype == np.complex128 else np.float32
    step = np.exp(np.asarray(log_step, ft).reshape(-1))
    lam = np.asarray(Lambda_re, ft) + 1j * np.asarray(Lambda_im, ft)
    return np.exp(lam * step).astype(dtype)


def diag_scan(lam, Bu, reverse=False, dtype=np.complex128):
    """h_t = lam * h_{t-1} + Bu_t, h_{-1} = 0: the recurrence that associative_scan(binary_operator_diag, ...)
    evaluates (models/lru.py:14-19, :95; models/s5.py:51-62, :82, :85).  lam (P,), Bu (..., T, P).
    reverse=True runs t = T-1 .. 0 (s5.py:85)."""
    lam = np.asarray(lam, dtype)
    Bu = np.asarray(Bu, dtype)
    h = np.empty_like(Bu)
    acc = np.zeros(Bu.shape[:-2] + Bu.shape[-1:], dtype)
    T = Bu.shape[-2]
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        acc = lam * acc + Bu[..., t, :]
        h[..., t, :] = acc
    return h


def lru_forward(params, u, dtype=np.complex128):
    """LRU.__call__ (models/lru.py:86-99) for u (B,T,d_model): returns (y, h, Bu)."""
    ft = np.float64 if dtype == np.complex128 else np.float32
    lam = lru_lambda(params["nu_log"], params["theta_log"], dtype)
    Bn = (np.asarray(params["B_re"], ft) + 1j * np.asarray(params["B_im"], ft)) * np.exp(np.asarray(params["gamma_log"], ft))[:, None]
    C = np.asarray(params["C_re"], ft) + 1j * np.asarray(params["C_im"], ft)
    u = np.asarray(u, ft)
    Bu = (u @ Bn.T).astype(dtype)
    h = diag_scan(lam, Bu, dtype=dtype)
    y = (h @ C.T).real + np.asarray(params["D"], ft) * u
    return y, h, Bu


def s5_discretize(Lambda, B_tilde, step, method="zoh"):
    """discretize_zoh / discretize_bilinear (models/s5.py:16-47)."""
    if method == "zoh":
        lam_bar = np.exp(Lambda * step)
        B_bar = (1.0 / Lambda * (lam_bar - 1.0))[:, None] * B_tilde
    elif method == "bilinear":
        BL = 1.0 / (1.0 - (step / 2.0) * Lambda)
        lam_bar = BL * (1.0 + (step / 2.0) * Lambda)
        B_bar = (BL * step)[:, None] * B_tilde
    else:
        raise NotImplementedError("Discretization method {} not implemented".format(method))
    return lam_bar, B_bar


def s5_forward(params, u, discretization="zoh", conj_sym=True, clip_eigs=False, bidirectional=False, dtype=np.complex128):
    """S5SSM.__call__ / apply_ssm (models/s5.py:65-93, :141-250) for u (B,T,H): returns (y, h, Bu)."""
    ft = np.float64 if dtype == np.complex128 else np.float32
    lre = np.asarray(params["Lambda_re"], ft)
    if clip_eigs:
        lre = np.minimum(lre, -1e-4)
    Lam = lre + 1j * np.asarray(params["Lambda_im"], ft)
    Bt = np.asarray(params["B"], ft)
    Bt = Bt[..., 0] + 1j * Bt[..., 1]
    if bidirectional and "C1" in params:
        C1 = np.asarray(params["C1"], ft); C2 = np.asarray(params["C2"], ft)
        Ct = np.concatenate((C1[..., 0] + 1j * C1[..., 1], C2[..., 0] + 1j * C2[..., 1]), axis=-1)
    else:
        Cc = np.asarray(params["C"], ft)
        Ct = Cc[..., 0] + 1j * Cc[..., 1]
    step = np.exp(np.asarray(params["log_step"], ft)[:, 0])
    lam_bar, B_bar = s5_discretize(Lam, Bt, step, discretization)
    u = np.asarray(u, ft)
    Bu = (u @ B_bar.T).astype(dtype)
    h = diag_scan(lam_bar, Bu, dtype=dtype)
    if bidirectional:
        h = np.concatenate((h, diag_scan(lam_bar, Bu, reverse=True, dtype=dtype)), axis=-1)
    y = (h @ Ct.T).real
    if conj_sym:
        y = 2 * y
    return y + np.asarray(params["D"], ft) * u, h, Bu


# --------------------------------------------------------------------------------------------------
# a8: S4 DPLR
# --------------------------------------------------------------------------------------------------

def make_hippo(N):
    """HiPPO-LegS matrix (models/common.py:180-191)."""
    p = np.sqrt(1 + 2 * np.arange(N))
    A = np.tril(p[:, None] * p[None, :]) - np.diag(np.arange(N))
    return -A


def make_nplr_hippo(N):
    """(models/common.py:193-212)."""
    return make_hippo(N), np.sqrt(np.arange(N) + 0.5), np.sqrt(2 * np.arange(N) + 1.0)


def make_dplr_hippo(N):
    """(models/common.py:215-241): Lambda, P, B, V, B_orig."""
    A, P, B = make_nplr_hippo(N)
    S = A + P[:, None] * P[None, :]
    lam_re = np.mean(np.diagonal(S)) * np.ones(N)
    lam_im, V = np.linalg.eigh(S * -1j)
    Pn = V.conj().T @ P
    Bn = V.conj().T @ B
    return lam_re + 1j * lam_im, Pn, Bn, V, B


def discrete_dplr_abar(Lambda, P, Q, step, dtype=np.complex128):
    """A-bar of discrete_DPLR (eval_eig.py:254-274; models/s4.py:16-36):
    A = diag(Lambda) - P Q^*;  A0 = (2/step) I + A;  D = diag(1/(2/step - Lambda));
    A1 = D - D P (1 + Q^* D P)^-1 Q^* D;  Abar = A1 A0.  (B-bar, C-bar are discarded by the analysis.)"""
    Lambda = np.asarray(Lambda, dtype); P = np.asarray(P, dtype); Q = np.asarray(Q, dtype)
    N = Lambda.shape[0]
    two_over = (2.0 / step)
    A = np.diag(Lambda) - np.outer(P, Q.conj())
    A0 = two_over * np.eye(N, dtype=dtype) + A
    d = 1.0 / (two_over - Lambda)
    qdp = np.sum(Q.conj() * d * P)
    A1 = np.diag(d) - np.outer(d * P, Q.conj() * d) / (1.0 + qdp)
    return (A1 @ A0).astype(dtype)


def s4_eigvals(layer, idx=1, dtype=np.complex128):
    """get_eigvals_ssm("s4") (eval_eig.py:282-301): feature `idx` of the vmapped parameters, Lambda_re clipped
    to <= -1e-4 (:288), Q = P (:294), np.linalg.eigvals of A-bar.  Returns (Abar, eigenvalues)."""
    ft = np.float64 if dtype == np.complex128 else np.float32
    step = np.exp(np.asarray(layer["log_step"], ft)[0, idx])
    Lam = np.minimum(np.asarray(layer["Lambda_re"], ft)[:, idx], ft(-1e-4)) + 1j * np.asarray(layer["Lambda_im"], ft)[:, idx]
    P = np.asarray(layer["P"])[:, idx].astype(dtype)
    Ab = discrete_dplr_abar(Lam.astype(dtype), P, P, step, dtype)
    return Ab, np.linalg.eigvals(Ab)


def s4_kernel_dplr(Lambda, P, Q, B, C, step, L):
    """kernel_DPLR (models/s4.py:50-69): the length-L convolution kernel of the DPLR SSM from its truncated generating function at the roots of
    unity (4 Cauchy sums, Woodbury correction, inverse FFT, real part).  complex128."""
    Lambda = np.asarray(Lambda, np.complex128); P = np.asarray(P, np.complex128); Q = np.asarray(Q, np.complex128)
    B = np.asarray(B, np.complex128); C = np.asarray(C, np.complex128)
    Omega = np.exp((-2j * np.pi) * (np.arange(L) / L))
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (2.0 / step) * ((1.0 - Omega) / (1.0 + Omega))
        c = 2.0 / (1.0 + Omega)
        cauchy = lambda v: (v[None, :] / (g[:, None] - Lambda[None, :])).sum(axis=1)
        k00 = cauchy(C.conj() * B); k01 = cauchy(C.conj() * P); k10 = cauchy(Q.conj() * B); k11 = cauchy(Q.conj() * P)
        at_roots = c * (k00 - k01 * (1.0 / (1.0 + k11)) * k10)
    return np.fft.ifft(at_roots, L).real


def s4_discrete_dplr(Lambda, P, Q, B, C, step, L):
    """discrete_DPLR with all three outputs (models/s4.py:16-40): Abar, Bbar and Cbar = conj(C~ (I - Abar^L)^-1 .conj())."""
    Lambda = np.asarray(Lambda, np.complex128); P = np.asarray(P, np.complex128); Q = np.asarray(Q, np.complex128)
    Bc = np.asarray(B, np.complex128)[:, None]; Ct = np.asarray(C, np.complex128)[None, :]
    N = Lambda.shape[0]
    I = np.eye(N)
    A = np.diag(Lambda) - P[:, None] @ Q[:, None].conj().T
    A0 = (2.0 / step) * I + A
    D = np.diag(1.0 / ((2.0 / step) - Lambda))
    Qc = Q.conj().reshape(1, -1); P2 = P.reshape(-1, 1)
    A1 = D - (D @ P2 * (1.0 / (1 + (Qc @ D @ P2))) * Qc @ D)
    Ab = A1 @ A0
    Bb = 2 * A1 @ Bc
    Cb = Ct @ np.linalg.inv(I - np.linalg.matrix_power(Ab, L)).conj()
    return Ab, Bb, Cb.conj()


def s4_kernel_recurrent(Lambda, P, Q, B, C, step, L):
    """K_t = Re(Cbar Abar^t Bbar), t < L: the impulse response of the discretised SSM.  Equals s4_kernel_dplr (the S4 kernel identity) -- the
    independent check that pins the restatement of kernel_DPLR."""
    Ab, Bb, Cb = s4_discrete_dplr(Lambda, P, Q, B, C, step, L)
    x = Bb
    out = np.empty(L)
    for t in range(L):
        out[t] = (Cb @ x).real.item()
        x = Ab @ x
    return out


def s4_forward(layer, u, dtype=np.float64):
    """S4 (vmapped S4Layer) CNN mode (models/s4.py:107-110, :140-146, :169-173): per feature h, y[:, h] = causal_convolution(u[:, h], K_h) + D_h u[:, h],
    K_h = kernel_DPLR(clip(Lambda_re) + i Lambda_im, P, P, B, C~, exp(log_step), L).  layer: vmapped parameter dict (axis 1 = feature); u (B,T,H)."""
    u = np.asarray(u, dtype)
    Bsz, T, H = u.shape
    y = np.empty_like(u)
    for h in range(H):
        Lam = np.minimum(np.asarray(layer["Lambda_re"], np.float64)[:, h], -1e-4) + 1j * np.asarray(layer["Lambda_im"], np.float64)[:, h]
        Pv = np.asarray(layer["P"])[:, h]; Bv = np.asarray(layer["B"])[:, h]
        Cc = np.asarray(layer["C"], np.float64)[:, h, 0] + 1j * np.asarray(layer["C"], np.float64)[:, h, 1]
        step = np.exp(np.asarray(layer["log_step"], np.float64)[0, h])
        K = s4_kernel_dplr(Lam, Pv, Pv, Bv, Cc, step, T)
        n = 2 * T
        conv = np.fft.irfft(np.fft.rfft(u[:, :, h], n, axis=1) * np.fft.rfft(K, n)[None, :], n, axis=1)[:, :T]
        y[:, :, h] = conv + np.asarray(layer["D"], np.float64)[0, h] * u[:, :, h]
    return y


def dplr_exact_spectrum(N, step):
    """At the HiPPO-LegS init diag(Lambda) - P P^* is unitarily similar to the lower-triangular HiPPO matrix, whose
    diagonal is -(k+1): the bilinear map gives the exact real spectrum (2/step - k)/(2/step + k), k = 1..N
    (SURVEY 7-H1)."""
    k = np.arange(1, N + 1, dtype=np.float64)
    return (2.0 / step - k) / (2.0 / step + k)


# --------------------------------------------------------------------------------------------------
# a12: SSD recurrence (third-party mamba-ssm 2.1.0 semantics; PARITY UNPINNED by the reference)
# --------------------------------------------------------------------------------------------------

def ssd_scan_sequential(x, dt, A, Bm, Cm, D=None, dtype=np.float64, return_state=False):
    """What `mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=D, z=None)` computes at its call site
    (models/mamba.py:138-150), restated from the published SSD recurrence:
        h_t[h,p,n] = exp(dt_t[h] A[h]) h_{t-1}[h,p,n] + dt_t[h] B_t[g(h),n] x_t[h,p]
        y_t[h,p]   = sum_n C_t[g(h),n] h_t[h,p,n] + D[h] x_t[h,p]
    x (B,T,H,P), dt (B,T,H) (already softplus'ed), A (H,), Bm/Cm (B,T,G,N), D (H,) -> y (B,T,H,P)."""
    x = np.asarray(x, dtype); dt = np.asarray(dt, dtype); A = np.asarray(A, dtype)
    Bm = np.asarray(Bm, dtype); Cm = np.asarray(Cm, dtype)
    Bsz, T, H, P = x.shape
    G, N = Bm.shape[2], Bm.shape[3]
    rep = H // G
    state = np.zeros((Bsz, H, P, N), dtype)
    y = np.empty_like(x)
    for t in range(T):
        decay = np.exp(dt[:, t] * A)                                  # (B,H)
        Bt = np.repeat(Bm[:, t], rep, axis=1)                         # (B,H,N)
        Ct = np.repeat(Cm[:, t], rep, axis=1)
        state = decay[:, :, None, None] * state + (dt[:, t][:, :, None] * x[:, t])[..., None] * Bt[:, :, None, :]
        y[:, t] = np.einsum("bhpn,bhn->bhp", state, Ct)
    if D is not None:
        y = y + np.asarray(D, dtype)[None, None, :, None] * x
    return (y, state) if return_state else y


def ssd_scan_chunked(x, dt, A, Bm, Cm, D=None, chunk=64):
    """Same quantity in the chunked 'dual' form, torch-CPU fp32, used only as the *timed CPU baseline*
    (a sequential Python loop over T would under-sell the CPU).  Restated from the SSD block decomposition:
    intra-chunk (L o C B^T) X plus inter-chunk state passing."""
    import torch
    x = torch.as_tensor(x); dt = torch.as_tensor(dt); A = torch.as_tensor(A)
    Bm = torch.as_tensor(Bm); Cm = torch.as_tensor(Cm)
    Bsz, T, H, P = x.shape
    G, N = Bm.shape[2], Bm.shape[3]
    assert T % chunk == 0
    nc = T // chunk
    rep = H // G
    a = (dt * A).reshape(Bsz, nc, chunk, H)                           # log decay per step
    cs = torch.cumsum(a, dim=2)                                       # inclusive
    xs = (x * dt[..., None]).reshape(Bsz, nc, chunk, H, P)
    Bc = Bm.repeat_interleave(rep, dim=2).reshape(Bsz, nc, chunk, H, N)
    Cc = Cm.repeat_interleave(rep, dim=2).reshape(Bsz, nc, chunk, H, N)
    # intra-chunk: y_t += sum_{s<=t} exp(cs_t - cs_s) (C_t.B_s) xs_s
    seg = cs[:, :, :, None, :] - cs[:, :, None, :, :]                 # (B,nc,t,s,H)
    mask = torch.tril(torch.ones(chunk, chunk, dtype=torch.bool))
    Lm = torch.where(mask[None, None, :, :, None], torch.exp(torch.where(mask[None, None, :, :, None], seg, torch.zeros((), dtype=seg.dtype))), torch.zeros((), dtype=seg.dtype))
    CB = torch.einsum("bcthn,bcshn->bctsh", Cc, Bc)
    y = torch.einsum("bctsh,bcshp->bcthp", CB * Lm, xs)
    # chunk states: S_c = sum_s exp(cs_last - cs_s) B_s xs_s
    decay_to_end = torch.exp(cs[:, :, -1:, :] - cs)                   # (B,nc,chunk,H)
    S = torch.einsum("bcsh,bcshn,bcshp->bchpn", decay_to_end, Bc, xs)
    chunk_decay = torch.exp(cs[:, :, -1, :])                          # (B,nc,H)
    state = torch.zeros(Bsz, H, P, N, dtype=x.dtype)
    outs = []
    for c in range(nc):
        outs.append(torch.einsum("bthn,bhpn,bth->bthp", Cc[:, c], state, torch.exp(cs[:, c])))
        state = chunk_decay[:, c][:, :, None, None] * state + S[:, c]
    y = y + torch.stack(outs, dim=1)
    y = y.reshape(Bsz, T, H, P)
    if D is not None:
        y = y + torch.as_tensor(D)[None, None, :, None] * x
    return y


# --------------------------------------------------------------------------------------------------
# block glue (torch semantics)
# --------------------------------------------------------------------------------------------------

def layer_norm(x, weight, bias, eps=1e-5):
    """torch.nn.LayerNorm over the last axis, biased variance, eps 1e-5 (models/mamba.py:321; transformer.py:84)."""
    x = np.asarray(x)
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * np.asarray(weight, x.dtype) + np.asarray(bias, x.dtype)


def causal_depthwise_conv_silu(x, weight, bias):
    """nn.Conv1d(groups=C, kernel k, padding k-1) over T, truncated to T, then SiLU (models/mamba.py:98-105,
    :129-133; attention.py:153-156).  x (B,T,C), weight (C,1,k) or (C,k), bias (C,).
    out_t = bias + sum_j w[j] x_{t-(k-1)+j}."""
    x = np.asarray(x)
    w = np.asarray(weight, x.dtype).reshape(x.shape[-1], -1)
    k = w.shape[1]
    T = x.shape[1]
    xp = np.concatenate([np.zeros(x.shape[:1] + (k - 1,) + x.shape[2:], x.dtype), x], axis=1)
    out = np.zeros_like(x) + np.asarray(bias, x.dtype)
    for j in range(k):
        out = out + xp[:, j:j + T, :] * w[:, j]
    return out * sigmoid(out)


def gelu_erf(x):
    """nn.GELU() exact erf form (models/mamba.py:318)."""
    from scipy.special import erf
    x = np.asarray(x)
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def glu(x, weight, bias):
    """GLU (models/common.py:50-58): linear to 2D, first half * sigmoid(second half)."""
    out = x @ np.asarray(weight, x.dtype).T + np.asarray(bias, x.dtype)
    D = x.shape[-1]
    return out[..., :D] * sigmoid(out[..., D:])


def token_embedding(ids, word_emb, pos_emb=None):
    """TokenEmbeddings.forward (models/common.py:160-176)."""
    e = np.asarray(word_emb)[np.asarray(ids)]
    if pos_emb is not None:
        e = e + np.asarray(pos_emb)[: e.shape[1]][None]
    return e


def ssd_mixer_forward(u, p, cfg, dtype=np.float64):
    """SSD.forward (models/mamba.py:111-154).  p: dict of the layer's tensors, cfg: d_inner, ngroups, d_state,
    nheads, headdim."""
    u = np.asarray(u, dtype)
    di, G, N, H, hd = cfg["d_inner"], cfg["ngroups"], cfg["d_state"], cfg["nheads"], cfg["headdim"]
    xbcdt = u @ np.asarray(p["in_proj.weight"], dtype).T
    xBC, dtr = xbcdt[..., : di + 2 * G * N], xbcdt[..., di + 2 * G * N:]
    dt = softplus(dtr + np.asarray(p["dt_bias"], dtype))
    if "conv1d.weight" in p:
        xBC = causal_depthwise_conv_silu(xBC, p["conv1d.weight"], p["conv1d.bias"])
    x, Bm, Cm = xBC[..., :di], xBC[..., di:di + G * N], xBC[..., di + G * N:]
    Bsz, T, _ = u.shape
    y = ssd_scan_sequential(x.reshape(Bsz, T, H, hd), dt, -np.exp(np.asarray(p["A_log"], dtype)),
                            Bm.reshape(Bsz, T, G, N), Cm.reshape(Bsz, T, G, N), np.asarray(p["D"], dtype), dtype)
    return y.reshape(Bsz, T, di) @ np.asarray(p["out_proj.weight"], dtype).T


def ssd_lti_mixer_forward(u, p, cfg, dtype=np.float64):
    """SSD_LTI.forward (models/mamba.py:248-299), ngroups = 1: ONE dt column, dt = softplus(dt_raw + dt_bias) (B,L,H) tiled khead_dim = d_state / H
    times over the state axis, B <- dt * B, then the scan with dt := beta, A := -softplus(A).  (The class cannot be constructed with current torch --
    nn.Parameter(A, device=...) raises -- so this restatement follows the source text; no golden vector can exist.)"""
    u = np.asarray(u, dtype)
    di, G, N, H, hd = cfg["d_inner"], cfg["ngroups"], cfg["d_state"], cfg["nheads"], cfg["headdim"]
    assert G == 1 and N % H == 0
    xbcdt = u @ np.asarray(p["in_proj.weight"], dtype).T
    xBC, dtr = xbcdt[..., : di + 2 * N], xbcdt[..., di + 2 * N:]
    dt = softplus(dtr + np.asarray(p["dt_bias"], dtype))                        # (B,L,1) + (H,) -> (B,L,H)
    if "conv1d.weight" in p:
        xBC = causal_depthwise_conv_silu(xBC, p["conv1d.weight"], p["conv1d.bias"])
    x, Bm, Cm = xBC[..., :di], xBC[..., di:di + N], xBC[..., di + N:]
    Bm = np.repeat(dt, N // H, axis=-1) * Bm                                    # dt.unsqueeze(-1).repeat(.., khead) -> "b l (h d)"
    Bsz, T, _ = u.shape
    beta = np.broadcast_to(np.asarray(p.get("beta", np.ones(H)), dtype), (Bsz, T, H))
    A = -softplus(np.asarray(p["A"], dtype))
    y = ssd_scan_sequential(x.reshape(Bsz, T, H, hd), beta, A, Bm.reshape(Bsz, T, 1, N), Cm.reshape(Bsz, T, 1, N), np.asarray(p["D"], dtype), dtype)
    return y.reshape(Bsz, T, di) @ np.asarray(p["out_proj.weight"], dtype).T


def mamba_block_forward(x, p, cfg, dtype=np.float64):
    """MambaBlock.forward in eval mode / dropout 0 (models/mamba.py:328-340).  p keys are the block's
    state_dict names ('mamba.*', 'glu.linear.*', 'norm.*')."""
    x = np.asarray(x, dtype)
    skip = x
    if cfg.get("prenorm", True):
        x = layer_norm(x, p["norm.weight"], p["norm.bias"])
    mp = {k[len("mamba."):]: v for k, v in p.items() if k.startswith("mamba.")}
    x = ssd_lti_mixer_forward(x, mp, cfg, dtype) if cfg.get("pseudoLTI", False) else ssd_mixer_forward(x, mp, cfg, dtype)
    x = gelu_erf(x)
    if "glu.linear.weight" in p:
        x = glu(x, p["glu.linear.weight"], p["glu.linear.bias"])
    x = x + skip
    if not cfg.get("prenorm", True):
        x = layer_norm(x, p["norm.weight"], p["norm.bias"])
    return x


def linattn_forward(x, p, cfg, dtype=np.float64):
    """MHA(lin_att=True, use_flash=False).forward -> SelfLinAttention (models/attention.py:63-83, :149-182)
    in its O(T) form: out_t = (phi(q_t) . KV_t) / (phi(q_t) . K_t)."""
    x = np.asarray(x, dtype)
    D, dqk, H = cfg["d_model"], cfg["d_qk"], cfg["num_heads"]
    qkv = x @ np.asarray(p["Wqkv.weight"], dtype).T + np.asarray(p["Wqkv.bias"], dtype)
    if "conv1d.weight" in p:
        qkv = causal_depthwise_conv_silu(qkv, p["conv1d.weight"], p["conv1d.bias"])
    Bsz, T, _ = x.shape
    d, dv = dqk // H, D // H
    q = elu(qkv[..., :dqk].reshape(Bsz, T, H, d)) + 1
    k = elu(qkv[..., dqk:2 * dqk].reshape(Bsz, T, H, d)) + 1
    v = qkv[..., 2 * dqk:].reshape(Bsz, T, H, dv)
    kv = np.zeros((Bsz, H, d, dv), dtype)
    ks = np.zeros((Bsz, H, d), dtype)
    out = np.empty((Bsz, T, H, dv), dtype)
    for t in range(T):
        kv = kv + k[:, t][..., None] * v[:, t][:, :, None, :]
        ks = ks + k[:, t]
        num = np.einsum("bhd,bhdt->bht", q[:, t], kv)
        den = np.einsum("bhd,bhd->bh", q[:, t], ks)
        out[:, t] = num / den[..., None]
    return out.reshape(Bsz, T, D) @ np.asarray(p["out_proj.weight"], dtype).T + np.asarray(p["out_proj.bias"], dtype)


def smattn_forward(x, p, cfg, dtype=np.float64):
    """MHA.forward with SelfAttention (models/attention.py:14-35, :149-182), the "naive" path: k scaled by 1/sqrt(d) before the
    product, additive -10000 causal mask, softmax over keys, P V, out_proj.  (use_flash=True rounds q,k,v to fp16 first.)"""
    x = np.asarray(x, dtype)
    D, dqk, H = cfg["d_model"], cfg["d_qk"], cfg["num_heads"]
    qkv = x @ np.asarray(p["Wqkv.weight"], dtype).T + np.asarray(p["Wqkv.bias"], dtype)
    if "conv1d.weight" in p:
        if cfg.get("conv_type", "full") == "full":
            qkv = causal_depthwise_conv_silu(qkv, p["conv1d.weight"], p["conv1d.bias"])
        else:
            qk_ = causal_depthwise_conv_silu(qkv[..., : 2 * dqk], p["conv1d.weight"], p["conv1d.bias"])
            qkv = np.concatenate([qk_, qkv[..., 2 * dqk:]], axis=-1)
    Bsz, T, _ = x.shape
    d, dv = dqk // H, D // H
    q = qkv[..., :dqk].reshape(Bsz, T, H, d)
    k = qkv[..., dqk:2 * dqk].reshape(Bsz, T, H, d)
    v = qkv[..., 2 * dqk:].reshape(Bsz, T, H, dv)
    sc = np.einsum("bthd,bshd->bhts", q, k * dtype(1.0 / math.sqrt(d))) + np.triu(np.full((T, T), -10000.0, dtype), 1)
    sc = sc - sc.max(axis=-1, keepdims=True)
    pr = np.exp(sc)
    pr = pr / pr.sum(axis=-1, keepdims=True)
    out = np.einsum("bhts,bshd->bthd", pr, v)
    return out.reshape(Bsz, T, D) @ np.asarray(p["out_proj.weight"], dtype).T + np.asarray(p["out_proj.bias"], dtype)


def normattn_forward(x, p, cfg, dtype=np.float64):
    """MHNA.forward -> SelfNormAttention (models/norm_attention.py:61-89, :230-258), O(T) form."""
    x = np.asarray(x, dtype)
    D, dqk, H = cfg["d_model"], cfg["d_qk"], cfg["num_heads"]
    vqkn = x @ np.asarray(p["Wvqkn.weight"], dtype).T + np.asarray(p["Wvqkn.bias"], dtype)
    vqk, n = vqkn[..., : D + 2 * dqk], vqkn[..., D + 2 * dqk:]
    if "conv1d.weight" in p:
        if cfg.get("conv_type", "full") == "full":
            vqk = causal_depthwise_conv_silu(vqk, p["conv1d.weight"], p["conv1d.bias"])
        else:
            qk_ = causal_depthwise_conv_silu(vqk[..., D:], p["conv1d.weight"], p["conv1d.bias"])
            vqk = np.concatenate([vqk[..., :D], qk_], axis=-1)
    Bsz, T, _ = x.shape
    d, dv = dqk // H, D // H
    v = vqk[..., :D].reshape(Bsz, T, H, dv)
    q = vqk[..., D:D + dqk].reshape(Bsz, T, H, d)
    k = vqk[..., D + dqk:].reshape(Bsz, T, H, d)
    if cfg.get("approx_fn", "none") == "elu":
        q = elu(q) + 1
        k = elu(k) + 1
    scale = 1.0 / math.sqrt(d) if cfg.get("scale_B", False) else 1.0
    kv = np.zeros((Bsz, H, d, dv), dtype)
    out = np.empty((Bsz, T, H, dv), dtype)
    for t in range(T):
        kv = kv + (k[:, t] * scale)[..., None] * v[:, t][:, :, None, :]
        out[:, t] = np.einsum("bhd,bhdt->bht", q[:, t], kv)
    raw = n + np.asarray(p["inner_attn.offset"], dtype) if "inner_attn.offset" in p else n
    nn_ = np.exp(-norm_fn_apply(cfg["norm_fn"], raw))
    out = nn_[..., None] * out
    return out.reshape(Bsz, T, D) @ np.asarray(p["out_proj.weight"], dtype).T + np.asarray(p["out_proj.bias"], dtype)


def transformer_block_forward(x, p, cfg, dtype=np.float64):
    """TransformerBlock.forward, eval mode (models/transformer.py:90-111): ONE LayerNorm reused for both
    sub-blocks (:94, :99); mixer 'none' drops the second skip (:73-75, :102-104)."""
    x = np.asarray(x, dtype)
    z = None
    if "Wz.weight" in p:
        z = x @ np.asarray(p["Wz.weight"], dtype).T + np.asarray(p["Wz.bias"], dtype)
    skip = x
    xn = layer_norm(x, p["norm.weight"], p["norm.bias"])
    ap = {k[len("attention."):]: v for k, v in p.items() if k.startswith("attention.")}
    if cfg["attention_fn"] == "lin-attention":
        a = linattn_forward(xn, ap, cfg, dtype)
    elif cfg["attention_fn"] == "norm-attention":
        a = normattn_forward(xn, ap, cfg, dtype)
    elif cfg["attention_fn"] == "sm-attention":
        a = smattn_forward(xn, ap, cfg, dtype)
    else:
        raise NotImplementedError(cfg["attention_fn"])
    x = a + skip
    y = layer_norm(x, p["norm.weight"], p["norm.bias"])
    mixer = cfg.get("mixer", "none")
    silu = lambda t: t * sigmoid(t)
    if mixer == "none":
        return y * silu(z) if z is not None else y
    if mixer == "glu":
        y = glu(y, p["mixer.linear.weight"], p["mixer.linear.bias"])
    elif mixer == "mlp":
        y = gelu_erf(y @ np.asarray(p["mixer.encoder.weight"], dtype).T + np.asarray(p["mixer.encoder.bias"], dtype))
        y = y @ np.asarray(p["mixer.decoder.weight"], dtype).T + np.asarray(p["mixer.decoder.bias"], dtype)
    elif mixer == "hybrid":                                        # LAMBDA, models/common.py:60-84
        xz = y @ np.asarray(p["mixer.encoder.weight"], dtype).T + np.asarray(p["mixer.encoder.bias"], dtype)
        a = sigmoid(np.asarray(p["mixer.alpha"], dtype).reshape(()))
        Dm = xz.shape[-1] // 2
        g = xz[..., :Dm] * sigmoid(xz[..., Dm:])
        m = gelu_erf(xz) @ np.asarray(p["mixer.decoder.weight"], dtype).T + np.asarray(p["mixer.decoder.bias"], dtype)
        y = a * g + (1 - a) * m
    else:
        raise NotImplementedError(mixer)
    return (x + y) * silu(z) if z is not None else x + y


# --------------------------------------------------------------------------------------------------
# the per-layer analysis loop (eval_eig.py:510-526, :584-600, :627-663)
# --------------------------------------------------------------------------------------------------

def _block_params(state_dict, prefix):
    return {k[len(prefix):]: np.asarray(v) for k, v in state_dict.items() if k.startswith(prefix)}


def mamba_eval_pass(ids_or_x, state_dict, cfg, dtype=np.float64):
    """One pass of the Mamba branch: encoder, then for each layer i: x = block_i(x); eig_i = get_eig_mamba2(x, block_i)
    -- the extractor sees the block OUTPUT and applies block i's own dt rows to it (eval_eig.py:512-517, :586-591).
    Returns eig (B,T,H,L) and the final activations."""
    if "encoder.word_embeddings.weight" in state_dict:
        x = token_embedding(ids_or_x, state_dict["encoder.word_embeddings.weight"]).astype(dtype)
    else:
        x = np.asarray(ids_or_x, dtype) @ np.asarray(state_dict["encoder.weight"], dtype).T + np.asarray(state_dict["encoder.bias"], dtype)
    eigs = []
    for i in range(cfg["num_layers"]):
        p = _block_params(state_dict, "blocks.%d." % i)
        x = mamba_block_forward(x, p, cfg, dtype)
        if cfg.get("pseudoLTI", False):
            eigs.append(mamba2_lti_eig(x.shape[0], x.shape[1], p["mamba.A"], p.get("mamba.beta", np.ones(cfg["nheads"])), dtype))
        else:
            eigs.append(mamba2_eig(x, p["mamba.in_proj.weight"], p["mamba.dt_bias"], p["mamba.A_log"],
                                   cfg["d_inner"], cfg["ngroups"], cfg["d_state"], cfg["nheads"], dtype))
    return np.concatenate(eigs, axis=-1), x


def transformer_eval_pass(ids_or_x, state_dict, cfg, dtype=np.float64, eta_dtype=np.float32):
    """One pass of the Transformer branch (eval_eig.py:528-564, :627-663)."""
    if "encoder.word_embeddings.weight" in state_dict:
        x = token_embedding(ids_or_x, state_dict["encoder.word_embeddings.weight"],
                            state_dict.get("encoder.position_embeddings.weight")).astype(dtype)
    else:
        x = np.asarray(ids_or_x, dtype) @ np.asarray(state_dict["encoder.weight"], dtype).T + np.asarray(state_dict["encoder.bias"], dtype)
    etas = []
    for i in range(cfg["num_layers"]):
        p = _block_params(state_dict, "layers.%d." % i)
        x = transformer_block_forward(x, p, cfg, dtype)
        if cfg["attention_fn"] == "lin-attention":
            q, k = linattn_qk(x, p["attention.Wqkv.weight"], p["attention.Wqkv.bias"], cfg["d_qk"], cfg["num_heads"], eta_dtype)
            etas.append(linattn_eta_prefix(q, k, eta_dtype))
        elif cfg["attention_fn"] == "sm-attention":
            q, k = linattn_qk(x, p["attention.Wqkv.weight"], p["attention.Wqkv.bias"], cfg["d_qk"], cfg["num_heads"], eta_dtype)
            etas.append(softmax_eta_closed(q, k, eta_dtype))
        else:
            etas.append(normattn_eta(x, p["attention.Wvqkn.weight"], p["attention.Wvqkn.bias"],
                                     p.get("attention.inner_attn.offset") if cfg.get("offset", False) else None,
                                     cfg["norm_fn"], cfg["d_model"], cfg["d_qk"], cfg["num_heads"], eta_dtype))
    return np.concatenate(etas, axis=-1), x


# --------------------------------------------------------------------------------------------------
# torch-CPU fp32 port of one Mamba analysis pass -- the TIMED CPU BASELINE (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------------------

def mamba_eval_pass_torch_cpu(ids, state_dict, cfg, chunk=64):
    """The reference's per-layer loop for the Mamba branch (eval_eig.py:575-618) the way it runs on a host without CUDA:
    torch fp32 on all cores for the block forward (F.layer_norm, F.linear, F.conv1d, chunked SSD in place of the
    CUDA-only mamba_ssm Triton kernel), get_eig_mamba2's projection/softplus/exp, `.numpy()` per layer, np.concatenate,
    then the radius and the threshold statistics in NumPy.  Returns (eig (B,T,H,L) f32, percentage (7,B,H,L), phase pct)."""
    import torch
    import torch.nn.functional as F
    sd = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))) for k, v in state_dict.items()}
    di, G, N, H, hd = cfg["d_inner"], cfg["ngroups"], cfg["d_state"], cfg["nheads"], cfg["headdim"]
    with torch.no_grad():
        ids = torch.as_tensor(ids)
        x = F.embedding(ids, sd["encoder.word_embeddings.weight"])
        Bsz, T, D = x.shape
        eig = None
        for i in range(cfg["num_layers"]):
            p = "blocks.%d." % i
            skip = x
            xn = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"]) if cfg.get("prenorm", True) else x
            z = F.linear(xn, sd[p + "mamba.in_proj.weight"])
            xBC, dt = z[..., : di + 2 * G * N], z[..., di + 2 * G * N:]
            dt = F.softplus(dt + sd[p + "mamba.dt_bias"])
            if p + "mamba.conv1d.weight" in sd:
                w = sd[p + "mamba.conv1d.weight"]
                k = w.shape[-1]
                xBC = F.silu(F.conv1d(xBC.transpose(1, 2), w, sd[p + "mamba.conv1d.bias"], padding=k - 1, groups=w.shape[0]).transpose(1, 2))[:, :T]
            xs, Bm, Cm = xBC[..., :di], xBC[..., di:di + G * N], xBC[..., di + G * N:]
            y = ssd_scan_chunked(xs.reshape(Bsz, T, H, hd), dt, -torch.exp(sd[p + "mamba.A_log"]), Bm.reshape(Bsz, T, G, N),
                                 Cm.reshape(Bsz, T, G, N), sd[p + "mamba.D"], chunk=chunk).reshape(Bsz, T, di)
            o = F.gelu(F.linear(y, sd[p + "mamba.out_proj.weight"]))
            if p + "glu.linear.weight" in sd:
                g = F.linear(o, sd[p + "glu.linear.weight"], sd[p + "glu.linear.bias"])
                o = g[..., :D] * torch.sigmoid(g[..., D:])
            x = o + skip
            if not cfg.get("prenorm", True):
                x = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"])
            # get_eig_mamba2 (eval_eig.py:176-190): the whole in_proj again on the block output
            zz = F.linear(x, sd[p + "mamba.in_proj.weight"])
            lam = torch.exp(F.softplus(zz[..., di + 2 * G * N:] + sd[p + "mamba.dt_bias"]) * -torch.exp(sd[p + "mamba.A_log"]))
            lam = np.expand_dims(lam.numpy(), axis=-1)
            eig = lam if eig is None else np.concatenate((eig, lam), axis=-1)
    rad = np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2))
    pct = threshold_analysis(rad, THRESHOLDS_RADIUS)
    ph = threshold_analysis(np.arctan2(eig.imag, eig.real) * 180 / np.pi, THRESHOLDS_PHASE)
    return eig, pct, ph
