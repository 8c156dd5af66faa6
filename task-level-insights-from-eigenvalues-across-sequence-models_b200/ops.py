"""Thin torch-tensor wrappers over the C ABI (include/eigb200.h).  torch is plumbing here: it owns device memory and
the current stream; every computation below is a hand-written sm_100a kernel inside libeigb200.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L

THRESHOLDS_RADIUS = (0.1, 0.5, 0.9, 1.0, 10.0, 100.0)      # analysis/eval_eig.py:603, :665, :724
THRESHOLDS_PHASE = (1.0, 10.0, 45.0, 90.0, 180.0)           # analysis/eval_eig.py:612, :671, :732
NSLOT = L.NSLOT


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _prep(t: torch.Tensor, dtype=None, name="tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise L.Eigb200Error("%s must be a CUDA tensor: the eigb200 path has no CPU fallback" % name)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _enter(t: torch.Tensor):
    lib = L.load()
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    L.check(lib.eigb200_set_device(idx), "eigb200_set_device")
    return lib


def _cmp(compare: str) -> int:
    return {"float64": L.CMP_F64, "float32": L.CMP_F32}[compare]


def _xdtype(x):
    if x.dtype == torch.float32:
        return L.F32
    if x.dtype == torch.bfloat16:
        return L.BF16
    raise L.Eigb200Error("activations must be float32 or bfloat16, got %s" % x.dtype)


LAUNCHES = {"n": 0}          # C-ABI compute calls enqueued through this module (bench.py reports it as gpu_launches)
PROFILE = None               # set to a list to record (name, start_event, end_event) per call (bench.py roofline leg)


def _call(lib, name, stream, *args, tag=None):
    """One C-ABI call: enqueue on `stream`, raise on a non-zero status, count it, optionally bracket it with CUDA events
    (recorded under `name` or `name[tag]`, e.g. the GEMM shape, so that bench.py can attribute time to ONE kernel configuration)."""
    fn = getattr(lib, name)
    if PROFILE is not None and tag:
        name_rec = "%s[%s]" % (name, tag)
    else:
        name_rec = name
    if PROFILE is not None:
        ext = torch.cuda.ExternalStream(stream.value) if stream.value else torch.cuda.default_stream()
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record(ext)
        rc = fn(stream, *args)
        s1.record(ext)
        PROFILE.append((name_rec, s0, s1))
    else:
        rc = fn(stream, *args)
    L.check(rc, name)
    LAUNCHES["n"] += 1


def _uniform_stride(t: torch.Tensor, shape) -> int:
    """Element stride s such that entry with row-major index i over `shape` lives at t.data_ptr() + i*s; raises otherwise."""
    assert tuple(t.shape) == tuple(shape), (t.shape, shape)
    s = t.stride(-1)
    expect = s
    for dim in range(t.dim() - 1, -1, -1):
        if t.shape[dim] != 1 and t.stride(dim) != expect:
            raise L.Eigb200Error("output view is not uniformly strided: shape %s strides %s" % (tuple(t.shape), t.stride()))
        expect *= t.shape[dim]
    return s


def new_counts(B: int, inner: int, device) -> torch.Tensor:
    return torch.zeros(B, inner, NSLOT, dtype=torch.int32, device=device)


def mamba2_eig(x, W_dt, dt_bias, A_log, thresholds: Sequence[float] = THRESHOLDS_RADIUS, want_lam=True,
               counts: Optional[torch.Tensor] = None, want_counts=True, compare="float64", lam_out=None,
               rowstats_out: Optional[torch.Tensor] = None, ln_eps: float = 1e-5):
    """K1.  x (B,T,D) f32|bf16 -> (lam (B,T,H) f32 | None, counts (B,H,8) int32 | None).
    rowstats_out (B,T,2) f32: also emit the LayerNorm (mean, rstd) of every row of x for the next block (see linear_ln)."""
    x = _prep(x, name="x")
    B, T, D = x.shape
    W_dt = _prep(W_dt, torch.float32); dt_bias = _prep(dt_bias, torch.float32); A_log = _prep(A_log, torch.float32)
    H = W_dt.shape[0]
    lib = _enter(x)
    lam = None
    stride = 1
    if want_lam:
        lam = lam_out if lam_out is not None else torch.empty(B, T, H, dtype=torch.float32, device=x.device)
        assert lam.dtype == torch.float32
        stride = _uniform_stride(lam, (B, T, H))
    if want_counts and counts is None:
        counts = new_counts(B, H, x.device)
    thr, n = L.thresholds_arg(thresholds)
    _call(lib, "eigb200_mamba2_eig", _stream(x), _p(x), _xdtype(x), B, T, D, _p(W_dt), _p(dt_bias), _p(A_log), H,
                                   _p(lam), stride, _p(counts if want_counts else None), thr, n, _cmp(compare),
          _p(rowstats_out), float(ln_eps))
    return lam, (counts if want_counts else None)


def mamba2_lti_eig(A, beta, B, T, thresholds=THRESHOLDS_RADIUS, want_lam=True, counts=None, compare="float64", lam_out=None):
    A = _prep(A, torch.float32); beta = _prep(beta, torch.float32)
    H = A.shape[0]
    lib = _enter(A)
    lam, stride = None, 1
    if want_lam:
        lam = lam_out if lam_out is not None else torch.empty(B, T, H, dtype=torch.float32, device=A.device)
        stride = _uniform_stride(lam, (B, T, H))
    if counts is None:
        counts = new_counts(B, H, A.device)
    thr, n = L.thresholds_arg(thresholds)
    _call(lib, "eigb200_mamba2_lti_eig", _stream(A), _p(A), _p(beta), B, T, H, _p(lam), stride, _p(counts), thr, n, _cmp(compare))
    return lam, counts


def normattn_gate(x, W_n, b_n, offset, norm_fn: str):
    """K1' first half: n (B,T,H) f32."""
    if norm_fn not in L.NORM_FN:
        raise RuntimeError("normalization function {0} not implemented!".format(norm_fn))     # eval_eig.py:151
    x = _prep(x, name="x")
    B, T, D = x.shape
    W_n = _prep(W_n, torch.float32); b_n = _prep(b_n, torch.float32)
    offset = _prep(offset, torch.float32) if offset is not None else None
    H = W_n.shape[0]
    lib = _enter(x)
    n = torch.empty(B, T, H, dtype=torch.float32, device=x.device)
    _call(lib, "eigb200_normattn_gate", _stream(x), _p(x), _xdtype(x), B, T, D, _p(W_n), _p(b_n), _p(offset), H,
                                      L.NORM_FN[norm_fn], _p(n))
    return n


def ratio_hist(a, mode: int, thresholds=THRESHOLDS_RADIUS, want_out=True, counts=None, compare="float64", out=None):
    """a (B,N,inner) f32|f64.  mode RATIO_NONE: plain threshold counts; ratio modes: eta (B,N-1,inner) f64 + counts."""
    a = _prep(a, name="a")
    if a.dtype not in (torch.float32, torch.float64):
        raise L.Eigb200Error("ratio_hist: float32 or float64 input required")
    B, N = a.shape[0], a.shape[1]
    inner = int(np.prod(a.shape[2:])) if a.dim() > 2 else 1
    lib = _enter(a)
    ostride = 1
    if mode != L.RATIO_NONE and want_out:
        if out is None:
            out = torch.empty((B, N - 1) + tuple(a.shape[2:]), dtype=torch.float64, device=a.device)
        assert out.dtype == torch.float64 and out.numel() == B * (N - 1) * inner
        ostride = _uniform_stride(out, tuple(out.shape))
    else:
        out = None
    if counts is None:
        counts = new_counts(B, inner, a.device)
    thr, n = L.thresholds_arg(thresholds)
    _call(lib, "eigb200_ratio_hist", _stream(a), _p(a), L.F32 if a.dtype == torch.float32 else L.F64, mode, B, N, inner,
                                   _p(out), ostride, _p(counts), thr, n, _cmp(compare))
    return out, counts


def log_hist(a, lo=1e-8, hi=1e2, nbins=512, hist=None):
    """Log-spaced histogram of a (B,N,inner...) f32|f64 array per inner column -> (inner, nbins + 3) int64, accumulated into `hist` when given."""
    a = _prep(a, name="a")
    if a.dtype not in (torch.float32, torch.float64):
        raise L.Eigb200Error("log_hist: float32 or float64 input required")
    B, N = a.shape[0], a.shape[1]
    inner = int(np.prod(a.shape[2:])) if a.dim() > 2 else 1
    lib = _enter(a)
    if hist is None:
        hist = torch.zeros(inner, nbins + 3, dtype=torch.int64, device=a.device)
    assert hist.is_contiguous() and hist.dtype == torch.int64 and tuple(hist.shape) == (inner, nbins + 3)
    _call(lib, "eigb200_log_hist", _stream(a), _p(a), L.F32 if a.dtype == torch.float32 else L.F64, B, N, inner, float(lo), float(hi), int(nbins), _p(hist))
    return hist


def hist_quantiles(hist, qs, lo=1e-8, hi=1e2):
    """Quantiles (inner, len(qs)) float64 from a log_hist histogram, on the device."""
    hist = _prep(hist, torch.int64)
    inner, nslot = hist.shape
    q = torch.as_tensor(np.asarray(qs, np.float64), device=hist.device)
    lib = _enter(hist)
    out = torch.empty(inner, q.numel(), dtype=torch.float64, device=hist.device)
    _call(lib, "eigb200_hist_quantiles", _stream(hist), _p(hist), inner, float(lo), float(hi), nslot - 3, _p(q), q.numel(), _p(out))
    return out


def count_moments(counts):
    counts = _prep(counts, torch.int32)
    B = counts.shape[0]
    inner = counts.numel() // (B * NSLOT)
    lib = _enter(counts)
    s = torch.empty(inner, NSLOT, dtype=torch.int64, device=counts.device)
    s2 = torch.empty_like(s)
    _call(lib, "eigb200_count_moments", _stream(counts), _p(counts), B, inner, _p(s), _p(s2))
    return s, s2


def count_moments_layers(counts, out=None):
    """counts (L,B,inner...,8) int32 of a whole pass -> moments (2,L,inner,8) int64: [0] = sum_b c, [1] = sum_b c^2, one launch.
    This buffer is all a rank contributes to the path's single all-reduce (dist.allreduce_moments)."""
    counts = _prep(counts, torch.int32)
    Ln, B = counts.shape[0], counts.shape[1]
    inner = counts.numel() // (Ln * B * NSLOT)
    lib = _enter(counts)
    if out is None:
        out = torch.empty(2, Ln, inner, NSLOT, dtype=torch.int64, device=counts.device)
    assert out.is_contiguous() and out.dtype == torch.int64 and out.numel() == 2 * Ln * inner * NSLOT
    _call(lib, "eigb200_count_moments_layers", _stream(counts), _p(counts), Ln, B, inner, _p(out[0]), _p(out[1]))
    return out


# ---- K1'': linear attention ---------------------------------------------------------------------------------------
def linattn_nu(qk_buf, ld: int, B: int, T: int, H: int, d: int, k_offset: int):
    """qk_buf: projection buffer (B*T, ld) f32 with q at column 0 and k at column `k_offset`.  -> nu (B,T,H) f64."""
    qk_buf = _prep(qk_buf, torch.float32)
    lib = _enter(qk_buf)
    nu = torch.empty(B, T, H, dtype=torch.float64, device=qk_buf.device)
    q_ptr = C.c_void_p(qk_buf.data_ptr())
    k_ptr = C.c_void_p(qk_buf.data_ptr() + 4 * k_offset)
    _call(lib, "eigb200_linattn_nu", _stream(qk_buf), q_ptr, k_ptr, ld, B, T, H, d, _p(nu))
    return nu


def softmax_nu(qk_buf, ld: int, B: int, T: int, H: int, d: int, k_offset: int):
    """Softmax-attention normaliser with the reference's masking quirk.  -> (nu (B,T,H) f64, m (B,T,H) f32)."""
    qk_buf = _prep(qk_buf, torch.float32)
    lib = _enter(qk_buf)
    nu = torch.empty(B, T, H, dtype=torch.float64, device=qk_buf.device)
    m = torch.empty(B, T, H, dtype=torch.float32, device=qk_buf.device)
    q_ptr = C.c_void_p(qk_buf.data_ptr())
    k_ptr = C.c_void_p(qk_buf.data_ptr() + 4 * k_offset)
    _call(lib, "eigb200_softmax_nu", _stream(qk_buf), q_ptr, k_ptr, ld, B, T, H, d, _p(nu), _p(m))
    return nu, m


def softmax_eta(nu, m, thresholds=THRESHOLDS_RADIUS, want_out=True, counts=None):
    """eta_t = nu_t / nu_{t+1} * exp(m_t - m_{t+1}) (B,T-1,H) f64 and its threshold counts (B,H,8)."""
    nu = _prep(nu, torch.float64); m = _prep(m, torch.float32)
    B, T, H = nu.shape
    lib = _enter(nu)
    out = torch.empty(B, T - 1, H, dtype=torch.float64, device=nu.device) if want_out else None
    if counts is None:
        counts = new_counts(B, H, nu.device)
    thr, n = L.thresholds_arg(thresholds)
    _call(lib, "eigb200_softmax_eta", _stream(nu), _p(nu), _p(m), B, T, H, _p(out), _p(counts), thr, n)
    return out, counts


def softmax_attn_forward(buf, ld, q_off, k_off, v_off, B, T, H, d, dv, scale):
    """Causal softmax attention over a projection buffer (B*T, ld); q/k/v at the given column offsets.  -> (B,T,H*dv)."""
    buf = _prep(buf, torch.float32)
    lib = _enter(buf)
    out = torch.empty(B, T, H * dv, dtype=torch.float32, device=buf.device)
    base = buf.data_ptr()
    _call(lib, "eigb200_softmax_attn_forward", _stream(buf), C.c_void_p(base + 4 * q_off), C.c_void_p(base + 4 * k_off), C.c_void_p(base + 4 * v_off),
          ld, float(scale), _p(out), H * dv, B, T, H, d, dv)
    return out


def linattn_forward(buf, ld, q_off, k_off, v_off, B, T, H, d, dv, gate=None, phi_elu=True, normalise=True, kscale=1.0):
    """Causal linear attention over a projection buffer (B*T, ld); q/k/v live at the given column offsets."""
    buf = _prep(buf, torch.float32)
    lib = _enter(buf)
    out = torch.empty(B, T, H * dv, dtype=torch.float32, device=buf.device)
    base = buf.data_ptr()
    gate = _prep(gate, torch.float32) if gate is not None else None
    _call(lib, "eigb200_linattn_forward", _stream(buf), C.c_void_p(base + 4 * q_off), C.c_void_p(base + 4 * k_off),
                                        C.c_void_p(base + 4 * v_off), ld, _p(gate), int(phi_elu), int(normalise), float(kscale),
                                        _p(out), H * dv, B, T, H, d, dv)
    return out


def linattn_conv_fusable(buf, ld, q_off, k_off, v_off, H, d, dv, kconv) -> bool:
    """True when linattn_forward_conv takes this shape (chunked tensor-core kernel: d = dv = 64, taps <= 4, aligned rows)."""
    lib = _enter(buf)
    base = buf.data_ptr()
    fn = lib.eigb200_linattn_conv_fusable
    return bool(fn(C.c_void_p(base + 4 * q_off), C.c_void_p(base + 4 * k_off), C.c_void_p(base + 4 * v_off), ld, C.c_void_p(256), H * dv, d, dv, kconv))


def linattn_forward_conv(buf, ld, q_off, k_off, v_off, B, T, H, d, dv, conv_w, conv_b, conv_ch_q, conv_ch_k, conv_ch_v,
                         gate=None, phi_elu=True, normalise=True, kscale=1.0):
    """conv1d + SiLU (depthwise, causal) of the q / k / v columns fused into the causal linear attention: buf holds the RAW projection.
    conv_w (C, k), conv_b (C); conv_ch_* = conv channel of each matrix's first column, -1 = not convolved."""
    buf = _prep(buf, torch.float32)
    conv_w = _prep(conv_w, torch.float32); conv_b = _prep(conv_b, torch.float32)
    lib = _enter(buf)
    out = torch.empty(B, T, H * dv, dtype=torch.float32, device=buf.device)
    base = buf.data_ptr()
    gate = _prep(gate, torch.float32) if gate is not None else None
    _call(lib, "eigb200_linattn_forward_conv", _stream(buf), C.c_void_p(base + 4 * q_off), C.c_void_p(base + 4 * k_off),
          C.c_void_p(base + 4 * v_off), ld, _p(gate), int(phi_elu), int(normalise), float(kscale),
          _p(conv_w), _p(conv_b), int(conv_w.shape[1]), int(conv_ch_q), int(conv_ch_k), int(conv_ch_v),
          _p(out), H * dv, B, T, H, d, dv)
    return out


# ---- K2 scans -----------------------------------------------------------------------------------------------------
def diag_scan(lam, Bu, reverse=False):
    """lam (P,) complex64, Bu (B,T,P) complex64 -> h (B,T,P) complex64  (h_t = lam h_{t-1} + Bu_t)."""
    lam = _prep(lam, torch.complex64); Bu = _prep(Bu, torch.complex64)
    B, T, P = Bu.shape
    lib = _enter(Bu)
    h = torch.empty_like(Bu)
    _call(lib, "eigb200_diag_scan", _stream(Bu), _p(torch.view_as_real(lam)), _p(torch.view_as_real(Bu)), _p(torch.view_as_real(h)),
                                  B, T, P, int(reverse))
    return h


def ssd_scan(x, dt, A, Bm, Cm, D=None, return_final_state=False):
    """mamba_chunk_scan_combined semantics.  x (B,T,H,P), dt (B,T,H), A (H), Bm/Cm (B,T,G,N), D (H) -> y (B,T,H,P)."""
    x = _prep(x, torch.float32); dt = _prep(dt, torch.float32); A = _prep(A, torch.float32)
    Bm = _prep(Bm, torch.float32); Cm = _prep(Cm, torch.float32)
    D = _prep(D, torch.float32) if D is not None else None
    B, T, H, P = x.shape
    G, N = Bm.shape[2], Bm.shape[3]
    lib = _enter(x)
    y = torch.empty_like(x)
    fs = torch.empty(B, H, P, N, dtype=torch.float32, device=x.device) if return_final_state else None
    _call(lib, "eigb200_ssd_scan", _stream(x), _p(x), H * P, _p(dt), _p(A), _p(Bm), _p(Cm), G * N, _p(D), _p(y), H * P, _p(fs),
                                 B, T, H, P, G, N)
    return (y, fs) if return_final_state else y


def ssd_scan_buffer(buf, ld, col_x, col_b, col_c, dt, A, D, B, T, H, P, G, N):
    """eigb200_ssd_scan over operands that live in one projection buffer (B*T, ld): x at column col_x (H*P wide), B / C at col_b / col_c (G*N wide).
    dt (B,T,H) dense, A (H), D (H) | None -> y (B,T,H*P)."""
    buf = _prep(buf, torch.float32); dt = _prep(dt, torch.float32); A = _prep(A, torch.float32)
    D = _prep(D, torch.float32) if D is not None else None
    lib = _enter(buf)
    y = torch.empty(B, T, H * P, dtype=torch.float32, device=buf.device)
    base = buf.data_ptr()
    _call(lib, "eigb200_ssd_scan", _stream(buf), C.c_void_p(base + 4 * col_x), ld, _p(dt), _p(A), C.c_void_p(base + 4 * col_b), C.c_void_p(base + 4 * col_c), ld,
          _p(D), _p(y), H * P, None, B, T, H, P, G, N)
    return y


def mamba_conv_ssd(xbcdt, ldz, conv_w, conv_b, dt_bias, A_log, D, B, T, H, P, G, N, out=None):
    """Fused conv+SiLU / softplus(dt) / SSD scan over the raw in_proj output (B*T, ldz) -> y (B,T,H*P)."""
    xbcdt = _prep(xbcdt, torch.float32)
    lib = _enter(xbcdt)
    k = 0
    if conv_w is not None:
        conv_w = _prep(conv_w, torch.float32).reshape(conv_w.shape[0], -1)
        conv_b = _prep(conv_b, torch.float32)
        k = conv_w.shape[1]
    y = out if out is not None else torch.empty(B, T, H * P, dtype=torch.float32, device=xbcdt.device)
    _call(lib, "eigb200_mamba_conv_ssd", _stream(xbcdt), _p(xbcdt), ldz, _p(conv_w), _p(conv_b), k, _p(_prep(dt_bias, torch.float32)),
                                       _p(_prep(A_log, torch.float32)), _p(_prep(D, torch.float32)), _p(y), H * P, B, T, H, P, G, N)
    return y


# ---- K4 + glue ----------------------------------------------------------------------------------------------------
EPILOGUES = {"none": L.EPI_NONE, "gelu": L.EPI_GELU, "glu_residual": L.EPI_GLU_RESIDUAL, "residual": L.EPI_RESIDUAL}
GEMM_MODES = {"auto": L.GEMM_AUTO, "simt": L.GEMM_SIMT_F32, "tc3": L.GEMM_TC_3XTF32, "tc1": L.GEMM_TC_TF32, "f16x3": L.GEMM_TC_F16X3}


def gemm_precision() -> str:
    """Operand precision the tensor-core GEMMs use by default (EIGB200_GEMM_PRECISION): "tf32x3" or "f16x3"."""
    return "f16x3" if int(L.load().eigb200_gemm_precision()) == 1 else "tf32x3"


def set_gemm_precision(name) -> None:
    """ "tf32x3" | "f16x3" | None (back to EIGB200_GEMM_PRECISION).  Prepared workspaces of the other precision become stale: MambaDev.invalidate_prepared()."""
    L.load().eigb200_set_gemm_precision({"tf32x3": 0, "f16x3": 1, None: -1}[name])


def gemm_overflow(device=None, reset=True) -> bool:
    """True when an fp16-split GEMM met an activation beyond its range since the last reset (its output holds inf / NaN there): rerun with 3xTF32.
    Synchronises the current stream of `device`."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lib = L.load()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    L.check(lib.eigb200_set_device(idx), "eigb200_set_device")
    flag = C.c_int(0)
    L.check(lib.eigb200_gemm_overflow(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), int(reset), C.byref(flag)), "eigb200_gemm_overflow")
    return flag.value != 0
_ws_cache = {}


def linear_workspace(N, K, device, M=0):
    """Workspace of the tensor-core GEMM for (M, N, K): split weights (+ the split copy of A when K > 256)."""
    lib = L.load()
    nbytes = int(lib.eigb200_linear_workspace_bytes_m(int(M), N, K))
    if nbytes == 0:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def linear_prepare(weight, bias=None, epilogue="none", ln_gamma=None, ln_beta=None):
    """Split `weight` once into the tensor-core operand layout (weights are constants of an analysis run): -> workspace tensor to pass as
    `prepared=` to linear / linear_ln / linear_glu_extract with the same weight, epilogue and LayerNorm, or None when the shape has no
    resident-weight plan (those keep preparing per call)."""
    weight = _prep(weight, torch.float32)
    assert weight.is_cuda
    N, K = weight.shape
    lib = _enter(weight)
    nbytes = int(lib.eigb200_linear_workspace_bytes(N, K))
    if nbytes == 0:
        return None
    ws = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    _call(lib, "eigb200_linear_prepare", _stream(weight), _p(weight), _p(_prep(bias, torch.float32) if bias is not None else None),
          _p(_prep(ln_gamma, torch.float32) if ln_gamma is not None else None), _p(_prep(ln_beta, torch.float32) if ln_beta is not None else None),
          N, K, EPILOGUES[epilogue], _p(ws), nbytes, tag="N%d K%d %s" % (N, K, epilogue))
    return ws


# narrow GEMMs with K > 256 as K slabs on the resident-weight kernel (see linear): opt-in since the streamed-operand kernel converts A on the SM
# (no split pre-pass): 0.229 ms against 0.305 ms for two slabs at the C h projection of C3 (262 144 x 128 x 512)
K_SLABS = os.environ.get("EIGB200_K_SLABS", "0") == "1"


def linear(a, weight, bias=None, epilogue="none", residual=None, mode="auto", out=None, ldc=None, workspace=None, prepared=None):
    """a (..., K) with a.stride(-2) as the row stride -> epilogue(a W^T + bias).  weight (N,K) torch layout.
    prepared: workspace from linear_prepare(weight, bias, epilogue) -- no per-call weight preparation kernels (tensor-core path only)."""
    assert a.is_cuda and a.dtype == torch.float32 and a.stride(-1) == 1
    K = a.shape[-1]
    M = a.numel() // K
    if a.is_contiguous():
        a2 = a.reshape(M, K); lda = K
    elif a.dim() == 2:
        a2 = a; lda = a.stride(0)                                   # strided 2-D view (column slice of a projection buffer): row stride passed through
    else:
        a2 = a.contiguous().reshape(M, K); lda = K                  # an N-D non-contiguous view has no single row stride: compact it first
    weight = _prep(weight, torch.float32)
    N = weight.shape[0]
    assert weight.shape[1] == K
    bias = _prep(bias, torch.float32) if bias is not None else None
    if (K > 256 and N <= 128 and epilogue in ("none", "residual") and mode in ("auto", "f16x3", "tc3") and M >= 1024 and prepared is None
            and K_SLABS and a2.data_ptr() % 16 == 0 and lda % 4 == 0):
        # K in slabs of <= 256 columns through the RESIDENT-weight kernel, each slab accumulating onto the previous result through the residual epilogue:
        # that kernel reads the raw fp32 rows by TMA (row stride lda) and splits them on the SM.  It beat the streamed-operand kernel while that one wrote and
        # re-read a split copy of A (0.48 -> 0.27 ms at the C h projection of C3); the streamed kernel now converts on the SM too and is faster (0.23 ms).
        nsl = (K + 255) // 256
        step = ((K + nsl - 1) // nsl + 3) // 4 * 4
        acc = residual if epilogue == "residual" else None
        for i, k0 in enumerate(range(0, K, step)):
            k1 = min(K, k0 + step)
            last = k1 == K
            acc = linear(a2[:, k0:k1], weight[:, k0:k1].contiguous(), bias if i == 0 else None, epilogue="residual" if acc is not None else "none",
                         residual=acc, mode=mode, out=out if last else None, ldc=ldc if last else None)
        return acc
    nout = N // 2 if epilogue == "glu_residual" else N
    lib = _enter(a)
    if out is None:
        ldc = ldc or nout
        out = torch.empty(M, ldc, dtype=torch.float32, device=a.device)
    else:
        ldc = out.stride(-2)
    ldr = 0
    if residual is not None:
        assert residual.is_cuda and residual.dtype == torch.float32 and residual.stride(-1) == 1
        ldr = residual.stride(-2) if residual.dim() >= 2 else nout
    if prepared is not None:
        ws, wsb = prepared, prepared.numel()
    else:
        ws, wsb = (workspace, workspace.numel()) if workspace is not None else linear_workspace(N, K, a.device, M)
    _call(lib, "eigb200_linear", _stream(a), _p(a2), lda, _p(weight if prepared is None else None), _p(bias), _p(out), ldc, _p(residual), ldr, M, N, K,
                               EPILOGUES[epilogue], GEMM_MODES[mode], _p(ws), wsb, tag="N%d K%d %s" % (N, K, epilogue))
    return out


def linear_glu_extract(a, weight, bias, residual, w_gate, partials=None, out=None, workspace=None, prepared=None):
    """GLU + residual GEMM that also leaves the extractor partials of its output rows: -> (out (M, N/2), partials (N/32, 3, M))."""
    assert a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()
    K = a.shape[-1]
    M = a.numel() // K
    weight = _prep(weight, torch.float32); bias = _prep(bias, torch.float32) if bias is not None else None
    N = weight.shape[0]
    nout = N // 2
    lib = _enter(a)
    if out is None:
        out = torch.empty(M, nout, dtype=torch.float32, device=a.device)
    if partials is None:
        partials = torch.empty(nout // 16, 3, M, dtype=torch.float32, device=a.device)
    if prepared is not None:
        ws, wsb = prepared, prepared.numel()
    else:
        ws, wsb = (workspace, workspace.numel()) if workspace is not None else linear_workspace(N, K, a.device, M)
    _call(lib, "eigb200_linear_glu_extract", _stream(a), _p(a), K, _p(weight if prepared is None else None), _p(bias), _p(out), out.stride(-2), _p(residual), residual.stride(-2),
          M, N, K, _p(_prep(w_gate, torch.float32)), _p(partials), _p(ws), wsb, tag="N%d K%d glu_residual+extract" % (N, K))
    return out, partials


def out_glu_fused_supported(D, K1) -> bool:
    return gemm_precision() == "f16x3" and int(L.load().eigb200_out_glu_fused_supported(int(D), int(K1))) == 1


def out_glu_fused(y, prepared_out, bias_out, prepared_glu, bias_glu, residual, w_gate=None, partials=None, out=None):
    """GLU(GELU(y W_out^T + b_out) W_glu^T + b_glu) + residual in one kernel (fp16-split precision, prepared operands): -> (out (M, D), partials | None)."""
    assert y.is_cuda and y.dtype == torch.float32 and y.is_contiguous() and residual.is_cuda and residual.dtype == torch.float32
    K1 = y.shape[-1]
    M = y.numel() // K1
    D = residual.shape[-1]
    lib = _enter(y)
    if out is None:
        out = torch.empty(M, D, dtype=torch.float32, device=y.device)
    if w_gate is not None and partials is None:
        partials = torch.empty(D // 16, 3, M, dtype=torch.float32, device=y.device)
    _call(lib, "eigb200_out_glu_fused", _stream(y), _p(y), K1, _p(prepared_out), _p(_prep(bias_out, torch.float32) if bias_out is not None else None),
          _p(prepared_glu), _p(_prep(bias_glu, torch.float32) if bias_glu is not None else None), _p(out), out.stride(-2), _p(residual), residual.stride(-2),
          M, D, K1, _p(_prep(w_gate, torch.float32) if w_gate is not None else None), _p(partials if w_gate is not None else None),
          tag="D%d K%d out_proj+gelu -> glu_residual%s" % (D, K1, "+extract" if w_gate is not None else ""))
    return out, (partials if w_gate is not None else None)


def mamba_front_fused_supported(D, d_inner, H, G, N, kconv) -> bool:
    return gemm_precision() == "f16x3" and int(L.load().eigb200_mamba_front_fused_supported(int(D), int(d_inner), int(H), int(G), int(N), int(kconv))) == 1


def mamba_front_fused(x, stats, prepared_in, conv_w, conv_b, dt_bias, A_log, D, d_inner, N, out=None):
    """LayerNorm -> in_proj -> conv + SiLU -> softplus(dt) -> SSD scan in one kernel (models/mamba.py:329-331, :118-150): x (B,T,D) -> y (B,T,d_inner).
    stats (B,T,2) = (mean, rstd) of the rows of x; prepared_in = linear_prepare(W_in, None, "none", gamma, beta) under the fp16-split precision."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and stats.is_cuda and stats.dtype == torch.float32 and stats.is_contiguous()
    B, T, Dm = x.shape
    lib = _enter(x)
    conv_w = _prep(conv_w, torch.float32).reshape(conv_w.shape[0], -1)
    conv_b = _prep(conv_b, torch.float32)
    y = out if out is not None else torch.empty(B, T, d_inner, dtype=torch.float32, device=x.device)
    _call(lib, "eigb200_mamba_front_fused", _stream(x), _p(x), Dm, _p(stats), _p(prepared_in), _p(conv_w), _p(conv_b), conv_w.shape[1],
          _p(_prep(dt_bias, torch.float32)), _p(_prep(A_log, torch.float32)), _p(_prep(D, torch.float32) if D is not None else None),
          _p(y), d_inner, B, T, Dm, d_inner, N, tag="D%d P%d N%d ln+in_proj+conv+ssd" % (Dm, d_inner, N))
    return y


def mamba2_eig_partials(partials, B, T, dt_bias, A_log, thresholds: Sequence[float] = THRESHOLDS_RADIUS, want_lam=True, counts=None,
                        compare="float64", lam_out=None, rowstats_out=None, ln_eps: float = 1e-5):
    """Finish the Mamba-2 extractor (H = 1) from the GLU epilogue's partials.  -> (lam (B,T,1) | None, counts (B,1,8))."""
    partials = _prep(partials, torch.float32)
    ng = partials.shape[0]
    lib = _enter(partials)
    lam, stride = None, 1
    if lam_out is not None:
        lam = lam_out; stride = _uniform_stride(lam_out, (B, T, 1))
    elif want_lam:
        lam = torch.empty(B, T, 1, dtype=torch.float32, device=partials.device)
    if counts is None:
        counts = new_counts(B, 1, partials.device)
    thr, n = L.thresholds_arg(thresholds)
    _call(lib, "eigb200_mamba2_eig_partials", _stream(partials), _p(partials), ng, B, T, _p(_prep(dt_bias, torch.float32)), _p(_prep(A_log, torch.float32)),
          _p(lam), stride, _p(counts), thr, n, _cmp(compare), _p(rowstats_out), float(ln_eps))
    return lam, counts


def rowstats(x, eps=1e-5):
    """(mean, rstd) of every row of x (..., D) -> (..., 2) float32."""
    x = _prep(x, torch.float32)
    D = x.shape[-1]
    lib = _enter(x)
    out = torch.empty(x.shape[:-1] + (2,), dtype=torch.float32, device=x.device)
    _call(lib, "eigb200_rowstats", _stream(x), _p(x), x.numel() // D, D, float(eps), _p(out))
    return out


def linear_ln(a, stats, gamma, beta, weight, bias=None, epilogue="none", residual=None, out=None, ldc=None, workspace=None, prepared=None):
    """epilogue(LayerNorm(a) W^T + bias) with the normalisation applied inside the GEMM's A-operand stage (tensor-core path)."""
    assert a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()
    K = a.shape[-1]
    M = a.numel() // K
    weight = _prep(weight, torch.float32)
    N = weight.shape[0]
    bias = _prep(bias, torch.float32) if bias is not None else None
    nout = N // 2 if epilogue == "glu_residual" else N
    lib = _enter(a)
    if out is None:
        ldc = ldc or nout
        out = torch.empty(M, ldc, dtype=torch.float32, device=a.device)
    else:
        ldc = out.stride(-2)
    ldr = residual.stride(-2) if residual is not None else 0
    if prepared is not None:
        ws, wsb = prepared, prepared.numel()
    else:
        ws, wsb = (workspace, workspace.numel()) if workspace is not None else linear_workspace(N, K, a.device, M)
    _call(lib, "eigb200_linear_ln", _stream(a), _p(a), K, _p(_prep(stats, torch.float32)), _p(_prep(gamma, torch.float32)),
          _p(_prep(beta, torch.float32)), _p(weight if prepared is None else None), _p(bias), _p(out), ldc, _p(residual), ldr, M, N, K, EPILOGUES[epilogue], _p(ws), wsb,
          tag="N%d K%d %s" % (N, K, epilogue))
    return out


def linear_ln_supported(N, K):
    return int(L.load().eigb200_linear_workspace_bytes_m(1024, N, K)) > 0


def embedding(ids, word, pos=None, rowstats_out=None, ln_eps=1e-5):
    ids = _prep(ids, torch.int64); word = _prep(word, torch.float32)
    pos = _prep(pos, torch.float32) if pos is not None else None
    B, T = ids.shape
    V, D = word.shape
    if pos is not None and pos.shape[0] < T:
        raise L.Eigb200Error("embedding: sequence length %d exceeds max_position_embeddings %d" % (T, pos.shape[0]))
    lib = _enter(ids)
    out = torch.empty(B, T, D, dtype=torch.float32, device=ids.device)
    if rowstats_out is not None:
        _call(lib, "eigb200_embedding_stats", _stream(ids), _p(ids), _p(word), _p(pos), _p(out), B, T, D, V, _p(rowstats_out), float(ln_eps))
    else:
        _call(lib, "eigb200_embedding", _stream(ids), _p(ids), _p(word), _p(pos), _p(out), B, T, D, V)
    return out


def layernorm(x, w, b, eps=1e-5):
    x = _prep(x, torch.float32)
    D = x.shape[-1]
    lib = _enter(x)
    out = torch.empty_like(x)
    _call(lib, "eigb200_layernorm", _stream(x), _p(x), _p(_prep(w, torch.float32)), _p(_prep(b, torch.float32)), float(eps), _p(out),
                                  x.numel() // D, D)
    return out


def conv_silu(x, ldx, w, b, B, T, Cn, out=None, ldo=None):
    """x: buffer (B*T, ldx) whose first Cn columns are convolved; -> (B*T, ldo)."""
    lib = _enter(x)
    w = _prep(w, torch.float32).reshape(w.shape[0], -1); b = _prep(b, torch.float32)
    if out is None:
        ldo = ldo or Cn
        out = torch.empty(B * T, ldo, dtype=torch.float32, device=x.device)
    _call(lib, "eigb200_conv_silu", _stream(x), _p(x), ldx, _p(w), _p(b), w.shape[1], _p(out), ldo, B, T, Cn)
    return out


def add(a, b):
    a = _prep(a, torch.float32); b = _prep(b, torch.float32)
    lib = _enter(a)
    out = torch.empty_like(a)
    _call(lib, "eigb200_add", _stream(a), _p(a), _p(b), _p(out), a.numel())
    return out


def mul_silu(y, z):
    y = _prep(y, torch.float32); z = _prep(z, torch.float32)
    lib = _enter(y)
    out = torch.empty_like(y)
    _call(lib, "eigb200_mul_silu", _stream(y), _p(y), _p(z), _p(out), y.numel())
    return out


def gelu(x):
    x = _prep(x, torch.float32)
    lib = _enter(x)
    out = torch.empty_like(x)
    _call(lib, "eigb200_gelu", _stream(x), _p(x), _p(out), x.numel())
    return out


def scale_cols(a, s):
    a = _prep(a, torch.float32); s = _prep(s, torch.float32)
    lib = _enter(a)
    out = torch.empty_like(a)
    cols = a.shape[-1]
    _call(lib, "eigb200_scale_cols", _stream(a), _p(a), _p(s), _p(out), a.numel() // cols, cols)
    return out


def lti_scale_b(buf, ld, col_b, col_dt, dt_bias, N, khead):
    """SSD_LTI: B <- softplus(dt_raw + dt_bias[head of column]) * B, in place on the projection buffer (rows, ld)."""
    assert buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous()
    lib = _enter(buf)
    _call(lib, "eigb200_lti_scale_b", _stream(buf), _p(buf), ld, col_b, col_dt, _p(_prep(dt_bias, torch.float32)), buf.numel() // ld, N, khead)
    return buf


def s4_kernel(Lambda, P, Q, Bv, Cv, step, L_):
    """kernel_DPLR for H features: (H,N) complex64 x5, step (H) f32 -> Kt (L,H) float32 (lag-major)."""
    Lambda = _prep(Lambda, torch.complex64); P = _prep(P, torch.complex64); Q = _prep(Q, torch.complex64)
    Bv = _prep(Bv, torch.complex64); Cv = _prep(Cv, torch.complex64); step = _prep(step, torch.float32).reshape(-1)
    H, N = Lambda.shape
    lib = _enter(Lambda)
    Kt = torch.empty(L_, H, dtype=torch.float32, device=Lambda.device)
    nb = int(lib.eigb200_s4_kernel_workspace_bytes(H, L_))
    ws = torch.empty(nb, dtype=torch.uint8, device=Lambda.device)
    r = torch.view_as_real
    _call(lib, "eigb200_s4_kernel", _stream(Lambda), _p(r(Lambda)), _p(r(P)), _p(r(Q)), _p(r(Bv)), _p(r(Cv)), _p(step), H, N, L_, _p(Kt), _p(ws), nb)
    return Kt


def s4_causal_conv(u, Kt, D=None):
    """y[b,t,h] = sum_{s<=t} Kt[t-s,h] u[b,s,h] + D[h] u[b,t,h]."""
    u = _prep(u, torch.float32); Kt = _prep(Kt, torch.float32)
    B, T, H = u.shape
    if Kt.shape != (T, H):
        raise L.Eigb200Error("s4_causal_conv: kernel shape %s does not match (T, H) = (%d, %d)" % (tuple(Kt.shape), T, H))
    lib = _enter(u)
    y = torch.empty_like(u)
    _call(lib, "eigb200_s4_causal_conv", _stream(u), _p(u), _p(Kt), _p(_prep(D, torch.float32)) if D is not None else None, _p(y), B, T, H)
    return y


SSM_KINDS = {"lru": 0, "s5_zoh": 1, "s5_bilinear": 2}


def ssm_lambda(kind: str, p0, p1, p2=None):
    """Parameter-only eigenvalues (P,) complex64 of LRU / S5 on the device."""
    p0 = _prep(p0, torch.float32).reshape(-1); p1 = _prep(p1, torch.float32).reshape(-1)
    p2 = _prep(p2, torch.float32).reshape(-1) if p2 is not None else None
    lib = _enter(p0)
    lam = torch.empty(p0.numel(), dtype=torch.complex64, device=p0.device)
    _call(lib, "eigb200_ssm_lambda", _stream(p0), SSM_KINDS[kind], _p(p0), _p(p1), _p(p2), p0.numel(), _p(torch.view_as_real(lam)))
    return lam


def dplr_abar(Lambda, Pv, Qv, step):
    """(nmat,N) complex64 x3, step (nmat,) f32 -> Abar (nmat,N,N) complex64."""
    Lambda = _prep(Lambda, torch.complex64); Pv = _prep(Pv, torch.complex64); Qv = _prep(Qv, torch.complex64)
    step = _prep(step, torch.float32).reshape(-1)
    nmat, N = Lambda.shape
    lib = _enter(Lambda)
    Ab = torch.empty(nmat, N, N, dtype=torch.complex64, device=Lambda.device)
    _call(lib, "eigb200_dplr_abar", _stream(Lambda), _p(torch.view_as_real(Lambda)), _p(torch.view_as_real(Pv)), _p(torch.view_as_real(Qv)),
                                  _p(step), nmat, N, _p(torch.view_as_real(Ab)))
    return Ab


def eigvals_c64(A):
    """np.linalg.eigvals for a batch (nmat,N,N) of complex64 matrices, N <= 64.  Returns (eig (nmat,N) complex64, info (nmat,) int32)."""
    A = _prep(A, torch.complex64).clone()
    nmat, N, _ = A.shape
    lib = _enter(A)
    ev = torch.empty(nmat, N, dtype=torch.complex64, device=A.device)
    info = torch.empty(nmat, dtype=torch.int32, device=A.device)
    _call(lib, "eigb200_eigvals_c64", _stream(A), _p(torch.view_as_real(A)), nmat, N, _p(torch.view_as_real(ev)), _p(info))
    return ev, info
