"""Thin torch-tensor wrappers over the C ABI (include/eigb200.h).  torch is plumbing here: it owns device memory and
the current stream; every computation below is a hand-written sm_100a kernel inside libeigb200.so."""
from __future__ import annotations

import ctypes as C
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


def new_counts(B: int, inner: int, device) -> torch.Tensor:
    return torch.zeros(B, inner, NSLOT, dtype=torch.int32, device=device)


def mamba2_eig(x, W_dt, dt_bias, A_log, thresholds: Sequence[float] = THRESHOLDS_RADIUS, want_lam=True,
               counts: Optional[torch.Tensor] = None, want_counts=True, compare="float64", lam_out=None):
    """K1.  x (B,T,D) f32|bf16 -> (lam (B,T,H) f32 | None, counts (B,H,8) int32 | None)."""
    x = _prep(x, name="x")
    B, T, D = x.shape
    W_dt = _prep(W_dt, torch.float32); dt_bias = _prep(dt_bias, torch.float32); A_log = _prep(A_log, torch.float32)
    H = W_dt.shape[0]
    lib = _enter(x)
    lam = None
    if want_lam:
        lam = lam_out if lam_out is not None else torch.empty(B, T, H, dtype=torch.float32, device=x.device)
        assert lam.is_contiguous() and lam.numel() == B * T * H and lam.dtype == torch.float32
    if want_counts and counts is None:
        counts = new_counts(B, H, x.device)
    thr, n = L.thresholds_arg(thresholds)
    L.check(lib.eigb200_mamba2_eig(_stream(x), _p(x), _xdtype(x), B, T, D, _p(W_dt), _p(dt_bias), _p(A_log), H,
                                   _p(lam), _p(counts if want_counts else None), thr, n, _cmp(compare)), "eigb200_mamba2_eig")
    return lam, (counts if want_counts else None)


def mamba2_lti_eig(A, beta, B, T, thresholds=THRESHOLDS_RADIUS, want_lam=True, counts=None, compare="float64"):
    A = _prep(A, torch.float32); beta = _prep(beta, torch.float32)
    H = A.shape[0]
    lib = _enter(A)
    lam = torch.empty(B, T, H, dtype=torch.float32, device=A.device) if want_lam else None
    if counts is None:
        counts = new_counts(B, H, A.device)
    thr, n = L.thresholds_arg(thresholds)
    L.check(lib.eigb200_mamba2_lti_eig(_stream(A), _p(A), _p(beta), B, T, H, _p(lam), _p(counts), thr, n, _cmp(compare)),
            "eigb200_mamba2_lti_eig")
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
    L.check(lib.eigb200_normattn_gate(_stream(x), _p(x), _xdtype(x), B, T, D, _p(W_n), _p(b_n), _p(offset), H,
                                      L.NORM_FN[norm_fn], _p(n)), "eigb200_normattn_gate")
    return n


def ratio_hist(a, mode: int, thresholds=THRESHOLDS_RADIUS, want_out=True, counts=None, compare="float64"):
    """a (B,N,inner) f32|f64.  mode RATIO_NONE: plain threshold counts; ratio modes: eta (B,N-1,inner) f64 + counts."""
    a = _prep(a, name="a")
    if a.dtype not in (torch.float32, torch.float64):
        raise L.Eigb200Error("ratio_hist: float32 or float64 input required")
    B, N = a.shape[0], a.shape[1]
    inner = int(np.prod(a.shape[2:])) if a.dim() > 2 else 1
    lib = _enter(a)
    out = None
    if mode != L.RATIO_NONE and want_out:
        out = torch.empty((B, N - 1) + tuple(a.shape[2:]), dtype=torch.float64, device=a.device)
    if counts is None:
        counts = new_counts(B, inner, a.device)
    thr, n = L.thresholds_arg(thresholds)
    L.check(lib.eigb200_ratio_hist(_stream(a), _p(a), L.F32 if a.dtype == torch.float32 else L.F64, mode, B, N, inner,
                                   _p(out), _p(counts), thr, n, _cmp(compare)), "eigb200_ratio_hist")
    return out, counts


def count_moments(counts):
    counts = _prep(counts, torch.int32)
    B = counts.shape[0]
    inner = counts.numel() // (B * NSLOT)
    lib = _enter(counts)
    s = torch.empty(inner, NSLOT, dtype=torch.int64, device=counts.device)
    s2 = torch.empty_like(s)
    L.check(lib.eigb200_count_moments(_stream(counts), _p(counts), B, inner, _p(s), _p(s2)), "eigb200_count_moments")
    return s, s2
