"""The `get_eig_*` extractors and `threshold_analysis*` with the reference's signatures (analysis/eval_eig.py:43-391),
executed by the eigb200 kernels.

Each extractor exists twice:
  * `get_eig_xxx(x, layer, ...)`        -- drop-in: same arguments, returns a HOST numpy array with the trailing singleton axis
                                           exactly like the reference (one device->host copy, as in the reference);
  * `get_eig_xxx_device(x, layer, ...)` -- returns (eig device tensor | None, counts (B,H,8) int32 device tensor): what
                                           `eval_eig` uses so that eigenvalues never leave HBM unless asked for.
`layer` is duck-typed like in the reference: `.mamba.{in_proj,A_log,dt_bias,d_inner,ngroups,d_state,nheads}` or
`.attention.{Wqkv|Wvqkn, head_dim, inner_attn.offset}` -- eigb200.layers objects or the reference's own nn.Modules.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops

THRESHOLDS_RADIUS = np.array([0.1, 0.5, 0.9, 1.0, 10, 100])
THRESHOLDS_PHASE = np.array([1, 10, 45, 90, 180])


def _w(t):
    return t.detach() if isinstance(t, torch.Tensor) else torch.as_tensor(t)


def _cuda(x):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise L.Eigb200Error("eigb200 extractors need a CUDA device; there is no CPU fallback")
        x = x.cuda()
    return x


# ---- Mamba-2 ---------------------------------------------------------------------------------------------------------
def _mamba_gate_rows(m):
    w = getattr(m, "W_dt", None)
    if w is not None:
        return w
    lo = m.d_inner + 2 * m.ngroups * m.d_state                     # eval_eig.py:179-181
    return _w(m.in_proj.weight)[lo:lo + m.nheads].contiguous()


def get_eig_mamba2_device(x, layer, want_eig=True, counts=None, compare="float64", lam_out=None, rowstats_out=None):
    m = layer.mamba
    x = _cuda(x)
    return ops.mamba2_eig(x, _cuda(_mamba_gate_rows(m)), _cuda(_w(m.dt_bias)), _cuda(_w(m.A_log)),
                          want_lam=want_eig, counts=counts, compare=compare, lam_out=lam_out, rowstats_out=rowstats_out)


def get_eig_mamba2(x, layer):
    """analysis/eval_eig.py:176-190.  (B,T,D) -> (B,T,H,1) float32 numpy."""
    lam, _ = get_eig_mamba2_device(x, layer)
    return np.expand_dims(lam.cpu().numpy(), axis=-1)


def get_eig_mamba2_LTI_device(x, layer, want_eig=True, counts=None, compare="float64", lam_out=None):
    m = layer.mamba
    x = _cuda(x)
    B, T, _ = x.shape
    return ops.mamba2_lti_eig(_cuda(_w(m.A)), _cuda(_w(m.beta)), B, T, want_lam=want_eig, counts=counts, compare=compare, lam_out=lam_out)


def get_eig_mamba2_LTI(x, layer):
    """analysis/eval_eig.py:192-205."""
    lam, _ = get_eig_mamba2_LTI_device(x, layer)
    return np.expand_dims(lam.cpu().numpy(), axis=-1)


# ---- normalised attention -----------------------------------------------------------------------------------------------
def get_eig_att_norm_device(x, layer, d_qk, num_heads, d_model, model_config, want_eig=True, counts=None, compare="float64"):
    norm_fn_cf = model_config["norm_fn"]
    if norm_fn_cf not in L.NORM_FN:
        raise RuntimeError("normalization function {0} not implemented!".format(norm_fn_cf))     # eval_eig.py:151
    att = layer.attention
    x = _cuda(x)
    W_n = getattr(att, "W_n", None)
    if W_n is None:
        lo = d_model + 2 * d_qk                                                                   # eval_eig.py:156-158
        W_n = _w(att.Wvqkn.weight)[lo:lo + num_heads].contiguous()
        b_n = _w(att.Wvqkn.bias)[lo:lo + num_heads].contiguous()
    else:
        b_n = att.b_n
    offset = _w(att.inner_attn.offset) if model_config["offset"] else None                        # eval_eig.py:160-163
    n = ops.normattn_gate(x, _cuda(W_n), _cuda(b_n), _cuda(offset) if offset is not None else None, norm_fn_cf)
    return ops.ratio_hist(n, L.RATIO_NEXT_OVER_CUR, want_out=want_eig, counts=counts, compare=compare)


def get_eig_att_norm(x, layer, d_qk, num_heads, d_model, model_config):
    """analysis/eval_eig.py:137-174.  -> (B,T-1,H,1) float64 numpy."""
    eta, _ = get_eig_att_norm_device(x, layer, d_qk, num_heads, d_model, model_config)
    return np.expand_dims(eta.cpu().numpy(), axis=-1)


# ---- linear attention -----------------------------------------------------------------------------------------------------
def get_eig_att_linear_device(x, layer, d_qk, num_heads, d_model, want_eig=True, counts=None, compare="float64"):
    att = layer.attention
    x = _cuda(x)
    B, T, _ = x.shape
    W_qk = getattr(att, "W_qk", None)
    if W_qk is None:
        W_qk = _w(att.Wqkv.weight)[: 2 * d_qk].contiguous()                                       # eval_eig.py:99-103
        b_qk = _w(att.Wqkv.bias)[: 2 * d_qk].contiguous() if att.Wqkv.bias is not None else None
    else:
        b_qk = att.b_qk
    qk = ops.linear(x, _cuda(W_qk), _cuda(b_qk) if b_qk is not None else None)                    # (B*T, 2*d_qk): [q (h d) | k (h d)]
    nu = ops.linattn_nu(qk, 2 * d_qk, B, T, num_heads, att.head_dim, d_qk)
    return ops.ratio_hist(nu, L.RATIO_CUR_OVER_NEXT, want_out=want_eig, counts=counts, compare=compare)


def get_eig_att_linear(x, layer, d_qk, num_heads, d_model):
    """analysis/eval_eig.py:97-135 in O(T) memory.  -> (B,T-1,H,1) float64 numpy."""
    eta, _ = get_eig_att_linear_device(x, layer, d_qk, num_heads, d_model)
    return np.expand_dims(eta.cpu().numpy(), axis=-1)


def get_eig_att_softmax_device(x, layer, d_qk, num_heads, d_model, want_eig=True, counts=None, compare="float64"):
    att = layer.attention
    x = _cuda(x)
    B, T, _ = x.shape
    W_qk = getattr(att, "W_qk", None)
    if W_qk is None:
        W_qk = _w(att.Wqkv.weight)[: 2 * d_qk].contiguous()                                       # eval_eig.py:46-48
        b_qk = _w(att.Wqkv.bias)[: 2 * d_qk].contiguous() if att.Wqkv.bias is not None else None
    else:
        b_qk = att.b_qk
    qk = ops.linear(x, _cuda(W_qk), _cuda(b_qk) if b_qk is not None else None)                    # (B*T, 2*d_qk): [q (h d) | k (h d)]
    nu, m = ops.softmax_nu(qk, 2 * d_qk, B, T, num_heads, att.head_dim, d_qk)
    return ops.softmax_eta(nu, m, want_out=want_eig, counts=counts)


def get_eig_att_softmax(x, layer, d_qk, num_heads, d_model):
    """analysis/eval_eig.py:43-95 in O(T) memory (no (B,T,T,H) tensors), reproducing the multiplicative-mask row maximum.  -> (B,T-1,H,1) float64."""
    eta, _ = get_eig_att_softmax_device(x, layer, d_qk, num_heads, d_model)
    return np.expand_dims(eta.cpu().numpy(), axis=-1)


def get_eig_from_qkv_att_softmax_device(q, k, v=None, want_eig=True, counts=None):
    """notebooks/lm_eigvals.ipynb, cell 13 (`get_eig_from_qkv_att_softmax`): the softmax-attention eigenvalues from hooked q_proj / k_proj outputs of a
    pretrained LM (cells 11-16: OLMo-3).  q (B,T,Hq,d), k (B,T,Hkv,d) in any float dtype (the notebook hooks fp16 activations); v is accepted and unused,
    as in the notebook.  Grouped-query attention: with Hkv < Hq every query head h reads key head h // (Hq / Hkv), the `repeat_kv` convention of the
    HF models the notebook loads (the notebook's own einsum needs Hkv == Hq; OLMo-3-7B has 32 / 32).  Same formula and masking quirk as
    get_eig_att_softmax (eval_eig.py:43-95), O(T) memory.  -> (eta (B,T-1,Hq) float64 device tensor, counts (B,Hq,8))."""
    q = _cuda(q).float(); k = _cuda(k).float()
    B, T, Hq, d = q.shape
    Hkv = k.shape[2]
    if k.shape[0] != B or k.shape[1] != T or k.shape[3] != d or Hq % Hkv != 0:
        raise L.Eigb200Error("get_eig_from_qkv_att_softmax: q %s and k %s do not describe (grouped-query) attention heads" % (tuple(q.shape), tuple(k.shape)))
    if Hkv != Hq:
        k = k.repeat_interleave(Hq // Hkv, dim=2)
    qk = torch.cat([q.reshape(B * T, Hq * d), k.reshape(B * T, Hq * d)], dim=1).contiguous()          # [q (h d) | k (h d)], the layout softmax_nu reads
    nu, m = ops.softmax_nu(qk, 2 * Hq * d, B, T, Hq, d, Hq * d)
    return ops.softmax_eta(nu, m, want_out=want_eig, counts=counts)


def get_eig_from_qkv_att_softmax(q, k, v=None):
    """Drop-in for the notebook function: -> eta (B,T-1,Hq,1) float64 numpy."""
    eta, _ = get_eig_from_qkv_att_softmax_device(q, k, v)
    return np.expand_dims(eta.cpu().numpy(), axis=-1)


# ---- threshold statistics ---------------------------------------------------------------------------------------------------
def threshold_counts_device(eig_val, thresholds, compare="float64"):
    """eig_val (B,N,...) device tensor -> counts (B, inner, 8) int32 (slots: see include/eigb200.h)."""
    _, counts = ops.ratio_hist(eig_val, L.RATIO_NONE, thresholds=[float(t) for t in np.asarray(thresholds).ravel()], compare=compare)
    return counts


def threshold_analysis(eig_val, thresholds, num_layers, num_heads, batch_size, compare="float64"):
    """analysis/eval_eig.py:335-362.  eig_val (B,N,H,L) numpy or torch -> percentages (n_thr+1, B, H, L) float64 numpy."""
    thresholds = np.asarray(thresholds).flatten()
    nb = thresholds.shape[0] + 1
    t = _cuda(eig_val if isinstance(eig_val, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(eig_val)))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    B, N = t.shape[0], t.shape[1]
    counts = threshold_counts_device(t.reshape(B, N, -1), thresholds, compare)              # (B, H*L, 8)
    c = counts[..., :nb].cpu().numpy().astype(np.float64)
    pct = np.moveaxis(c, -1, 0).reshape((nb, batch_size, num_heads, num_layers)) / N * 100
    return pct


def threshold_analysis_ssm(eig_val, thresholds, num_layers, compare="float64"):
    """analysis/eval_eig.py:364-391.  eig_val (P,L) -> (n_thr+1, L)."""
    thresholds = np.asarray(thresholds).flatten()
    nb = thresholds.shape[0] + 1
    t = _cuda(eig_val if isinstance(eig_val, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(eig_val)))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    P = t.shape[0]
    counts = threshold_counts_device(t.reshape(1, P, -1), thresholds, compare)              # (1, L, 8)
    c = counts[0, :, :nb].cpu().numpy().astype(np.float64)
    return c.T.reshape(nb, num_layers) / P * 100


def percentages_from_counts(counts, n_per_seq, nb):
    """counts (L, B, H, 8) int -> percentages (nb, B, H, L) float64, count / N * 100 (eval_eig.py:351)."""
    c = counts[..., :nb].astype(np.float64)
    return np.transpose(c, (3, 1, 2, 0)) / n_per_seq * 100


def phase_percentages_from_counts(counts, n_per_seq, nb_phase):
    """First phase bin from slot 7, the other bins are empty for real non-negative / zeroed values (eval_eig.py:612-618, :673-674)."""
    L_, B, H, _ = counts.shape
    out = np.zeros((nb_phase, B, H, L_), np.float64)
    out[0] = np.transpose(counts[..., 7].astype(np.float64), (1, 2, 0)) / n_per_seq * 100
    return out
