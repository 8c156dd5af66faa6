"""Device-side mirrors of the reference's torch layers, driven by a reference-format state_dict.

Only what the analysis path needs: parameter containers with the reference's attribute names (so the `get_eig_*`
extractors can be handed either these objects or the reference's own modules) and a `__call__` that propagates the
activations through one block with the eigb200 kernels.

Reference: models/mamba.py (SSD :25-154, MambaBlock :301-340, Mamba :342-389), models/transformer.py
(TransformerBlock :22-111, Transformer :113-161), models/attention.py (MHA :85-182), models/norm_attention.py
(MHNA :160-258), models/common.py (GLU :50-58, MLP :33-48, TokenEmbeddings :117-176).
Eval-mode semantics (dropout is the identity).  The reference's INIT pass runs in train mode (eval_eig.py:484-564 never calls model.eval()), so with
dropout > 0 its eig_init is one random sample; eval_eig warns in that case.
"""
from __future__ import annotations

import os

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _pad8(n: int) -> int:
    """Row pitch of GEMM outputs: a multiple of 8 floats keeps every row 32-byte aligned for the 256-bit epilogue stores."""
    return (n + 7) // 8 * 8


class _Lin:
    """nn.Linear-shaped parameter holder: .weight (N,K), .bias (N,) | None; callable through the eigb200 GEMM."""

    def __init__(self, weight, bias=None):
        self.weight = weight
        self.bias = bias

    def __call__(self, x, **kw):
        out = ops.linear(x, self.weight, self.bias, **kw)
        return out.reshape(x.shape[:-1] + (out.shape[-1],)) if out.shape[-1] == self.weight.shape[0] else out


def _dev(sd: Dict[str, torch.Tensor], key: str, device, dtype=torch.float32) -> Optional[torch.Tensor]:
    if key not in sd:
        return None
    t = sd[key]
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device=device, dtype=dtype).contiguous()


# ======================================================================================================================
# Mamba-2
# ======================================================================================================================

class SSDParams(SimpleNamespace):
    """Attribute names follow models/mamba.py:SSD so that get_eig_mamba2(x, layer) reads them unchanged."""


class MambaBlockDev:
    def __init__(self, sd, prefix, cfg, device):
        D = cfg["hidden_dim"]
        headdim = D // cfg["num_heads"]
        d_inner = cfg["expansion"] * D
        N = cfg["state_dim"]
        self.prenorm = cfg["prenorm"]
        m = SSDParams()
        m.d_model, m.d_inner, m.d_state, m.ngroups, m.headdim = D, d_inner, N, 1, headdim
        m.nheads = d_inner // headdim
        m.in_proj = _Lin(_dev(sd, prefix + "mamba.in_proj.weight", device))
        m.out_proj = _Lin(_dev(sd, prefix + "mamba.out_proj.weight", device))
        m.dt_bias = _dev(sd, prefix + "mamba.dt_bias", device)
        self.pseudoLTI = bool(cfg.get("pseudoLTI", False))
        if self.pseudoLTI:                                            # SSD_LTI (models/mamba.py:156-299): A, beta instead of A_log
            m.A = _dev(sd, prefix + "mamba.A", device)
            m.beta = _dev(sd, prefix + "mamba.beta", device)
            if m.beta is None:
                m.beta = torch.ones_like(m.A)                         # register_buffer("beta", ones) (:228)
            if (N * m.ngroups) % m.nheads != 0:
                raise RuntimeError("SSD_LTI: d_state * ngroups must be divisible by nheads (models/mamba.py:200)")
            m.khead_dim = (N * m.ngroups) // m.nheads
        m.A_log = _dev(sd, prefix + "mamba.A_log", device)
        m.D = _dev(sd, prefix + "mamba.D", device)
        cw = _dev(sd, prefix + "mamba.conv1d.weight", device)
        m.conv_w = cw.reshape(cw.shape[0], -1).contiguous() if cw is not None else None
        m.conv_b = _dev(sd, prefix + "mamba.conv1d.bias", device)
        m.use_conv = cw is not None
        # the dt rows of in_proj: split order [x | B | C | dt] (models/mamba.py:62-63)
        lo = d_inner + 2 * m.ngroups * N
        m.W_dt = m.in_proj.weight[lo:lo + (m.ngroups if self.pseudoLTI else m.nheads)].contiguous()
        self.mamba = m
        gw = _dev(sd, prefix + "glu.linear.weight", device)
        self.glu = SimpleNamespace(linear=_Lin(gw, _dev(sd, prefix + "glu.linear.bias", device))) if gw is not None else None
        self.norm = SimpleNamespace(weight=_dev(sd, prefix + "norm.weight", device), bias=_dev(sd, prefix + "norm.bias", device))
        self.gemm_mode = cfg.get("_gemm_mode", "auto")
        # the parameters are constants of an analysis run: their tensor-core operand form is built once per layer (ops.linear_prepare), not per call
        self.prepare_weights = os.environ.get("EIGB200_PREPARE_WEIGHTS", "1") != "0"
        self.fuse_front = os.environ.get("EIGB200_FUSE_FRONT", "1") != "0"    # LayerNorm + in_proj + conv + SSD scan as one kernel
        self.fuse_tail = os.environ.get("EIGB200_FUSE_TAIL", "1") != "0"      # out_proj + GELU + GLU + residual (+ extractor partials) as one kernel
        self._prep_ws = {}

    def _prepared(self, key, weight, bias, epilogue, gamma=None, beta=None):
        """Per-layer prepared workspace of one GEMM (None when its shape has no resident-weight plan).  invalidate_prepared() after changing parameters."""
        key = (key, ops.gemm_precision())                           # the operand form is precision-specific (tf32 hi / lo or scaled fp16 hi / lo)
        if key not in self._prep_ws:
            self._prep_ws[key] = ops.linear_prepare(weight, bias, epilogue, gamma, beta)
        return self._prep_ws[key]

    def invalidate_prepared(self):
        self._prep_ws = {}

    def fuses_layernorm(self):
        """True when the prenorm LayerNorm can ride inside the in_proj GEMM (row statistics supplied by the producer of x)."""
        m = self.mamba
        return (self.prenorm and not self.pseudoLTI and self.gemm_mode in ("auto", "tc3") and m.d_model % 4 == 0 and
                ops.linear_ln_supported(m.in_proj.weight.shape[0], m.d_model))

    def _ssd_lti(self, z, ldz, B, T):
        """SSD_LTI.forward after in_proj (models/mamba.py:262-295): conv + SiLU of [x | B | C], B <- softplus(dt + dt_bias) * B with the single dt
        column tiled over the state columns, then the scan with dt := beta and A := -softplus(A)."""
        m = self.mamba
        di, N, H = m.d_inner, m.d_state, m.nheads
        Cn = di + 2 * m.ngroups * N
        if m.conv_w is not None:
            zc = z.clone()
            ops.conv_silu(z, ldz, m.conv_w, m.conv_b, B, T, Cn, out=zc, ldo=ldz)
        else:
            zc = z.clone()
        ops.lti_scale_b(zc, ldz, di, Cn, m.dt_bias, N, m.khead_dim)
        beta = m.beta.reshape(1, 1, H).expand(B, T, H).contiguous()
        A = -torch.nn.functional.softplus(m.A)                        # parameter-sized host-side prep (:275)
        return ops.ssd_scan_buffer(zc, ldz, 0, di, di + m.ngroups * N, beta, A, m.D, B, T, H, m.headdim, m.ngroups, N)

    def fuses_extractor(self):
        """True when the eigenvalue extractor of this block's OUTPUT can ride in the GLU + residual GEMM's epilogue (one head, prenorm, GLU present,
        d_model / 16 a power of two <= 16, tensor-core GEMM): x_out is then never re-read."""
        m = self.mamba
        return (self.fuses_layernorm() and self.glu is not None and m.nheads == 1 and not self.pseudoLTI and m.d_model % 16 == 0 and
                (m.d_model // 16) in (2, 4, 8, 16))

    def __call__(self, x, stats=None, extract_partials=None):
        """MambaBlock.forward (models/mamba.py:328-340) with SSD.forward (:111-154) inlined.
        stats: optional (B,T,2) LayerNorm (mean, rstd) of x from the kernel that produced x (embedding / previous extractor);
        with it the normalised activations are formed inside the in_proj GEMM and never written to HBM.
        extract_partials: optional (d_model/16, 3, B*T) buffer; the GLU epilogue then also leaves the extractor partials of the output rows in it."""
        m = self.mamba
        B, T, D = x.shape
        skip = x
        d_in_proj = m.in_proj.weight.shape[0]
        ldz = _pad8(d_in_proj)
        tc = self.gemm_mode in ("auto", "tc3") and B * T >= 1024 and self.prepare_weights
        y = None
        if (stats is not None and self.fuses_layernorm() and tc and self.fuse_front and m.conv_w is not None and
                ops.mamba_front_fused_supported(D, m.d_inner, m.nheads, m.ngroups, m.d_state, m.conv_w.shape[1])):
            # LayerNorm -> in_proj -> conv + SiLU -> SSD scan (:329-331, :118-150) in one kernel: the projection never reaches HBM
            y = ops.mamba_front_fused(x, stats, self._prepared("in_ln", m.in_proj.weight, None, "none", self.norm.weight, self.norm.bias),
                                      m.conv_w, m.conv_b, m.dt_bias, m.A_log, m.D, m.d_inner, m.d_state)
        elif stats is not None and self.fuses_layernorm():
            z = ops.linear_ln(x, stats, self.norm.weight, self.norm.bias, m.in_proj.weight, None, ldc=ldz,
                              prepared=self._prepared("in_ln", m.in_proj.weight, None, "none", self.norm.weight, self.norm.bias) if tc else None)
        else:
            xn = ops.layernorm(x, self.norm.weight, self.norm.bias) if self.prenorm else x
            z = ops.linear(xn, m.in_proj.weight, None, ldc=ldz, mode=self.gemm_mode)              # (B*T, ldz) = [x | B | C | dt | pad]
        if y is not None:
            pass
        elif self.pseudoLTI:
            y = self._ssd_lti(z, ldz, B, T)
        else:
            y = ops.mamba_conv_ssd(z, ldz, m.conv_w, m.conv_b, m.dt_bias, m.A_log, m.D, B, T, m.nheads, m.headdim, m.ngroups, m.d_state)
        want_x = self.glu is not None and extract_partials is not None and self.fuses_extractor() and B * T >= 1024
        if (tc and self.glu is not None and self.prenorm and self.fuse_tail and ops.out_glu_fused_supported(D, m.d_inner) and D == m.out_proj.weight.shape[0]):
            # GELU(out_proj(y)) -> GLU + skip (:333-337) in one kernel: the intermediate never reaches HBM (fp16-split precision, d_model 128)
            out, _ = ops.out_glu_fused(y, self._prepared("out", m.out_proj.weight, m.out_proj.bias, "gelu"), m.out_proj.bias,
                                       self._prepared("glu", self.glu.linear.weight, self.glu.linear.bias, "glu_residual"), self.glu.linear.bias,
                                       skip.reshape(B * T, D), m.W_dt[0] if want_x else None, extract_partials if want_x else None)
            return out.reshape(B, T, D)
        o = ops.linear(y, m.out_proj.weight, m.out_proj.bias, epilogue="gelu", mode=self.gemm_mode,
                       prepared=self._prepared("out", m.out_proj.weight, m.out_proj.bias, "gelu") if tc else None)    # GELU(out_proj(y))  (:333)
        if want_x:
            out, _ = ops.linear_glu_extract(o, self.glu.linear.weight, self.glu.linear.bias, skip.reshape(B * T, D), m.W_dt[0], partials=extract_partials,
                                            prepared=self._prepared("glu", self.glu.linear.weight, self.glu.linear.bias, "glu_residual") if tc else None)
        elif self.glu is not None:
            out = ops.linear(o, self.glu.linear.weight, self.glu.linear.bias, epilogue="glu_residual",
                             residual=skip.reshape(B * T, D), mode=self.gemm_mode,
                             prepared=self._prepared("glu", self.glu.linear.weight, self.glu.linear.bias, "glu_residual") if tc else None)   # GLU + skip (:335-337)
        else:
            out = ops.add(o, skip.reshape(B * T, D))
        out = out.reshape(B, T, D)
        if not self.prenorm:
            out = ops.layernorm(out, self.norm.weight, self.norm.bias)
        return out


class TokenEmbeddingsDev:
    def __init__(self, sd, prefix, device):
        self.word = _dev(sd, prefix + "word_embeddings.weight", device)
        self.pos = _dev(sd, prefix + "position_embeddings.weight", device)

    def __call__(self, ids, rowstats_out=None):
        return ops.embedding(ids, self.word, self.pos, rowstats_out=rowstats_out)


class LinearEncoderDev:
    def __init__(self, sd, prefix, device):
        self.lin = _Lin(_dev(sd, prefix + "weight", device), _dev(sd, prefix + "bias", device))

    def __call__(self, x):
        out = ops.linear(x.float(), self.lin.weight, self.lin.bias)
        return out.reshape(x.shape[:-1] + (out.shape[-1],))


def _make_encoder(sd, device):
    if "encoder.word_embeddings.weight" in sd:
        return TokenEmbeddingsDev(sd, "encoder.", device)
    return LinearEncoderDev(sd, "encoder.", device)


class MambaDev:
    """Mirror of models/mamba.py:Mamba for the analysis loop: .encoder, .blocks[i]."""

    def __init__(self, cfg, state_dict, device="cuda"):
        if cfg.get("version", "mamba2") != "mamba2":
            raise RuntimeError("Non supported version")                               # models/mamba.py:313 (Mamba-1 is train-only)
        self.cfg = dict(cfg)
        self.encoder = _make_encoder(state_dict, device)
        self.blocks = [MambaBlockDev(state_dict, "blocks.%d." % i, cfg, device) for i in range(cfg["num_layers"])]

    def invalidate_prepared(self):
        """Drop the per-layer prepared GEMM operands (call after changing any parameter tensor in place; a captured MambaPassGraph must be rebuilt)."""
        for blk in self.blocks:
            blk.invalidate_prepared()


# ======================================================================================================================
# Transformer (linear attention / normalised attention)
# ======================================================================================================================

class AttentionDev:
    """MHA(lin_att=True, use_flash=False) or MHNA parameters + forward."""

    def __init__(self, sd, prefix, cfg, device):
        self.d_model = cfg["hidden_dim"]
        self.d_qk = cfg["state_dim"]
        self.num_heads = cfg["num_heads"]
        self.head_dim = self.d_qk // self.num_heads
        self.v_dim = self.d_model // self.num_heads
        self.kind = cfg["attention_fn"]
        self.conv_type = cfg.get("conv_type", "full")
        if self.kind == "norm-attention":
            self.Wvqkn = _Lin(_dev(sd, prefix + "Wvqkn.weight", device), _dev(sd, prefix + "Wvqkn.bias", device))
            self.inner_attn = SimpleNamespace(offset=_dev(sd, prefix + "inner_attn.offset", device))
            self.norm_fn = cfg["norm_fn"]
            self.approx_fn = cfg["approx_fn"]
            self.scale_B = cfg["scale_B"]
            lo = self.d_model + 2 * self.d_qk
            self.W_n = self.Wvqkn.weight[lo:lo + self.num_heads].contiguous()
            self.b_n = self.Wvqkn.bias[lo:lo + self.num_heads].contiguous()
            # the layer forward projects only [v | q | k]: the H gate columns come from eigb200_normattn_gate (fp32 dot products, the extractor's arithmetic), so
            # the GEMM drops them -- at C5 that is N = 1536 = 6 tiles of 256 columns instead of 1544 = 7 tiles (one of them 97 % padding)
            self.W_vqk = self.Wvqkn.weight[:lo].contiguous()
            self.b_vqk = self.Wvqkn.bias[:lo].contiguous() if self.Wvqkn.bias is not None else None
        elif self.kind in ("lin-attention", "sm-attention"):
            self.Wqkv = _Lin(_dev(sd, prefix + "Wqkv.weight", device), _dev(sd, prefix + "Wqkv.bias", device))
            self.W_qk = self.Wqkv.weight[: 2 * self.d_qk].contiguous()
            self.b_qk = self.Wqkv.bias[: 2 * self.d_qk].contiguous() if self.Wqkv.bias is not None else None
        else:
            raise RuntimeError("{0} is not a valid model option".format(self.kind))
        self.out_proj = _Lin(_dev(sd, prefix + "out_proj.weight", device), _dev(sd, prefix + "out_proj.bias", device))
        cw = _dev(sd, prefix + "conv1d.weight", device)
        self.conv_w = cw.reshape(cw.shape[0], -1).contiguous() if cw is not None else None
        self.conv_b = _dev(sd, prefix + "conv1d.bias", device)

    def _conv(self, buf, ld, col0, ncols, B, T):
        """conv+SiLU over columns [col0, col0+ncols) of buf (B*T, ld), in a fresh buffer that keeps the other columns."""
        out = buf.clone() if ncols < buf.shape[1] else torch.empty_like(buf)
        ops.conv_silu(buf[:, col0:], ld, self.conv_w, self.conv_b, B, T, ncols, out=out[:, col0:], ldo=ld)
        return out

    def _fuse_conv(self, buf, ld, q_off, k_off, v_off, H, d, dv):
        """The conv + SiLU rides in the chunked attention kernel when that kernel takes the shape (EIGB200_LINATTN_FORM=col / EIGB200_LINATTN_CONV=split: no)."""
        if os.environ.get("EIGB200_LINATTN_FORM") == "col" or os.environ.get("EIGB200_LINATTN_CONV") == "split":
            return False
        return ops.linattn_conv_fusable(buf, ld, q_off, k_off, v_off, H, d, dv, int(self.conv_w.shape[1]))

    def forward_residual(self, xn, skip):
        """out_proj(attention(xn)) + skip   (MHA.forward models/attention.py:149-182 / MHNA.forward norm_attention.py:230-258)."""
        B, T, D = xn.shape
        H, d, dv, dqk = self.num_heads, self.head_dim, self.v_dim, self.d_qk
        if self.kind == "lin-attention":
            ld = 2 * dqk + D
            buf = ops.linear(xn, self.Wqkv.weight, self.Wqkv.bias)
            if self.conv_w is not None and self._fuse_conv(buf, ld, 0, dqk, 2 * dqk, H, d, dv):
                # conv + SiLU inside the attention kernel's tile loader: q at conv channel 0, k at dqk, v at 2 dqk ("full") or untouched
                ctx = ops.linattn_forward_conv(buf, ld, 0, dqk, 2 * dqk, B, T, H, d, dv, self.conv_w, self.conv_b, 0, dqk,
                                               2 * dqk if self.conv_type == "full" else -1, phi_elu=True, normalise=True)
            else:
                if self.conv_w is not None:
                    buf = self._conv(buf, ld, 0, ld if self.conv_type == "full" else 2 * dqk, B, T)
                ctx = ops.linattn_forward(buf, ld, 0, dqk, 2 * dqk, B, T, H, d, dv, phi_elu=True, normalise=True)
        elif self.kind == "norm-attention":
            ld = D + 2 * dqk
            buf = ops.linear(xn, self.W_vqk, self.b_vqk)
            gate = ops.normattn_gate(xn, self.W_n, self.b_n, self.inner_attn.offset, self.norm_fn)
            kscale = 1.0 / math.sqrt(d) if self.scale_B else 1.0
            if self.conv_w is not None and self._fuse_conv(buf, ld, D, D + dqk, 0, H, d, dv):
                # buffer columns [v | q | k]; "full": conv channel = column, else the conv covers q, k only (channel = column - D)
                full = self.conv_type == "full"
                ctx = ops.linattn_forward_conv(buf, ld, D, D + dqk, 0, B, T, H, d, dv, self.conv_w, self.conv_b,
                                               D if full else 0, D + dqk if full else dqk, 0 if full else -1,
                                               gate=gate, phi_elu=(self.approx_fn == "elu"), normalise=False, kscale=kscale)
            else:
                if self.conv_w is not None:
                    if self.conv_type == "full":
                        buf = self._conv(buf, ld, 0, D + 2 * dqk, B, T)
                    else:
                        buf = self._conv(buf, ld, D, 2 * dqk, B, T)
                ctx = ops.linattn_forward(buf, ld, D, D + dqk, 0, B, T, H, d, dv, gate=gate, phi_elu=(self.approx_fn == "elu"),
                                          normalise=False, kscale=kscale)
        elif self.kind == "sm-attention":
            # SelfAttention.forward (models/attention.py:14-35), the float32 "naive" path; with use_flash the reference rounds q,k,v to fp16 first
            ld = 2 * dqk + D
            buf = ops.linear(xn, self.Wqkv.weight, self.Wqkv.bias)
            if self.conv_w is not None:
                buf = self._conv(buf, ld, 0, ld if self.conv_type == "full" else 2 * dqk, B, T)
            ctx = ops.softmax_attn_forward(buf, ld, 0, dqk, 2 * dqk, B, T, H, d, dv, 1.0 / math.sqrt(d))
        else:
            raise RuntimeError("{0} is not a valid model option".format(self.kind))
        out = ops.linear(ctx, self.out_proj.weight, self.out_proj.bias, epilogue="residual", residual=skip.reshape(B * T, D))
        return out.reshape(B, T, D)


class TransformerBlockDev:
    def __init__(self, sd, prefix, cfg, device):
        self.attention = AttentionDev(sd, prefix + "attention.", cfg, device)
        self.norm = SimpleNamespace(weight=_dev(sd, prefix + "norm.weight", device), bias=_dev(sd, prefix + "norm.bias", device))
        self.mixer_kind = cfg["mixer"]
        self.use_gate = cfg.get("use_gate", False)
        if self.use_gate:
            self.Wz = _Lin(_dev(sd, prefix + "Wz.weight", device), _dev(sd, prefix + "Wz.bias", device))
        if self.mixer_kind == "mlp":
            self.mixer = SimpleNamespace(encoder=_Lin(_dev(sd, prefix + "mixer.encoder.weight", device), _dev(sd, prefix + "mixer.encoder.bias", device)),
                                         decoder=_Lin(_dev(sd, prefix + "mixer.decoder.weight", device), _dev(sd, prefix + "mixer.decoder.bias", device)))
        elif self.mixer_kind == "glu":
            self.mixer = SimpleNamespace(linear=_Lin(_dev(sd, prefix + "mixer.linear.weight", device), _dev(sd, prefix + "mixer.linear.bias", device)))
        elif self.mixer_kind == "hybrid":
            # LAMBDA (models/common.py:60-84): a * glu(xz) + (1 - a) * decoder(gelu(xz)), xz = encoder(y), a = sigmoid(alpha).  The scalar blend is
            # folded into the weights once: value rows of the encoder scaled by a (GLU epilogue then yields a * glu), decoder scaled by 1 - a.
            enc_w, enc_b = _dev(sd, prefix + "mixer.encoder.weight", device), _dev(sd, prefix + "mixer.encoder.bias", device)
            dec_w, dec_b = _dev(sd, prefix + "mixer.decoder.weight", device), _dev(sd, prefix + "mixer.decoder.bias", device)
            a = torch.sigmoid(_dev(sd, prefix + "mixer.alpha", device).reshape(()))
            Dm = enc_w.shape[1]
            sc = torch.cat([a.expand(Dm), torch.ones(Dm, device=enc_w.device)])
            self.mixer = SimpleNamespace(encoder=_Lin(enc_w, enc_b), glu_scaled=_Lin((enc_w * sc[:, None]).contiguous(), (enc_b * sc).contiguous()),
                                         decoder_scaled=_Lin((dec_w * (1 - a)).contiguous(), (dec_b * (1 - a)).contiguous()))
        elif self.mixer_kind == "none":
            self.mixer = None
        else:
            raise RuntimeError("{0} mixer not implemented yet!".format(self.mixer_kind))     # models/transformer.py:76-77

    def __call__(self, x):
        """TransformerBlock.forward (models/transformer.py:90-111): one LayerNorm reused for both sub-blocks."""
        B, T, D = x.shape
        z = ops.linear(x, self.Wz.weight, self.Wz.bias).reshape(B, T, D) if self.use_gate else None
        xn = ops.layernorm(x, self.norm.weight, self.norm.bias)
        x = self.attention.forward_residual(xn, x)
        y = ops.layernorm(x, self.norm.weight, self.norm.bias)
        if self.mixer_kind == "none":                                                     # drop_skip (:73-75, :102-104)
            return ops.mul_silu(y, z) if z is not None else y
        x2 = x.reshape(B * T, D)
        if self.mixer_kind == "glu":
            out = ops.linear(y, self.mixer.linear.weight, self.mixer.linear.bias, epilogue="glu_residual", residual=x2)
        elif self.mixer_kind == "hybrid":
            part = ops.linear(y, self.mixer.glu_scaled.weight, self.mixer.glu_scaled.bias, epilogue="glu_residual", residual=x2)   # x + a * glu(xz)
            hmid = ops.linear(y, self.mixer.encoder.weight, self.mixer.encoder.bias, epilogue="gelu")                             # gelu(xz)
            out = ops.linear(hmid, self.mixer.decoder_scaled.weight, self.mixer.decoder_scaled.bias, epilogue="residual", residual=part)
        else:
            hmid = ops.linear(y, self.mixer.encoder.weight, self.mixer.encoder.bias, epilogue="gelu")
            out = ops.linear(hmid, self.mixer.decoder.weight, self.mixer.decoder.bias, epilogue="residual", residual=x2)
        out = out.reshape(B, T, D)
        return ops.mul_silu(out, z) if z is not None else out


class TransformerDev:
    """Mirror of models/transformer.py:Transformer for the analysis loop: .encoder, .layers[i]."""

    def __init__(self, cfg, state_dict, device="cuda"):
        self.cfg = dict(cfg)
        self.encoder = _make_encoder(state_dict, device)
        self.layers = [TransformerBlockDev(state_dict, "layers.%d." % i, cfg, device) for i in range(cfg["num_layers"])]


# ======================================================================================================================
# initial parameters: the reference draws them from torch's CPU generator in constructor order (eval_eig.py:484-497)
# ======================================================================================================================

def init_mamba_state_dict(cfg, seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Parameters of models/mamba.py:Mamba(cfg) at construction, drawn in the reference's order so that
    torch.manual_seed(seed) gives the same tensors (embedding; per block: in_proj, dt, A, conv1d, out_proj, GLU; decoder)."""
    if seed is not None:
        torch.manual_seed(seed)
    D = cfg["hidden_dim"]; N = cfg["state_dim"]; k = cfg["conv_dim"]
    headdim = D // cfg["num_heads"]; d_inner = cfg["expansion"] * D; H = d_inner // headdim; G = 1
    sd = {}
    if cfg["token_embedding"]:
        sd["encoder.word_embeddings.weight"] = nn.Embedding(cfg["vocab_size"], D).weight.detach()
    else:
        enc = nn.Linear(cfg["input_dim"], D)
        sd["encoder.weight"], sd["encoder.bias"] = enc.weight.detach(), enc.bias.detach()
    for i in range(cfg["num_layers"]):
        p = "blocks.%d." % i
        lti = bool(cfg.get("pseudoLTI", False))
        sd[p + "mamba.in_proj.weight"] = nn.Linear(D, d_inner + 2 * G * N + (G if lti else H), bias=False).weight.detach()   # SSD_LTI: ngroups dt columns (:204)
        dt = torch.exp(torch.rand(H) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))        # models/mamba.py:71-77
        dt = torch.clamp(dt, min=1e-4)
        sd[p + "mamba.dt_bias"] = dt + torch.log(-torch.expm1(-dt))
        if lti:                                                                                     # :221-228 (constructor order: dt, D, A, beta)
            sd[p + "mamba.D"] = torch.ones(H)
            sd[p + "mamba.A"] = torch.empty(H, dtype=torch.float32).uniform_(-8, -2)
            sd[p + "mamba.beta"] = torch.ones(H)
        else:
            sd[p + "mamba.A_log"] = torch.log(torch.empty(H, dtype=torch.float32).uniform_(1, 16)) # :85-86
            sd[p + "mamba.D"] = torch.ones(H)
        if k > 0:
            conv = nn.Conv1d(d_inner + 2 * G * N, d_inner + 2 * G * N, kernel_size=k, groups=d_inner + 2 * G * N, padding=k - 1)
            sd[p + "mamba.conv1d.weight"], sd[p + "mamba.conv1d.bias"] = conv.weight.detach(), conv.bias.detach()
        sd[p + "mamba.out_proj.weight"] = nn.Linear(d_inner, D, bias=False).weight.detach()
        if cfg["glu"]:
            glu = nn.Linear(D, 2 * D)
            sd[p + "glu.linear.weight"], sd[p + "glu.linear.bias"] = glu.weight.detach(), glu.bias.detach()
        sd[p + "norm.weight"], sd[p + "norm.bias"] = torch.ones(D), torch.zeros(D)
    dec = nn.Linear(D, cfg["output_dim"])
    sd["decoder.weight"], sd["decoder.bias"] = dec.weight.detach(), dec.bias.detach()
    return sd


def init_transformer_state_dict(cfg, seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Parameters of models/transformer.py:Transformer(cfg) at construction, in the reference's drawing order."""
    if seed is not None:
        torch.manual_seed(seed)
    D = cfg["hidden_dim"]; dqk = cfg["state_dim"]; H = cfg["num_heads"]
    sd = {}
    if cfg["embedding"]:
        sd["encoder.word_embeddings.weight"] = nn.Embedding(cfg["vocab_size"], D).weight.detach()
        if cfg["max_pos_embed"] > 0:
            sd["encoder.position_embeddings.weight"] = nn.Embedding(cfg["max_pos_embed"], D).weight.detach()
    else:
        enc = nn.Linear(cfg["input_dim"], D)
        sd["encoder.weight"], sd["encoder.bias"] = enc.weight.detach(), enc.bias.detach()
    for i in range(cfg["num_layers"]):
        p = "layers.%d." % i
        kind = cfg["attention_fn"]
        dim_conv = cfg.get("dim_conv", 0)
        conv_type = cfg.get("conv_type", "full")
        if kind == "norm-attention":
            lin = nn.Linear(D, D + 2 * dqk + H)
            sd[p + "attention.Wvqkn.weight"], sd[p + "attention.Wvqkn.bias"] = lin.weight.detach(), lin.bias.detach()
            if cfg["offset"]:
                if cfg["offset_init"] == "exp":
                    sd[p + "attention.inner_attn.offset"] = torch.linspace(4, 9, H)               # norm_attention.py:52
                elif cfg["offset_init"] == "uniform":
                    if H == 1:
                        off = torch.tensor([(14. - 8.) / 2])                                       # norm_attention.py:17-19
                    else:
                        x = torch.log(torch.expm1(torch.linspace(0.02, 0.1, H)))
                        x = (x - x.min()) / (x.max() - x.min())
                        off = x * abs(14. - 8.) + 8.
                    sd[p + "attention.inner_attn.offset"] = off
                else:
                    raise RuntimeError("Invalid init option {0}".format(cfg["offset_init"]))
        else:
            lin = nn.Linear(D, 2 * dqk + D)
            sd[p + "attention.Wqkv.weight"], sd[p + "attention.Wqkv.bias"] = lin.weight.detach(), lin.bias.detach()
        op = nn.Linear(D, D)
        sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"] = op.weight.detach(), op.bias.detach()
        if dim_conv > 0:
            cd = D + 2 * dqk if conv_type == "full" else 2 * dqk
            conv = nn.Conv1d(cd, cd, kernel_size=dim_conv, groups=cd, padding=dim_conv - 1)
            sd[p + "attention.conv1d.weight"], sd[p + "attention.conv1d.bias"] = conv.weight.detach(), conv.bias.detach()
        if cfg.get("use_gate", False):
            wz = nn.Linear(D, D)
            nn.init.constant_(wz.bias, 1.0)
            nn.init.xavier_uniform_(wz.weight, gain=0.1)
            sd[p + "Wz.weight"], sd[p + "Wz.bias"] = wz.weight.detach(), wz.bias.detach()
        if cfg["mixer"] == "mlp":
            e = nn.Linear(D, cfg["mixer_dim"]); dcd = nn.Linear(cfg["mixer_dim"], D)
            sd[p + "mixer.encoder.weight"], sd[p + "mixer.encoder.bias"] = e.weight.detach(), e.bias.detach()
            sd[p + "mixer.decoder.weight"], sd[p + "mixer.decoder.bias"] = dcd.weight.detach(), dcd.bias.detach()
        elif cfg["mixer"] == "glu":
            gl = nn.Linear(D, 2 * D)
            sd[p + "mixer.linear.weight"], sd[p + "mixer.linear.bias"] = gl.weight.detach(), gl.bias.detach()
        elif cfg["mixer"] == "hybrid":                                                              # LAMBDA(hidden_dim, init=0.2), transformer.py:75-77
            e = nn.Linear(D, 2 * D); dcd = nn.Linear(2 * D, D)
            sd[p + "mixer.alpha"] = torch.ones(1) * (-math.log(1 / 0.2 - 1))
            sd[p + "mixer.encoder.weight"], sd[p + "mixer.encoder.bias"] = e.weight.detach(), e.bias.detach()
            sd[p + "mixer.decoder.weight"], sd[p + "mixer.decoder.bias"] = dcd.weight.detach(), dcd.bias.detach()
        sd[p + "norm.weight"], sd[p + "norm.bias"] = torch.ones(D), torch.zeros(D)
    if cfg["classifier"]:
        if cfg["mixer_dim"] != 0:
            e = nn.Linear(D, cfg["mixer_dim"]); dcd = nn.Linear(cfg["mixer_dim"], cfg["output_dim"])
            sd["classifier.encoder.weight"], sd["classifier.encoder.bias"] = e.weight.detach(), e.bias.detach()
            sd["classifier.decoder.weight"], sd["classifier.decoder.bias"] = dcd.weight.detach(), dcd.bias.detach()
    else:
        sd["decoder.weight"] = nn.Linear(D, cfg["output_dim"], bias=False).weight.detach()
    sd["norm.weight"], sd["norm.bias"] = torch.ones(D), torch.zeros(D)
    return sd
