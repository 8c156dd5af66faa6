"""LRU / S5 / S4 side of the analysis (analysis/eval_eig.py:207-333, :684-745) and the diagonal-recurrence layer calls
(models/lru.py:86-99; models/s5.py:65-93, :236-250) on the eigb200 kernels.

The Flax modules cannot be instantiated here (no JAX), so layers are plain dicts with the Flax parameter names
(`params["encoder"]["layers_i"]["seq"]`, eval_eig.py:234-237) holding numpy arrays or torch tensors.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List

import numpy as np
import torch

from . import _lib as L
from . import ops


def _t(a, dtype=torch.float32):
    if isinstance(a, torch.Tensor):
        return a.detach().to("cuda", dtype)
    return torch.as_tensor(np.asarray(a)).to("cuda", dtype)


# ---- eigenvalues -----------------------------------------------------------------------------------------------------
def get_eigvals_ssm(model, layer_list, layer_nr, idx, SEQ_LEN):
    """analysis/eval_eig.py:281-333.  Returns a HOST complex64 array (P,1) like the reference."""
    if model in ["s4"]:
        layer = layer_list[layer_nr]
        lam_re = np.asarray(layer["Lambda_re"])[:, idx]
        lam_im = np.asarray(layer["Lambda_im"])[:, idx]
        Lam = (np.minimum(lam_re, np.float32(-1e-4)) + 1j * lam_im).astype(np.complex64)          # :288 clip
        Pv = np.asarray(layer["P"])[:, idx].astype(np.complex64)
        step = np.exp(np.asarray(layer["log_step"], np.float32)[0, idx])
        Ab = ops.dplr_abar(_t(Lam[None], torch.complex64), _t(Pv[None], torch.complex64), _t(Pv[None], torch.complex64),
                           _t(np.array([step], np.float32)))
        ev, info = ops.eigvals_c64(Ab)
        if int(info.max()) != 0:
            raise L.Eigb200Error("eigvals_c64: QR iteration did not converge for %d eigenvalue(s)" % int(info.max()))
        return np.expand_dims(ev[0].cpu().numpy(), axis=-1)
    if model in ["s5"]:
        layer = layer_list[layer_nr]
        lam = ops.ssm_lambda("s5_zoh", _t(layer["Lambda_re"]), _t(layer["Lambda_im"]), _t(np.asarray(layer["log_step"]).reshape(-1)))
        return np.expand_dims(lam.cpu().numpy(), axis=-1)
    elif model in ["lru"]:
        layer = layer_list[layer_nr]
        lam = ops.ssm_lambda("lru", _t(layer["nu_log"]), _t(layer["theta_log"]))
        return np.expand_dims(lam.cpu().numpy(), axis=-1)
    else:
        print("model type {0} is not supported!".format(model))
        return None


def eigvals_s4_all_features(layer_list):
    """Batched variant: every feature of every layer at once -> (L, H, N) complex64 on the device."""
    Ls, Ps, steps = [], [], []
    for layer in layer_list:
        lam = np.minimum(np.asarray(layer["Lambda_re"]), np.float32(-1e-4)) + 1j * np.asarray(layer["Lambda_im"])
        Ls.append(lam.T.astype(np.complex64)); Ps.append(np.asarray(layer["P"]).T.astype(np.complex64))
        steps.append(np.exp(np.asarray(layer["log_step"], np.float32)[0]))
    Lam = _t(np.concatenate(Ls), torch.complex64); Pv = _t(np.concatenate(Ps), torch.complex64)
    Ab = ops.dplr_abar(Lam, Pv, Pv, _t(np.concatenate(steps)))
    ev, info = ops.eigvals_c64(Ab)
    nl = len(layer_list)
    return ev.reshape(nl, -1, ev.shape[-1]), info.reshape(nl, -1), Ab


def radius_phase(eig):
    """|lambda| and arg in degrees exactly as eval_eig.py:726-727, :734-735 (host NumPy on the (P, L) array)."""
    eig = np.asarray(eig)
    rad = np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2))
    ph = np.arctan2(eig.imag, eig.real) * 180 / np.pi
    return rad, ph


# ---- parameters ------------------------------------------------------------------------------------------------------
def _make_dplr_hippo(N):
    """HiPPO-LegS in DPLR form (models/common.py:180-241), NumPy."""
    p = np.sqrt(1 + 2 * np.arange(N))
    A = -(np.tril(p[:, None] * p[None, :]) - np.diag(np.arange(N)))
    Pl = np.sqrt(np.arange(N) + 0.5)
    Bv = np.sqrt(2 * np.arange(N) + 1.0)
    S = A + Pl[:, None] * Pl[None, :]
    lam_re = np.mean(np.diagonal(S)) * np.ones(N)
    lam_im, V = np.linalg.eigh(S * -1j)
    return lam_re + 1j * lam_im, V.conj().T @ Pl, V.conj().T @ Bv, V


def get_init_layers_ssm(seed, data_config, train_config, model_config, SEQ_LEN, layer_type, batch_size) -> List[Dict[str, np.ndarray]]:
    """Initial `seq` parameters of every layer (analysis/eval_eig.py:207-239).  The reference draws them from JAX's PRNG, which
    cannot be reproduced without JAX: the DETERMINISTIC parts (HiPPO Lambda / P / B of S4 and S5) are identical, the random
    parts follow the same distributions from numpy.random.default_rng(seed) (eig_init parity is distributional)."""
    rng = np.random.default_rng(seed)
    P, H, nl = model_config["state_dim"], model_config["hidden_dim"], model_config["num_layers"]
    layers = []
    for _ in range(nl):
        if layer_type == "lru":
            r_min, r_max, max_phase = model_config.get("r_min", 0.0), model_config.get("r_max", 1.0), model_config.get("max_phase", 6.28)
            u1, u2 = rng.uniform(size=P), rng.uniform(size=P)
            nu_log = np.log(-0.5 * np.log(u1 * (r_max ** 2 - r_min ** 2) + r_min ** 2))          # models/lru.py:26-28
            theta_log = np.log(max_phase * u2)                                                     # :31-33
            lam_abs2 = np.exp(-2 * np.exp(nu_log))
            layers.append(dict(nu_log=nu_log.astype(np.float32), theta_log=theta_log.astype(np.float32),
                               gamma_log=np.log(np.sqrt(1 - lam_abs2)).astype(np.float32),
                               B_re=(rng.normal(size=(P, H)) / np.sqrt(2 * H)).astype(np.float32),
                               B_im=(rng.normal(size=(P, H)) / np.sqrt(2 * H)).astype(np.float32),
                               C_re=(rng.normal(size=(H, P)) / np.sqrt(P)).astype(np.float32),
                               C_im=(rng.normal(size=(H, P)) / np.sqrt(P)).astype(np.float32),
                               D=rng.normal(size=H).astype(np.float32)))
        elif layer_type == "s5":
            blocks = model_config.get("num_blocks", 8)
            conj_sym = model_config.get("conj_sym", True)
            bs = int(P / blocks)
            Lam, _, _, V = _make_dplr_hippo(bs)                                                    # models/s5.py:264-284
            Pe = P
            if conj_sym:
                bs //= 2; Pe = P // 2
            Lam = np.tile(Lam[:bs], blocks)
            dt_min, dt_max = model_config.get("dt_min", 0.001), model_config.get("dt_max", 0.1)
            local_P = 2 * Pe if conj_sym else Pe
            layers.append(dict(Lambda_re=Lam.real.astype(np.float32), Lambda_im=Lam.imag.astype(np.float32),
                               B=(rng.normal(size=(Pe, H, 2)) / np.sqrt(local_P)).astype(np.float32),
                               C=(rng.normal(size=(H, Pe, 2)) / np.sqrt(local_P)).astype(np.float32),
                               D=rng.normal(size=H).astype(np.float32),
                               log_step=(rng.uniform(size=(Pe, 1)) * (np.log(dt_max) - np.log(dt_min)) + np.log(dt_min)).astype(np.float32)))
        elif layer_type == "s4":
            Lam, Pv, Bv, _ = _make_dplr_hippo(P)                                                   # models/s4.py:192-215
            dt_min, dt_max = model_config.get("dt_min", 0.001), model_config.get("dt_max", 0.1)
            layers.append(dict(Lambda_re=np.repeat(Lam.real[:, None], H, 1).astype(np.float32),
                               Lambda_im=np.repeat(Lam.imag[:, None], H, 1).astype(np.float32),
                               P=np.repeat(Pv[:, None], H, 1).astype(np.complex64), B=np.repeat(Bv[:, None], H, 1).astype(np.complex64),
                               C=(rng.normal(size=(P, H, 2)) * 0.5 ** 0.5).astype(np.float32), D=np.ones((1, H), np.float32),
                               log_step=(rng.uniform(size=(1, H)) * (np.log(dt_max) - np.log(dt_min)) + np.log(dt_min)).astype(np.float32)))
        else:
            raise RuntimeError("{0} is not a valid model option".format(layer_type))
    return layers


def _layers_from_params(params):
    layers = []
    for name in params["encoder"]:
        if name.startswith("layers"):
            layers.append((name, params["encoder"][name]["seq"]))
    layers.sort(key=lambda kv: int(kv[0].split("_")[-1]))                  # orbax restores dict keys sorted as strings
    return [v for _, v in layers]


def _unflatten(flat):
    tree = {}
    for key, val in flat.items():
        node = tree
        parts = key.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return tree


def get_trained_layers_ssm(path):
    """analysis/eval_eig.py:241-252.  Accepts the reference's orbax PyTree checkpoint when orbax is importable, and otherwise a
    mirror of the same tree: a `.npz` with '/'-joined keys (model/params/encoder/layers_0/seq/nu_log ...) or a torch-saved
    nested dict {"model": {"params": ...}} of tensors (a directory may hold `params.npz`).
    A checkpoint path is user input: nothing here unpickles arbitrary objects (np.load(allow_pickle=False),
    torch.load(weights_only=True)) -- the reference restores through orbax, which does not execute code either.  Legacy pickle /
    non-tensor torch files load only with EIGB200_ALLOW_PICKLE=1 in the environment."""
    if os.path.isdir(path) and not os.path.exists(os.path.join(path, "params.npz")):
        try:
            import orbax.checkpoint as ocp
        except ImportError as e:
            raise L.Eigb200Error("'%s' looks like an orbax checkpoint but orbax is not installed; export it as params.npz" % path) from e
        raw = ocp.PyTreeCheckpointer().restore(path)
        return _layers_from_params(raw["model"]["params"])
    if os.path.isdir(path):
        path = os.path.join(path, "params.npz")
    allow_pickle = os.environ.get("EIGB200_ALLOW_PICKLE", "0") == "1"
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        tree = _unflatten({k: z[k] for k in z.files})
    elif path.endswith((".pkl", ".pickle")):
        if not allow_pickle:
            raise L.Eigb200Error("'%s': pickle checkpoints execute code on load and are refused; export the parameter tree as .npz "
                                 "(or set EIGB200_ALLOW_PICKLE=1 for a file you trust)" % path)
        with open(path, "rb") as f:
            tree = pickle.load(f)
    else:
        tree = torch.load(path, weights_only=not allow_pickle, map_location="cpu")
    params = tree["model"]["params"] if "model" in tree else tree.get("params", tree)
    return _layers_from_params(params)


# ---- layer calls: the diagonal recurrence the eigenvalues drive -----------------------------------------------------------
def _interleave_rows(re, im):
    """(P,H),(P,H) -> (2P,H) rows [re_0, im_0, re_1, im_1, ...] so that a real GEMM emits interleaved complex64."""
    P, H = re.shape
    return torch.stack([re, im], dim=1).reshape(2 * P, H).contiguous()


def lru_forward(params, u, return_states=False):
    """LRU.__call__ (models/lru.py:86-99) for a batch u (B,T,d_model) float32 on the device."""
    u = _t(u)
    B, T, Hd = u.shape
    lam = ops.ssm_lambda("lru", _t(params["nu_log"]), _t(params["theta_log"]))
    gamma = torch.exp(_t(params["gamma_log"]))[:, None]
    Wb = _interleave_rows(_t(params["B_re"]) * gamma, _t(params["B_im"]) * gamma)                # B_norm (lru.py:89)
    Bu = ops.linear(u, Wb)                                                                        # (B*T, 2P) == complex64 (B,T,P)
    P = lam.shape[0]
    h = ops.diag_scan(lam, torch.view_as_complex(Bu.reshape(B, T, P, 2)))
    Wc = torch.stack([_t(params["C_re"]), -_t(params["C_im"])], dim=2).reshape(Hd, 2 * P).contiguous()   # Re(C h)
    Du = ops.scale_cols(u.reshape(B * T, Hd), _t(params["D"]))
    y = ops.linear(torch.view_as_real(h).reshape(B * T, 2 * P), Wc, None, epilogue="residual", residual=Du).reshape(B, T, Hd)
    return (y, h) if return_states else y


def s5_forward(params, u, discretization="zoh", conj_sym=True, clip_eigs=False, bidirectional=False, return_states=False):
    """S5SSM.__call__ / apply_ssm (models/s5.py:65-93, :141-250) for a batch u (B,T,H)."""
    u = _t(u)
    B, T, Hd = u.shape
    lre = _t(params["Lambda_re"])
    if clip_eigs:
        lre = torch.clamp(lre, max=-1e-4)
    lim = _t(params["Lambda_im"]); log_step = _t(np.asarray(params["log_step"]).reshape(-1) if not isinstance(params["log_step"], torch.Tensor) else params["log_step"].reshape(-1))
    lam_bar = ops.ssm_lambda("s5_zoh" if discretization == "zoh" else "s5_bilinear", lre, lim, log_step)
    Lam = torch.complex(lre, lim)
    step = torch.exp(log_step)
    Bt = _t(params["B"]); Bt = torch.complex(Bt[..., 0], Bt[..., 1])
    if discretization == "zoh":
        B_bar = ((lam_bar - 1.0) / Lam)[:, None] * Bt                                             # s5.py:46
    elif discretization == "bilinear":
        B_bar = ((1.0 / (1.0 - (step / 2.0) * Lam)) * step)[:, None] * Bt                         # s5.py:27-30
    else:
        raise NotImplementedError("Discretization method {} not implemented".format(discretization))
    P = lam_bar.shape[0]
    Wb = _interleave_rows(B_bar.real.contiguous(), B_bar.imag.contiguous())
    Bu = torch.view_as_complex(ops.linear(u, Wb).reshape(B, T, P, 2))
    h = ops.diag_scan(lam_bar, Bu)
    if bidirectional:
        h = torch.cat([h, ops.diag_scan(lam_bar, Bu, reverse=True)], dim=-1)
        C1 = _t(params["C1"]); C2 = _t(params["C2"])
        Cre = torch.cat([C1[..., 0], C2[..., 0]], dim=-1); Cim = torch.cat([C1[..., 1], C2[..., 1]], dim=-1)
    else:
        Cc = _t(params["C"]); Cre, Cim = Cc[..., 0], Cc[..., 1]
    scale = 2.0 if conj_sym else 1.0
    PP = h.shape[-1]
    Wc = (scale * torch.stack([Cre, -Cim], dim=2)).reshape(Hd, 2 * PP).contiguous()
    Du = ops.scale_cols(u.reshape(B * T, Hd), _t(params["D"]))
    y = ops.linear(torch.view_as_real(h.contiguous()).reshape(B * T, 2 * PP), Wc, None, epilogue="residual", residual=Du).reshape(B, T, Hd)
    return (y, h) if return_states else y


def s4_forward(layer, u, return_kernel=False):
    """S4.__call__ in CNN mode (models/s4.py:169-173; vmapped over features, :182-188) for a batch u (B,T,H) float32 on the device:
    y[:, :, h] = causal_convolution(u[:, :, h], K_h) + D_h u[:, :, h], K_h = kernel_DPLR(clip(Lambda_re) + i Lambda_im, P, P, B, C~, exp(log_step), T)
    (:107-139).  layer: the vmapped parameter dict (feature axis 1), as returned by get_init_layers_ssm / get_trained_layers_ssm."""
    u = _t(u)
    Bsz, T, H = u.shape
    lam = (np.minimum(np.asarray(layer["Lambda_re"], np.float32), np.float32(-1e-4)) + 1j * np.asarray(layer["Lambda_im"], np.float32)).T
    Pv = np.asarray(layer["P"]).T.astype(np.complex64)
    Bv = np.asarray(layer["B"]).T.astype(np.complex64)
    Cf = np.asarray(layer["C"], np.float32)
    Cv = (Cf[..., 0] + 1j * Cf[..., 1]).T.astype(np.complex64)
    step = np.exp(np.asarray(layer["log_step"], np.float32)).reshape(-1)
    if not (lam.shape[0] == H and step.shape[0] == H):
        raise L.Eigb200Error("s4_forward: parameters describe %d features, input has %d" % (lam.shape[0], H))
    Kt = ops.s4_kernel(_t(np.ascontiguousarray(lam.astype(np.complex64)), torch.complex64), _t(np.ascontiguousarray(Pv), torch.complex64),
                       _t(np.ascontiguousarray(Pv), torch.complex64), _t(np.ascontiguousarray(Bv), torch.complex64),
                       _t(np.ascontiguousarray(Cv), torch.complex64), _t(step), T)
    y = ops.s4_causal_conv(u, Kt, _t(np.asarray(layer["D"], np.float32).reshape(-1)))
    return (y, Kt) if return_kernel else y
