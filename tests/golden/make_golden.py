#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN CODE on CPU.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

Nothing from the reference is copied into this repository: its functions / classes are pulled out of
the files where they lie with `ast` (or `importlib` when a file imports cleanly) and executed in a
namespace that supplies the missing third-party pieces:

  * `jnp` := numpy, `inv`/`matrix_power` := numpy.linalg   (jax is not installed)
  * `torch` := a proxy that drops `device='cuda'` keyword arguments (eval_eig.py:109-110 hard-codes it)
  * `mamba_chunk_scan_combined` := fla.ops.simple_gla.naive.naive_recurrent_simple_gla with the mapping
    q=C, k=B, v=dt*x, g=dt*A, scale=1, + D*x -- an INDEPENDENT third-party implementation of the SSD
    recurrence (mamba-ssm itself is not installed), so the Mamba goldens do not depend on our own oracle.

The script also cross-checks the oracle's SSD restatement against HF transformers' Mamba2 torch path.
Outputs: small .npz files + MANIFEST.json (shapes, seeds, reference file:line of what produced them).
"""
import ast
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import einops

REF = os.environ.get("EIGB200_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
MANIFEST = {}


# ------------------------------------------------------------------------------------------------
# helpers: pull definitions out of reference files without importing them
# ------------------------------------------------------------------------------------------------

def extract(path, names, namespace):
    """exec only the FunctionDef/ClassDef nodes called `names` from `path` inside `namespace`."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    assert not missing, missing
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, os.path.join(REF, path), "exec"), namespace)
    return namespace


class _TorchCpuProxy:
    """`torch` with any device= keyword dropped, so the reference's hard-coded 'cuda' lands on CPU."""
    def __getattr__(self, name):
        attr = getattr(torch, name)
        if callable(attr) and name in ("ones", "zeros", "full", "arange", "tril", "triu", "empty", "rand"):
            def wrapped(*a, **k):
                k.pop("device", None)
                return attr(*a, **k)
            return wrapped
        return attr


def save(name, cite, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    MANIFEST[name] = {"cite": cite, "arrays": {k: [list(np.shape(v)), str(np.asarray(v).dtype)] for k, v in arrays.items()}}
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def sd_numpy(module, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


# ------------------------------------------------------------------------------------------------
# the reference's analysis functions
# ------------------------------------------------------------------------------------------------

EV = extract("analysis/eval_eig.py",
             ["get_eig_att_softmax", "get_eig_att_linear", "get_eig_att_norm", "get_eig_mamba2", "get_eig_mamba2_LTI",
              "discrete_DPLR", "get_eigvals_ssm", "threshold_analysis", "threshold_analysis_ssm"],
             {"torch": _TorchCpuProxy(), "F": F, "einops": einops, "np": np, "jnp": np,
              "inv": np.linalg.inv, "matrix_power": np.linalg.matrix_power})


def gold_thresholds():
    rng = np.random.default_rng(7)
    thr_r = np.array([0.1, 0.5, 0.9, 1.0, 10, 100])
    thr_p = np.array([1, 10, 45, 90, 180])
    B, N, H, L = 5, 37, 2, 3
    v32 = np.exp(rng.normal(0, 2.5, (B, N, H, L))).astype(np.float32)
    # plant exact-edge, negative, nan, inf values (double counting / dropped values)
    flat = v32.reshape(-1)
    plant = np.array([0.1, 0.5, 0.9, 1.0, 10, 100, 0.0, -0.25, np.nan, np.inf, np.float32(0.1), np.nextafter(np.float32(0.5), np.float32(1))], np.float32)
    flat[: plant.size] = plant
    v64 = np.exp(rng.normal(0, 2.5, (B, N, H, L)))
    v64.reshape(-1)[:10] = [0.1, 0.5, 0.9, 1.0, 10, 100, 0.0, -3.0, np.nan, np.inf]
    ph = rng.uniform(-180, 180, (B, N, H, L)).astype(np.float32)
    ph.reshape(-1)[:6] = [1, 10, 45, 90, 180, 0]
    out = dict(v32=v32, v64=v64, ph=ph)
    with np.errstate(invalid="ignore"):
        out["p_v32"] = EV["threshold_analysis"](v32, thr_r, L, H, B)
        out["p_v64"] = EV["threshold_analysis"](v64, thr_r, L, H, B)
        out["p_ph"] = EV["threshold_analysis"](ph, thr_p, L, H, B)
        out["p_half"] = EV["threshold_analysis"](np.full((2, 8, 1, 1), 0.5), thr_r, 1, 1, 2)
        s = np.exp(rng.normal(0, 1.5, (64, 4)))
        out["s"] = s
        out["p_s"] = EV["threshold_analysis_ssm"](s, thr_r, 4)
        sp = rng.uniform(-180, 180, (64, 4))
        out["sp"] = sp
        out["p_sp"] = EV["threshold_analysis_ssm"](sp, thr_p, 4)
    save("thresholds", "analysis/eval_eig.py:335-391 executed with numpy %s" % np.__version__, **out)


class _NS(types.SimpleNamespace):
    pass


def gold_mamba2_extractor():
    torch.manual_seed(11)
    B, T, D, H, G, N = 3, 29, 48, 4, 1, 8
    d_in_proj = D + 2 * G * N + H
    lin = nn.Linear(D, d_in_proj, bias=False)
    layer = _NS(mamba=_NS(in_proj=lin, A_log=torch.log(torch.empty(H).uniform_(1, 16)),
                          dt_bias=torch.randn(H) * 2 - 3, d_inner=D, ngroups=G, d_state=N, nheads=H))
    x = torch.randn(B, T, D) * 1.5
    x[0, 0] *= 30.0          # drive softplus through its threshold-20 branch
    lam = EV["get_eig_mamba2"](x, layer)
    save("mamba2_extractor", "analysis/eval_eig.py:176-190", x=x.numpy(), in_proj_weight=lin.weight.detach().numpy(),
         A_log=layer.mamba.A_log.numpy(), dt_bias=layer.mamba.dt_bias.numpy(), dims=np.array([D, G, N, H]), lam=lam)
    # LTI
    layer2 = _NS(mamba=_NS(in_proj=lin, A=torch.empty(H).uniform_(-8, -2), beta=torch.ones(H)))
    lam2 = EV["get_eig_mamba2_LTI"](x, layer2)
    save("mamba2_lti_extractor", "analysis/eval_eig.py:192-205", A=layer2.mamba.A.numpy(), beta=layer2.mamba.beta.numpy(),
         shape=np.array([B, T]), lam=lam2)


def gold_norm_extractor():
    torch.manual_seed(12)
    B, T, D, dqk, H = 3, 21, 32, 16, 4
    lin = nn.Linear(D, D + 2 * dqk + H, bias=True)
    with torch.no_grad():
        lin.weight[-H:] *= 6.0
    x = torch.randn(B, T, D) * 2
    x[1, 3] *= 40.0          # exp(-exp(big)) underflows to 0 in fp32 -> the 2e-23 patch
    offs = torch.linspace(4, 9, H)
    arrays = dict(x=x.numpy(), weight=lin.weight.detach().numpy(), bias=lin.bias.detach().numpy(), offset=offs.numpy(),
                  dims=np.array([D, dqk, H]))
    for fn in ["exp", "elu", "softplus", "sigmoid"]:
        for use_off in (False, True):
            layer = _NS(attention=_NS(Wvqkn=lin, inner_attn=_NS(offset=offs)))
            cfg = {"norm_fn": fn, "approx_fn": "elu", "offset": use_off}
            with np.errstate(all="ignore"):
                eta = EV["get_eig_att_norm"](x, layer, dqk, H, D, cfg)
            arrays["eta_%s_%d" % (fn, int(use_off))] = eta
    save("norm_extractor", "analysis/eval_eig.py:137-174", **arrays)


def gold_lin_softmax_extractor():
    torch.manual_seed(13)
    B, T, D, dqk, H = 3, 24, 32, 32, 4
    lin = nn.Linear(D, 2 * dqk + D, bias=True)
    x = torch.randn(B, T, D) * 1.7
    layer = _NS(attention=_NS(Wqkv=lin, head_dim=dqk // H))
    with np.errstate(all="ignore"):
        eta_lin = EV["get_eig_att_linear"](x, layer, dqk, H, D)
        eta_sm = EV["get_eig_att_softmax"](x, layer, dqk, H, D)
    save("lin_softmax_extractor", "analysis/eval_eig.py:43-135 (device='cuda' at :109-110 redirected to CPU)",
         x=x.numpy(), weight=lin.weight.detach().numpy(), bias=lin.bias.detach().numpy(), dims=np.array([D, dqk, H]),
         eta_lin=eta_lin, eta_sm=eta_sm)


def gold_ssm_eigs():
    rng = np.random.default_rng(14)
    P, L = 24, 3
    lru = [dict(nu_log=rng.normal(-1.5, 0.7, P).astype(np.float32), theta_log=rng.normal(-0.5, 1.0, P).astype(np.float32)) for _ in range(L)]
    s5 = [dict(Lambda_re=(-np.abs(rng.normal(0.5, 0.3, P))).astype(np.float32), Lambda_im=rng.normal(0, 8, P).astype(np.float32),
               log_step=rng.uniform(np.log(1e-3), np.log(1e-1), (P, 1)).astype(np.float32)) for _ in range(L)]
    arrays = {}
    for i in range(L):
        arrays["lru_nu_%d" % i] = lru[i]["nu_log"]; arrays["lru_theta_%d" % i] = lru[i]["theta_log"]
        arrays["lru_eig_%d" % i] = EV["get_eigvals_ssm"]("lru", lru, i, 1, 100)
        for k in s5[i]:
            arrays["s5_%s_%d" % (k, i)] = s5[i][k]
        arrays["s5_eig_%d" % i] = EV["get_eigvals_ssm"]("s5", s5, i, 1, 100)
    save("lru_s5_eigs", "analysis/eval_eig.py:303-329 (jnp := numpy)", **arrays)


def gold_hippo_s4():
    ns = extract("models/common.py", ["make_HiPPO", "make_NPLR_HiPPO", "make_DPLR_HiPPO"], {"np": np, "eigh": np.linalg.eigh})
    s5ns = extract("models/s5.py", ["discretize_bilinear", "discretize_zoh"], {"jnp": np})
    arrays = {}
    for N in (8, 16, 64):
        Lam, P, Bv, V, Bo = ns["make_DPLR_HiPPO"](N)
        arrays["hippo_%d" % N] = ns["make_HiPPO"](N)
        arrays["Lambda_%d" % N] = Lam; arrays["P_%d" % N] = P; arrays["B_%d" % N] = Bv
    rng = np.random.default_rng(15)
    # S4 layers: vmapped over features on axis 1 (models/s4.py:183-189); complex64 P/B as in JAX
    for N, Hf in ((8, 3), (16, 3), (64, 2)):
        Lam, P, Bv, _, _ = ns["make_DPLR_HiPPO"](N)
        layers = []
        for l in range(2):
            pert = 0.0 if l == 0 else 0.1
            layer = dict(
                Lambda_re=(np.repeat(Lam.real[:, None], Hf, 1) + pert * rng.normal(size=(N, Hf))).astype(np.float32),
                Lambda_im=(np.repeat(Lam.imag[:, None], Hf, 1) + pert * rng.normal(size=(N, Hf))).astype(np.float32),
                P=(np.repeat(P[:, None], Hf, 1) + pert * (rng.normal(size=(N, Hf)) + 1j * rng.normal(size=(N, Hf)))).astype(np.complex64),
                B=np.repeat(Bv[:, None], Hf, 1).astype(np.complex64),
                C=rng.normal(size=(N, Hf, 2)).astype(np.float32),
                log_step=rng.uniform(np.log(1e-3), np.log(1e-1), (1, Hf)).astype(np.float32))
            layers.append(layer)
        for l in range(2):
            for k, v in layers[l].items():
                arrays["s4_N%d_l%d_%s" % (N, l, k)] = v
            # A-bar straight from the reference's discrete_DPLR on the same slices get_eigvals_ssm takes
            lay = layers[l]; idx = 1
            step = np.exp(lay["log_step"][0, idx])
            Lm = np.clip(lay["Lambda_re"][:, idx], None, -1e-4) + 1j * lay["Lambda_im"][:, idx]
            Ct = lay["C"][:, idx, 0] + 1j * lay["C"][:, idx, 1]
            Ab, _, _ = EV["discrete_DPLR"](Lm, lay["P"][:, idx], lay["P"][:, idx], lay["B"][:, idx], Ct, step, 64)
            arrays["s4_N%d_l%d_Abar" % (N, l)] = Ab
            arrays["s4_N%d_l%d_eig" % (N, l)] = EV["get_eigvals_ssm"]("s4", layers, l, idx, 64)
    # S5 discretisations
    Pn = 12
    Lam = (-np.abs(rng.normal(0.5, 0.3, Pn)) + 1j * rng.normal(0, 5, Pn)).astype(np.complex64)
    Bt = (rng.normal(size=(Pn, 5)) + 1j * rng.normal(size=(Pn, 5))).astype(np.complex64)
    dl = rng.uniform(1e-3, 1e-1, Pn).astype(np.float32)
    arrays["disc_Lambda"] = Lam; arrays["disc_B"] = Bt; arrays["disc_Delta"] = dl
    arrays["zoh_L"], arrays["zoh_B"] = s5ns["discretize_zoh"](Lam, Bt, dl)
    arrays["bil_L"], arrays["bil_B"] = s5ns["discretize_bilinear"](Lam, Bt, dl)
    save("hippo_s4", "models/common.py:180-241; analysis/eval_eig.py:254-301; models/s5.py:16-47 (jnp := numpy)", **arrays)


# ------------------------------------------------------------------------------------------------
# whole models: the reference classes, the reference per-layer loop
# ------------------------------------------------------------------------------------------------

def fla_ssd(x, dt, A, B, C, chunk_size=None, D=None, z=None, seq_idx=None, initial_states=None, **kw):
    """Stand-in for mamba_ssm's mamba_chunk_scan_combined built on fla's naive recurrent simple-GLA."""
    from fla.ops.simple_gla.naive import naive_recurrent_simple_gla
    assert z is None and seq_idx is None and initial_states is None
    b, l, h, p = x.shape
    g = B.shape[2]
    rep = h // g
    q = C.repeat_interleave(rep, dim=2)
    k = B.repeat_interleave(rep, dim=2)
    v = x * dt[..., None]
    gk = dt * A
    try:
        o, _ = naive_recurrent_simple_gla(q, k, v, gk, scale=1.0)
    except TypeError:
        o, _ = naive_recurrent_simple_gla(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), gk.transpose(1, 2), scale=1.0)
        o = o.transpose(1, 2)
    o = o.to(x.dtype)
    if D is not None:
        o = o + D[None, None, :, None] * x
    return o


def load_reference_models():
    common = extract("models/common.py", ["MATCH", "MLP", "GLU", "LAMBDA", "ClassifierHead", "TokenEmbeddings"],
                     {"torch": torch, "nn": nn, "F": F, "math": __import__("math")})
    spec = importlib.util.spec_from_file_location("ref_attention", os.path.join(REF, "models/attention.py"))
    att = importlib.util.module_from_spec(spec); spec.loader.exec_module(att)
    for name in ["mamba_ssm", "mamba_ssm.ops", "mamba_ssm.ops.selective_scan_interface", "mamba_ssm.ops.triton",
                 "mamba_ssm.ops.triton.ssd_combined"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mamba_ssm.ops.selective_scan_interface"].selective_scan_fn = None
    sys.modules["mamba_ssm.ops.triton.ssd_combined"].mamba_chunk_scan_combined = fla_ssd
    spec = importlib.util.spec_from_file_location("ref_norm_attention", os.path.join(REF, "models/norm_attention.py"))
    natt = importlib.util.module_from_spec(spec); spec.loader.exec_module(natt)
    ns = dict(common)
    ns.update({"torch": torch, "nn": nn, "F": F, "math": __import__("math"), "rearrange": einops.rearrange,
               "repeat": einops.repeat, "MHA": att.MHA, "MHNA": natt.MHNA, "mamba_chunk_scan_combined": fla_ssd,
               "Tensor": torch.Tensor, "Optional": __import__("typing").Optional})
    extract("models/transformer.py", ["TransformerBlock", "Transformer"], ns)
    extract("models/mamba.py", ["SSD", "MambaBlock", "Mamba"], ns)
    return ns


def reference_layer_loop(model, layers, X, extractor):
    """eval_eig.py:501-526 / :575-600, verbatim control flow: extractor applied to each block's OUTPUT."""
    x = model.encoder(X)
    acts = [x.detach().numpy()]
    eig = None
    for i, layer in enumerate(layers):
        x = layer(x)
        acts.append(x.detach().numpy())
        e = extractor(x, layer)
        eig = e if eig is None else np.concatenate((eig, e), axis=-1)
    return eig, acts


def gold_models(only=None):
    ns = load_reference_models()
    if only is None:
        _gold_mamba_model(ns)
    _gold_transformer_models(ns, only)


def _gold_mamba_model(ns):
    # ---- Mamba-2 (models/mamba.py) ----
    torch.manual_seed(1919)
    cfg = dict(version="mamba2", num_layers=3, input_dim=1, output_dim=64, hidden_dim=32, num_heads=2, state_dim=8,
               conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", prenorm=True, dual=False, pooling="none",
               token_embedding=True, vocab_size=64)
    model = ns["Mamba"](cfg).eval()
    with torch.no_grad():       # move the dt rows off their tiny init so lambda spreads over several bins
        for blk in model.blocks:
            blk.mamba.in_proj.weight[-blk.mamba.nheads:] *= 4.0
    X = torch.randint(0, 64, (8, 40))
    with torch.no_grad():
        eig, acts = reference_layer_loop(model, list(model.blocks), X, EV["get_eig_mamba2"])
    arrays = {"sd::" + k: v for k, v in sd_numpy(model).items()}
    arrays.update(X=X.numpy(), eig=eig, **{"act_%d" % i: a for i, a in enumerate(acts)})
    arrays["cfg_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    with np.errstate(invalid="ignore"):
        arrays["percentage"] = EV["threshold_analysis"](np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2)),
                                                        np.array([0.1, 0.5, 0.9, 1.0, 10, 100]), 3, 2, 8)
    save("model_mamba2", "models/mamba.py:25-154,301-389 + analysis/eval_eig.py:501-526 (SSD kernel := fla naive recurrence)", **arrays)



def _gold_transformer_models(ns, only=None):
    # ---- Transformer, linear attention (C1-like), normalised attention and softmax attention ----
    base = dict(input_dim=1, output_dim=64, num_layers=2, hidden_dim=32, embedding=True, vocab_size=64, max_pos_embed=24,
                pooling="none", dual=False, classifier=False, mixer_dim=64, norm="layer", dropout=0.0, state_dim=32,
                num_heads=2, att_dropout=0.0, use_flash=False)
    variants = {
        "model_linattn": dict(base, attention_fn="lin-attention", mixer="none"),
        "model_linattn_glu_conv": dict(base, attention_fn="lin-attention", mixer="glu", dim_conv=4),
        "model_normattn": dict(base, attention_fn="norm-attention", mixer="mlp", mode="attention", norm_fn="softplus",
                               approx_fn="elu", scale_B=False, offset=True, offset_init="exp", learn_A=False, dim_conv=4),
        "model_normattn_exp": dict(base, attention_fn="norm-attention", mixer="none", mode="attention", norm_fn="exp",
                                   approx_fn="none", scale_B=True, offset=False, offset_init="uniform", learn_A=False, dim_conv=0,
                                   max_pos_embed=0),
        "model_smattn": dict(base, attention_fn="sm-attention", mixer="mlp"),
        "model_linattn_hybrid": dict(base, attention_fn="lin-attention", mixer="hybrid"),
    }
    for name, c in variants.items():
        if only is not None and name not in only:
            continue
        torch.manual_seed(1919)
        model = ns["Transformer"](dict(c)).eval()
        model.encoder.device = "cpu"          # TokenEmbeddings hard-codes device='cuda' for position ids (common.py:126)
        X = torch.randint(0, 64, (8, 24))
        dqk, H, D = c["state_dim"], c["num_heads"], c["hidden_dim"]
        if c["attention_fn"] == "lin-attention":
            ext = lambda x, layer: EV["get_eig_att_linear"](x, layer, dqk, H, D)
        elif c["attention_fn"] == "sm-attention":
            ext = lambda x, layer: EV["get_eig_att_softmax"](x, layer, dqk, H, D)
        else:
            ext = lambda x, layer, c=c: EV["get_eig_att_norm"](x, layer, dqk, H, D, c)
        with torch.no_grad(), np.errstate(all="ignore"):
            eig, acts = reference_layer_loop(model, list(model.layers), X, ext)
        arrays = {"sd::" + k: v for k, v in sd_numpy(model).items()}
        arrays.update(X=X.numpy(), eig=eig, **{"act_%d" % i: a for i, a in enumerate(acts)})
        arrays["cfg_json"] = np.frombuffer(json.dumps(c).encode(), dtype=np.uint8)
        with np.errstate(invalid="ignore"):
            arrays["percentage"] = EV["threshold_analysis"](eig, np.array([0.1, 0.5, 0.9, 1.0, 10, 100]), 2, H, 8)
            arrays["percentage_phase"] = EV["threshold_analysis"](0 * eig, np.array([1, 10, 45, 90, 180]), 2, H, 8)
        save(name, "models/transformer.py:22-161, attention.py / norm_attention.py + analysis/eval_eig.py:528-564", **arrays)


def gold_c1():
    """BASELINE config C1 at its exact shapes: linear attention on MQAR-shaped tokens, seq 64, d_model = d_qk = 64, 1 head, 2 layers, vocab 8192,
    analysis batch 8.  The 4 MB of embedding / decoder weights are NOT stored: the model is the reference's own construction under
    torch.manual_seed(1919) (eval_eig.py:484-497), which eigb200.layers.init_transformer_state_dict reproduces bit for bit -- the fixture keeps a
    SHA-256 of every parameter so that the test proves it, plus the token ids, the activations after every block and the eigenvalue array."""
    import hashlib
    ns = load_reference_models()
    cfg = dict(input_dim=1, output_dim=8192, num_layers=2, hidden_dim=64, embedding=True, vocab_size=8192, max_pos_embed=64, pooling="none", dual=False,
               classifier=False, mixer_dim=128, norm="layer", dropout=0.0, state_dim=64, num_heads=1, att_dropout=0.0, use_flash=False,
               attention_fn="lin-attention", mixer="none")
    torch.manual_seed(1919)
    model = ns["Transformer"](dict(cfg)).eval()
    model.encoder.device = "cpu"
    rng = np.random.default_rng(42)
    X = torch.from_numpy(rng.integers(0, 8192, (8, 64)))
    ext = lambda x, layer: EV["get_eig_att_linear"](x, layer, 64, 1, 64)
    with torch.no_grad(), np.errstate(all="ignore"):
        eig, acts = reference_layer_loop(model, list(model.layers), X, ext)
    hashes = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in sd_numpy(model).items()}
    arrays = dict(X=X.numpy(), eig=eig, **{"act_%d" % i: a for i, a in enumerate(acts)})
    arrays["cfg_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    arrays["sd_sha256_json"] = np.frombuffer(json.dumps(hashes).encode(), dtype=np.uint8)
    with np.errstate(invalid="ignore"):
        arrays["percentage"] = EV["threshold_analysis"](eig, np.array([0.1, 0.5, 0.9, 1.0, 10, 100]), 2, 1, 8)
    save("c1_linattn_mqar", "BASELINE configs[0]: models/transformer.py + attention.py + analysis/eval_eig.py:528-552 at seq 64, d_model 64, 2 layers", **arrays)


def crosscheck_ssd_against_hf():
    """Independent check of the oracle's SSD restatement against HF transformers' pure-torch Mamba2 step
    (modeling_mamba2.py torch_forward): same parameters, same input, outputs compared.  Result recorded in MANIFEST."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ssd_scan_sequential, ssd_scan_chunked
    torch.manual_seed(5)
    b, l, h, p, g, n = 2, 64, 4, 8, 1, 16
    x = torch.randn(b, l, h, p); dt = F.softplus(torch.randn(b, l, h) - 1); A = -torch.empty(h).uniform_(1, 16)
    Bm = torch.randn(b, l, g, n); Cm = torch.randn(b, l, g, n); D = torch.randn(h)
    y_or = ssd_scan_sequential(x.numpy(), dt.numpy(), A.numpy(), Bm.numpy(), Cm.numpy(), D.numpy())
    y_fla = fla_ssd(x, dt, A, Bm, Cm, D=D).numpy()
    y_ch = ssd_scan_chunked(x, dt, A, Bm, Cm, D, chunk=16).numpy()
    # HF-style single-step recurrence (state*exp(dt*A) + (dt*B) (x) x ; y = C.state + D*x), written against their variable flow
    st = torch.zeros(b, h, p, n, dtype=torch.float64)
    ys = []
    for t in range(l):
        dA = torch.exp(dt[:, t].double() * A.double())
        dB = dt[:, t].double()[..., None] * Bm[:, t].double().repeat_interleave(h // g, 1)
        st = st * dA[..., None, None] + dB[:, :, None, :] * x[:, t].double()[..., None]
        ys.append(torch.einsum("bhpn,bhn->bhp", st, Cm[:, t].double().repeat_interleave(h // g, 1)) + D.double()[None, :, None] * x[:, t].double())
    y_hf = torch.stack(ys, 1).numpy()
    res = dict(oracle_vs_fla=float(np.abs(y_or - y_fla).max()), oracle_vs_hfstep=float(np.abs(y_or - y_hf).max()),
               chunked_vs_oracle=float(np.abs(y_or - y_ch).max()), scale=float(np.abs(y_or).max()))
    MANIFEST["ssd_crosscheck"] = res
    print("ssd crosscheck", res)
    save("ssd_small", "models/mamba.py:138-150 call site; outputs of fla naive recurrence (third-party) for the same inputs",
         x=x.numpy(), dt=dt.numpy(), A=A.numpy(), Bm=Bm.numpy(), Cm=Cm.numpy(), D=D.numpy(), y_fla=y_fla)


def gold_report_files():
    """The text reports written by create_file_percentage / create_file_percentage_ssm (eval_eig.py:393-459), byte for byte."""
    ns = extract("analysis/eval_eig.py", ["create_file_percentage", "create_file_percentage_ssm"], {"np": np})
    rng = np.random.default_rng(21)
    B, H, L = 8, 2, 3
    pct = rng.uniform(0, 100, (7, B, H, L)); pct_i = rng.uniform(0, 100, (7, B, H, L))
    thr = np.array([0.1, 0.5, 0.9, 1.0, 10, 100]); thp = np.array([1, 10, 45, 90, 180])
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            ns["create_file_percentage"](thr, pct, pct_i, pct.mean(1), pct_i.mean(1), pct.std(1), pct_i.std(1))
            txt1 = open("percentage_file.txt").read()
            ps = rng.uniform(0, 100, (7, L)); psi = rng.uniform(0, 100, (7, L)); pp = rng.uniform(0, 100, (6, L)); ppi = rng.uniform(0, 100, (6, L))
            ns["create_file_percentage_ssm"](thr, thp, ps, psi, pp, ppi)
            txt2 = open("percentage_file.txt").read()
        finally:
            os.chdir(cwd)
    save("report_files", "analysis/eval_eig.py:393-459", pct=pct, pct_i=pct_i, ps=ps, psi=psi, pp=pp, ppi=ppi,
         txt_torch=np.frombuffer(txt1.encode(), dtype=np.uint8), txt_ssm=np.frombuffer(txt2.encode(), dtype=np.uint8))


if __name__ == "__main__":
    if "--only-c1" in sys.argv:
        MANIFEST.update(json.load(open(os.path.join(OUT, "MANIFEST.json"))))
        gold_c1()
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
            json.dump(MANIFEST, f, indent=1, sort_keys=True)
        sys.exit(0)
    if "--only-models" in sys.argv:                 # (re)generate the named transformer models without touching the other vectors
        MANIFEST.update(json.load(open(os.path.join(OUT, "MANIFEST.json"))))
        gold_models(only=sys.argv[sys.argv.index("--only-models") + 1].split(","))
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
            json.dump(MANIFEST, f, indent=1, sort_keys=True)
        sys.exit(0)
    if "--only-reports" in sys.argv:
        MANIFEST.update(json.load(open(os.path.join(OUT, "MANIFEST.json"))))
        gold_report_files()
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
            json.dump(MANIFEST, f, indent=1, sort_keys=True)
        sys.exit(0)
    gold_thresholds()
    gold_mamba2_extractor()
    gold_norm_extractor()
    gold_lin_softmax_extractor()
    gold_ssm_eigs()
    gold_hippo_s4()
    crosscheck_ssd_against_hf()
    gold_models()
    gold_c1()
    gold_report_files()
    MANIFEST["_env"] = {"numpy": np.__version__, "torch": torch.__version__}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(MANIFEST, f, indent=1, sort_keys=True)
