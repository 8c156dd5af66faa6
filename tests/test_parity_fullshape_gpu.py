"""Full-shape parity (VERDICT r1 "what's weak" 1-2): the BASELINE configs at their real sequence lengths and depths on sampled sequences against the fp64
oracle, eta against the fp64 oracle under a conditioning-aware 1e-5 bound (the loose 3e-4 stays only against the reference's own fp32 goldens, with the
measured worst case printed), S5 at the C3 layer shape, and a 2-GPU eval_eig run against the 1-GPU run."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch

import oracle as O
from conftest import assert_eig_close, golden_model, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eig():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import eigb200.analysis as A
    import eigb200.extractors as E
    import eigb200.layers as Ly
    import eigb200.ops as ops
    import eigb200.ssm as S
    return A, Ly, E, S, ops


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------------------------------------------------
# eta against the fp64 oracle, conditioning-aware
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fn", ["exp", "elu", "softplus", "sigmoid"])
@pytest.mark.parametrize("use_off", [0, 1])
def test_normattn_eta_vs_fp64_conditioning_bound(eig, fn, use_off):
    """eta_t = n_{t+1} / n_t, n = exp(-norm_fn(z)), z = x . W_n + b + offset (eval_eig.py:137-174).  With z formed in fp32 (|dz| <= eps * (sum |x||w| + |b| +
    |offset|)) the relative error of eta is  |f'(z_t)| dz_t + |f'(z_{t+1})| dz_{t+1}:  the bound below is the north star's 1e-5 times (1 + that condition
    number), against the fp64 evaluation of the same formula.  Against the reference's own fp32 output only the loose bound can hold (its n carries the
    same error); the measured worst case is printed."""
    A, Ly, E, S, ops = eig
    g = load_golden("norm_extractor")
    D, dqk, H = [int(v) for v in g["dims"]]
    rows = O.normattn_rows(D, dqk, H)
    W, b, off = g["weight"][rows].astype(np.float64), g["bias"][rows].astype(np.float64), g["offset"].astype(np.float64) if use_off else 0.0
    n = ops.normattn_gate(_dev(g["x"]), _dev(g["weight"][rows]), _dev(g["bias"][rows]), _dev(g["offset"]) if use_off else None, fn)
    eta = ops.ratio_hist(n, 1)[0].cpu().numpy()
    x64 = g["x"].astype(np.float64)
    z = x64 @ W.T + b + off
    S_abs = np.abs(x64) @ np.abs(W).T + np.abs(b) + np.abs(off)                     # what an fp32 dot product's error is relative to
    h = 1e-6
    fprime = np.abs(O.norm_fn_apply(fn, z + h) - O.norm_fn_apply(fn, z - h)) / (2 * h)
    with np.errstate(over="ignore", under="ignore", divide="ignore", invalid="ignore"):
        n64 = np.exp(-O.norm_fn_apply(fn, z))
        eta64 = n64[:, 1:] / n64[:, :-1]
    cond = fprime * S_abs
    ok = (n64[:, 1:] > 1e-30) & (n64[:, :-1] > 1e-30) & np.isfinite(eta64)           # away from the fp32 underflow patch (n == 0 -> 2e-23, :167)
    bound = 1e-5 * np.abs(eta64) * (1.0 + cond[:, 1:] + cond[:, :-1])
    err = np.abs(eta - eta64)
    assert ok.mean() > 0.1                                           # norm_fn = exp with the offsets 4..9 underflows most of n in fp32 and fp64 alike
    assert (err[ok] <= bound[ok]).all(), float((err[ok] / bound[ok]).max())
    ref = g["eta_%s_%d" % (fn, use_off)][..., 0]
    fin = np.isfinite(ref) & ok
    worst = float((np.abs(eta[fin] - ref[fin]) / np.abs(ref[fin])).max())
    print("norm_fn %s offset %d: worst relative difference to the reference's fp32 eta %.2e (bound 2e-4 for exp, 2e-5 otherwise), worst ratio to the fp64 conditioning bound %.3f"
          % (fn, use_off, worst, float((err[ok] / bound[ok]).max())))
    assert worst <= (2e-4 if fn == "exp" else 2e-5)               # measured 1.0e-4 (exp: d ln n = e^z dz) / 4.8e-6


def test_linattn_and_softmax_eta_vs_fp64(eig):
    """eta of linear / softmax attention from the SAME fp32 q, k the device formed, evaluated in fp64: nu_t are sums of positive terms, so eta is well
    conditioned and the flat 1e-5 holds (the 3e-4 of the model-level tests is the price of comparing with the reference's own fp32 block forward)."""
    A, Ly, E, S, ops = eig
    g = load_golden("lin_softmax_extractor")
    D, dqk, H = [int(v) for v in g["dims"]]
    B, T, _ = g["x"].shape
    qk = ops.linear(_dev(g["x"]), _dev(g["weight"][:2 * dqk]), _dev(g["bias"][:2 * dqk]), mode="simt")
    qk_h = qk.cpu().numpy().astype(np.float64).reshape(B, T, 2, H, dqk // H)
    q, k = qk_h[:, :, 0], qk_h[:, :, 1]
    nu = ops.linattn_nu(qk, 2 * dqk, B, T, H, dqk // H, dqk)
    eta = ops.ratio_hist(nu, 2)[0].cpu().numpy()
    ref = O.linattn_eta_prefix(q, k, np.float64)[..., 0]
    np.testing.assert_allclose(eta, ref, rtol=1e-5)
    nu_s, m_s = ops.softmax_nu(qk, 2 * dqk, B, T, H, dqk // H, dqk)
    eta_s = ops.softmax_eta(nu_s, m_s)[0].cpu().numpy()
    ref_s = O.softmax_eta_closed(q, k, np.float64)[..., 0]
    fin = np.isfinite(ref_s)
    np.testing.assert_allclose(eta_s[fin], ref_s[fin], rtol=1e-5)


@pytest.mark.parametrize("name", ["model_linattn", "model_normattn", "model_normattn_exp", "model_smattn"])
def test_transformer_extractor_on_device_activations_vs_fp64(eig, name):
    """Model level, split in two: (1) the extractor alone -- eta from the DEVICE's block output x_i against the fp64 oracle on that same x_i -- holds the
    conditioning-aware 1e-5; (2) the block forward is compared separately (activations, test_models_gpu.py).  What remains against the reference's golden
    eta is the reference's own fp32 forward error; its measured worst case is printed."""
    A, Ly, E, S, ops = eig
    sd, cfg, g = golden_model(name)
    model = Ly.TransformerDev(cfg, {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, "cuda")
    x = model.encoder(torch.from_numpy(g["X"]).cuda())
    D, dqk, H = cfg["hidden_dim"], cfg["state_dim"], cfg["num_heads"]
    worst_ref = 0.0
    for i, layer in enumerate(model.layers):
        x = layer(x)
        xh = x.cpu().numpy().astype(np.float64)
        p = {k[len("layers.%d." % i):]: v for k, v in sd.items() if k.startswith("layers.%d." % i)}
        if cfg["attention_fn"] == "norm-attention":
            eta = E.get_eig_att_norm(x, layer, dqk, H, D, cfg)[..., 0]
            rows = O.normattn_rows(D, dqk, H)
            W = p["attention.Wvqkn.weight"][rows].astype(np.float64); b = p["attention.Wvqkn.bias"][rows].astype(np.float64)
            off = p["attention.inner_attn.offset"].astype(np.float64) if cfg.get("offset", False) else 0.0
            z = xh @ W.T + b + off
            S_abs = np.abs(xh) @ np.abs(W).T + np.abs(b) + np.abs(off)
            fp = np.abs(O.norm_fn_apply(cfg["norm_fn"], z + 1e-6) - O.norm_fn_apply(cfg["norm_fn"], z - 1e-6)) / 2e-6
            with np.errstate(over="ignore", under="ignore", divide="ignore", invalid="ignore"):
                n64 = np.exp(-O.norm_fn_apply(cfg["norm_fn"], z)); ref = n64[:, 1:] / n64[:, :-1]
            cond = fp * S_abs
            ok = (n64[:, 1:] > 1e-30) & (n64[:, :-1] > 1e-30) & np.isfinite(ref)
            bound = 1e-5 * np.abs(ref) * (1.0 + cond[:, 1:] + cond[:, :-1])
        else:
            W = p["attention.Wqkv.weight"][:2 * dqk].astype(np.float64); b = p["attention.Wqkv.bias"][:2 * dqk].astype(np.float64)
            qk = (xh @ W.T + b).reshape(xh.shape[0], xh.shape[1], 2, H, dqk // H)
            if cfg["attention_fn"] == "lin-attention":
                eta = E.get_eig_att_linear(x, layer, dqk, H, D)[..., 0]
                ref = O.linattn_eta_prefix(qk[:, :, 0], qk[:, :, 1], np.float64)[..., 0]
            else:
                eta = E.get_eig_att_softmax(x, layer, dqk, H, D)[..., 0]
                ref = O.softmax_eta_closed(qk[:, :, 0], qk[:, :, 1], np.float64)[..., 0]
            ok = np.isfinite(ref)
            # q, k here are fp64 products of the device's fp32 x: the device forms them in fp32 (error eps * sum |x||w| per component); 4e-5 covers it
            bound = 4e-5 * np.abs(ref)
        err = np.abs(eta - ref)
        assert (err[ok] <= bound[ok]).all(), "layer %d: worst ratio %.3g" % (i, float((err[ok] / bound[ok]).max()))
        r = g["eig"][..., i]
        fin = np.isfinite(r) & ok
        worst_ref = max(worst_ref, float((np.abs(eta[fin] - r[fin]) / np.abs(r[fin])).max()))
    print("%s: worst relative difference of eta to the reference's fp32 golden: %.2e (model-level bound 5e-5)" % (name, worst_ref))
    assert worst_ref <= 5e-5


@pytest.mark.parametrize("Hq,Hkv,dtype", [(4, 4, np.float32), (8, 2, np.float32), (6, 1, np.float16)])
def test_notebook_qkv_softmax_extractor_gqa(eig, Hq, Hkv, dtype):
    """notebooks/lm_eigvals.ipynb cell 13 on hooked q / k projections, incl. grouped-query heads (query head h -> key head h // (Hq / Hkv)) and fp16 inputs:
    against the oracle's closed form of the same formula on the widened keys."""
    A, Ly, E, S, ops = eig
    rng = np.random.default_rng(Hq * 10 + Hkv)
    B, T, d = 2, 150, 32
    q = (rng.normal(size=(B, T, Hq, d)) * 0.5).astype(dtype); k = (rng.normal(size=(B, T, Hkv, d)) * 0.5).astype(dtype)
    eta = E.get_eig_from_qkv_att_softmax(torch.from_numpy(q), torch.from_numpy(k), None)
    assert eta.shape == (B, T - 1, Hq, 1) and eta.dtype == np.float64
    kw = np.repeat(k.astype(np.float32), Hq // Hkv, axis=2)
    ref = O.softmax_eta_closed(q.astype(np.float32).astype(np.float64), kw.astype(np.float64), np.float64)
    fin = np.isfinite(ref)
    np.testing.assert_allclose(eta[fin], ref[fin], rtol=1e-5)
    with pytest.raises(Exception):
        E.get_eig_from_qkv_att_softmax(torch.from_numpy(q[:, :, :3]), torch.from_numpy(k[:, :, :2] if Hkv >= 2 else np.concatenate([k, k], 2)), None)


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE C5 at full sequence length and depth (sampled sequences vs the fp64 oracle)
# ----------------------------------------------------------------------------------------------------------------------
def test_c5_mamba_full_depth_vs_oracle(eig):
    """configs[4], Mamba arm: d_model 512, 8 heads, d_state 16, conv 4, GLU, T = 1024, 12 layers.  8 sequences on the device (8192 rows: streamed-operand
    tcgen05 GEMMs, 8-head scan, generic extractor), 2 of them through the fp64 oracle.  Twelve residual blocks of 3xTF32 GEMMs and fp32 scans: the
    accumulated deviation is asserted at 1e-5 of the activation scale and rel 1e-5 (conditioning-aware) on the eigenvalues; measured values printed."""
    A, Ly, E, S, ops = eig
    cfg = dict(layer="mamba", version="mamba2", num_layers=12, num_heads=8, input_dim=1, output_dim=64, hidden_dim=512, state_dim=16, conv_dim=4,
               expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none", token_embedding=True, vocab_size=1000)
    sd = Ly.init_mamba_state_dict(cfg, 7)
    model = Ly.MambaDev(cfg, sd, "cuda")
    X = torch.randint(0, 1000, (8, 1024), generator=torch.Generator().manual_seed(4))
    res = A.mamba_pass(model, X.cuda())
    sel = [0, 7]
    ocfg = dict(num_layers=12, d_inner=512, ngroups=1, d_state=16, nheads=8, headdim=64, prenorm=True)
    ref, xr = O.mamba_eval_pass(X[sel].numpy(), {k: v.numpy() for k, v in sd.items()}, ocfg, np.float64)
    xd = res.x_last.cpu().numpy()[sel]
    dev = np.abs(xd - xr).max() / np.abs(xr).max()
    e = res.eig_host()[sel]
    with np.errstate(divide="ignore"):
        ratio = np.abs(e - ref) / (np.abs(ref) * (1 + np.abs(np.log(np.maximum(ref, 1e-300)))) + 1e-44)
    print("C5 mamba T=1024 x 12 layers: activation deviation %.2e of scale, eigenvalue worst conditioned rel. error %.2e (per layer: %s)"
          % (dev, ratio.max(), " ".join("%.1e" % ratio[..., l].max() for l in range(12))))
    assert dev <= 1e-5                                             # measured 6.6e-7
    assert_eig_close(e, ref, rtol=1e-5)                            # measured 7.0e-7: the flat north-star bound holds through all 12 blocks
    assert (res.counts.cpu().numpy()[..., 7] == 1024).all()


def test_c5_norm_attention_full_depth_vs_oracle(eig):
    """configs[4], normalised-attention arm: d_model = d_qk = 512, 8 heads, conv 4, softplus gate, elu feature map, GLU mixer, T = 1024, 12 layers;
    4 sequences on the device, 1 through the fp64 oracle."""
    A, Ly, E, S, ops = eig
    cfg = dict(layer="transformer", input_dim=1, output_dim=64, num_layers=12, hidden_dim=512, embedding=True, vocab_size=1000, max_pos_embed=1024,
               pooling="none", dual=False, classifier=False, mixer_dim=1024, norm="layer", dropout=0.0, state_dim=512, num_heads=8, att_dropout=0.0,
               use_flash=False, attention_fn="norm-attention", mixer="glu", mode="attention", norm_fn="softplus", approx_fn="elu", scale_B=False,
               offset=True, offset_init="exp", learn_A=False, dim_conv=4)
    sd = Ly.init_transformer_state_dict(cfg, 7)
    model = Ly.TransformerDev(cfg, sd, "cuda")
    X = torch.randint(0, 1000, (4, 1024), generator=torch.Generator().manual_seed(4))
    res = A.transformer_pass(model, X.cuda(), cfg)
    sel = [3]
    ocfg = dict(cfg, d_model=512, d_qk=512)
    ref, xr = O.transformer_eval_pass(X[sel].numpy(), {k: v.numpy() for k, v in sd.items()}, ocfg, np.float64, eta_dtype=np.float64)
    xd = res.x_last.cpu().numpy()[sel]
    dev = np.abs(xd - xr).max() / np.abs(xr).max()
    e = res.eig_host()[sel]
    ref32, _ = O.transformer_eval_pass(X[sel].numpy(), {k: v.numpy() for k, v in sd.items()}, ocfg, np.float64)   # eta formed with the reference's fp32 n (underflow patch)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        fin = np.isfinite(ref) & np.isfinite(ref32) & (np.abs(ref32 / ref - 1.0) < 1e-3)   # away from fp32 underflow of n (n == 0 -> 2e-23, eval_eig.py:167)
    assert fin.mean() > 0.5
    rel = np.abs(e[fin] - ref[fin]) / np.abs(ref[fin])
    print("C5 norm-attention T=1024 x 12 layers: activation deviation %.2e of scale; eta relative error: median %.2e, 99 %% %.2e, worst %.2e"
          % (dev, np.median(rel), np.quantile(rel, 0.99), rel.max()))
    assert dev <= 2.5e-5                                           # measured after 12 blocks: 6.3e-6 (chunked tensor-core attention), 4.7e-6 (EIGB200_LINATTN_FORM=col)
    # eta = n_{t+1} / n_t with n = exp(-softplus(z + offset)), offsets 4..9: d ln eta = dz_{t+1} - dz_t and dz = W_n . dx, so the activation deviation
    # above (x |W_n| sqrt(D)) is amplified into a ~1e-3 tail -- conditioning of the quantity, not of the kernel.  The kernel itself: eta from the DEVICE's
    # last-layer activations against the fp64 formula on the same activations holds the conditioning-aware 1e-5 (below).
    assert np.median(rel) <= 1e-5 and np.quantile(rel, 0.99) <= 1e-3 and rel.max() <= 2e-2
    x = res.x_last
    layer = model.layers[-1]
    eta = E.get_eig_att_norm(x, layer, 512, 8, 512, cfg)[..., 0]
    xh = x.cpu().numpy().astype(np.float64)
    rows = O.normattn_rows(512, 512, 8)
    W = sd["layers.11.attention.Wvqkn.weight"].numpy()[rows].astype(np.float64); b = sd["layers.11.attention.Wvqkn.bias"].numpy()[rows].astype(np.float64)
    off = sd["layers.11.attention.inner_attn.offset"].numpy().astype(np.float64)
    z = xh @ W.T + b + off
    S_abs = np.abs(xh) @ np.abs(W).T + np.abs(b) + np.abs(off)
    fp = np.abs(O.norm_fn_apply("softplus", z + 1e-6) - O.norm_fn_apply("softplus", z - 1e-6)) / 2e-6
    n64 = np.exp(-O.norm_fn_apply("softplus", z)); r64 = n64[:, 1:] / n64[:, :-1]
    cond = fp * S_abs
    okm = (n64[:, 1:] > 1e-30) & (n64[:, :-1] > 1e-30)
    bound = 1e-5 * np.abs(r64) * (1.0 + cond[:, 1:] + cond[:, :-1])
    errm = np.abs(eta - r64)
    assert (errm[okm] <= bound[okm]).all(), float((errm[okm] / bound[okm]).max())


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE C3, S5 arm at the layer shape
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("disc", ["zoh", "bilinear"])
def test_c3_shaped_s5_layer_vs_oracle(eig, disc):
    """configs[2], S5: T = 2048, d_model 128, P = 256 states (conj_sym: 2 Re(C h)), both discretisations."""
    A, Ly, E, S, ops = eig
    rng = np.random.default_rng(5)
    P, H, B, T = 256, 128, 2, 2048
    prm = dict(Lambda_re=(-np.abs(rng.normal(0.5, 0.2, P))).astype(np.float32), Lambda_im=rng.normal(0, 6, P).astype(np.float32),
               B=(rng.normal(size=(P, H, 2)) / np.sqrt(H)).astype(np.float32), C=(rng.normal(size=(H, P, 2)) / np.sqrt(P)).astype(np.float32),
               D=rng.normal(size=H).astype(np.float32), log_step=rng.uniform(np.log(1e-3), np.log(1e-1), (P, 1)).astype(np.float32))
    u = rng.normal(size=(B, T, H)).astype(np.float32)
    y, h = S.s5_forward(prm, u, discretization=disc, conj_sym=True, return_states=True)
    yr, hr, _ = O.s5_forward(prm, u, discretization=disc, conj_sym=True)
    scale = np.abs(hr).max(axis=1, keepdims=True)
    dh = (np.abs(h.cpu().numpy() - hr) / scale).max()
    dy = np.abs(y.cpu().numpy() - yr).max() / np.abs(yr).max()
    print("C3 S5 %s P=256 T=2048: states %.2e of scale, output %.2e" % (disc, dh, dy))
    assert dh <= 2e-4 and dy <= 2e-5                                 # measured 7e-5 / 3e-6; states: fp32 discretisation ((lam_bar - 1)/Lambda cancels), |lam| ~ 0.9999 over 2048 steps
    lam = ops.ssm_lambda("s5_zoh" if disc == "zoh" else "s5_bilinear", _dev(prm["Lambda_re"]), _dev(prm["Lambda_im"]), _dev(prm["log_step"])).cpu().numpy()
    if disc == "zoh":
        np.testing.assert_allclose(lam, O.s5_lambda(prm["Lambda_re"], prm["Lambda_im"], prm["log_step"]), rtol=1e-5, atol=1e-7)


# ----------------------------------------------------------------------------------------------------------------------
# eval_eig on 2 GPUs == eval_eig on 1 GPU
# ----------------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _eval_eig_worker(rank, world, port, ckpt, out_dir, cfg_json):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import eigb200.analysis as A
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world)
    cfg = json.loads(cfg_json)
    wd = os.path.join(out_dir, "w%d_r%d" % (world, rank))
    os.makedirs(wd, exist_ok=True)
    os.chdir(wd)
    g = torch.Generator().manual_seed(9)
    X = torch.randint(0, cfg["vocab_size"], (32, 80), generator=g)
    loader = [(X, torch.zeros(32), None)]
    args = dict(model=dict(cfg), train=dict(lr=1e-3), dataset=dict(name="mqar"), seed=1919)
    out = A.eval_eig(args, dict(batch_size=32, save_path=wd + "/"), None, args["dataset"], loader, ckpt, 0.5)
    if rank == 0:
        np.savez(os.path.join(out_dir, "res_w%d.npz" % world), eig=out[0], eig_init=out[1], pct=out[2], pct_init=out[3], ph=out[4], ph_init=out[5],
                 report=np.frombuffer(open("percentage_file.txt", "rb").read(), dtype=np.uint8))
    if world > 1:
        dist.destroy_process_group()


def test_eval_eig_two_gpus_equals_one_gpu(eig, tmp_path):
    """analysis.eval_eig with world_size 2 (NCCL, batch sharded 16 + 16: 1280 rows per GPU, the tensor-core paths on both sides, the integer all-reduce of the statistics, the gathered eigenvalue array) against the
    1-GPU run on the same checkpoint and batch: bin counts / percentages / report file bit for bit, eigenvalues bit for bit as well (every sequence is
    computed by the same kernels whichever GPU owns it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    A, Ly, E, S, ops = eig
    cfg = dict(layer="mamba", version="mamba2", num_layers=2, num_heads=1, input_dim=1, output_dim=64, hidden_dim=128, state_dim=16, conv_dim=4,
               expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none", token_embedding=True, vocab_size=64)
    sd = Ly.init_mamba_state_dict(cfg, 3)
    ckpt = str(tmp_path / "model.pth")
    torch.save(sd, ckpt)
    for world in (1, 2):
        mp.spawn(_eval_eig_worker, args=(world, _free_port(), ckpt, str(tmp_path), json.dumps(cfg)), nprocs=world, join=True)
    a = np.load(str(tmp_path / "res_w1.npz")); b = np.load(str(tmp_path / "res_w2.npz"))
    for k in ("pct", "pct_init", "ph", "ph_init", "report"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert a["eig"].shape == b["eig"].shape == (32, 80, 1, 2)
    np.testing.assert_array_equal(a["eig"], b["eig"])
    np.testing.assert_array_equal(a["eig_init"], b["eig_init"])
