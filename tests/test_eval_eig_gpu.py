"""GPU end-to-end: eval_eig(args, conf_args, wandb_config, data_config, loader, path_file, perf) with a fake loader and a synthetic
checkpoint returns the reference's 6-tuple (shapes, dtypes, values vs the oracle) and writes the same files."""
import copy
import os

import numpy as np
import pytest
import torch
import yaml

import oracle as O
from conftest import golden_model, assert_eig_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import eigb200.analysis as A
    return A


def _mamba_ocfg(cfg):
    D = cfg["hidden_dim"]; hd = D // cfg["num_heads"]
    return dict(num_layers=cfg["num_layers"], d_inner=D, ngroups=1, d_state=cfg["state_dim"], nheads=D // hd, headdim=hd, prenorm=cfg["prenorm"])


def test_eval_eig_mamba_end_to_end(A, tmp_path, monkeypatch):
    import eigb200.layers as Ly
    sd, cfg, g = golden_model("model_mamba2")
    ckpt = str(tmp_path / "model-perf0.900.pth")
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, ckpt)
    X = torch.from_numpy(g["X"])
    loader = [(X, torch.zeros(X.shape[0]), None)]
    model_cfg = dict(cfg, layer="mamba", seq_len=X.shape[1])
    args = {"seed": 1919, "model": model_cfg, "train": {"lr": 0.01}, "dataset": {"name": "MQAR"}}
    save_dir = str(tmp_path) + "/"
    conf = {"batch_size": X.shape[0], "save_path": save_dir}
    monkeypatch.chdir(tmp_path)
    out = A.eval_eig(args, conf, None, args["dataset"], loader, ckpt, 0.9)
    eig, eig_init, pct, pct_init, pct_ph, pct_ph_init = out
    assert "layer" not in args["model"]                                           # popped, like the reference (eval_eig.py:479)
    B, T, H, L = X.shape[0], X.shape[1], cfg["num_heads"], cfg["num_layers"]
    assert eig.shape == (B, T, H, L) and eig.dtype == np.float32 and eig_init.shape == eig.shape
    assert pct.shape == (7, B, H, L) and pct_ph.shape == (6, B, H, L) and pct.dtype == np.float64
    assert_eig_close(eig, g["eig"], rtol=3e-5, what="trained-pass eigenvalues vs the reference's")
    # init pass == the reference's model at construction under the same seed (goldens scaled the dt rows afterwards)
    init_sd = {k: v.numpy() for k, v in Ly.init_mamba_state_dict(cfg, 1919).items()}
    ref_init, _ = O.mamba_eval_pass(g["X"], init_sd, _mamba_ocfg(cfg), np.float64)
    assert_eig_close(eig_init, ref_init, rtol=3e-5, what="init-pass eigenvalues vs the oracle")
    rad = np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2))
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(pct, O.threshold_analysis(rad, O.THRESHOLDS_RADIUS))
        np.testing.assert_array_equal(pct_ph, O.threshold_analysis(np.arctan2(eig.imag, eig.real) * 180 / np.pi, O.THRESHOLDS_PHASE))
    # files: the reference's names under save_path + name (eval_eig.py:807-851) and ./percentage_file.txt
    name = "MQARdmodel{0}-seed{3}-num_layers{4}-dqk{1}-conv_dim{5}-lr{2}".format(cfg["hidden_dim"], cfg["state_dim"], 0.01, 1919, L, 0) + "-perf0.900"
    d = os.path.join(save_dir, name)
    for f in ["eig", "eig_init", "percentage", "percentage_init", "percentage_phase", "percentage_phase_init", "percentage_mean",
              "percentage_init_mean", "percentage_std", "percentage_init_std"]:
        assert os.path.exists(os.path.join(d, f + ".npy")), f
    np.testing.assert_array_equal(np.load(os.path.join(d, "eig.npy")), eig)
    np.testing.assert_array_equal(np.load(os.path.join(d, "percentage_mean.npy")), np.mean(pct, axis=1))
    np.testing.assert_array_equal(np.load(os.path.join(d, "percentage_std.npy")), np.std(pct, axis=1))
    used = yaml.safe_load(open(os.path.join(d, "used_config.yaml")))
    assert used["seed"] == 1919 and "layer" not in used["model"]
    assert open(tmp_path / "percentage_file.txt").read().startswith("threshold radius: [  0.1   0.5   0.9   1.   10.  100. ]")


@pytest.mark.parametrize("name", ["model_linattn", "model_normattn", "model_smattn"])
def test_eval_eig_transformer_end_to_end(A, tmp_path, monkeypatch, name):
    sd, cfg, g = golden_model(name)
    ckpt = str(tmp_path / "tf.pth")
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, ckpt)
    X = torch.from_numpy(g["X"])
    loader = [(X, torch.zeros(X.shape[0]), None)]
    args = {"seed": 1919, "model": dict(cfg, layer="transformer", seq_len=X.shape[1]), "train": {"lr": 0.001}, "dataset": {"name": "MQAR"}}
    monkeypatch.chdir(tmp_path)
    eig, eig_init, pct, pct_init, pct_ph, pct_ph_init = A.eval_eig(args, {"batch_size": X.shape[0], "save_path": str(tmp_path) + "/"}, None,
                                                                   args["dataset"], loader, ckpt, 0.5)
    assert eig.shape == g["eig"].shape and eig.dtype == np.float64
    fin = np.isfinite(g["eig"])
    np.testing.assert_allclose(eig[fin], g["eig"][fin], rtol=5e-5)
    # the reference's golden model IS the seed-1919 construction, so the init pass must reproduce it as well
    np.testing.assert_allclose(eig_init[fin], g["eig"][fin], rtol=5e-5)
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(pct, O.threshold_analysis(eig, O.THRESHOLDS_RADIUS))
    np.testing.assert_array_equal(pct_ph, g["percentage_phase"])


@pytest.mark.parametrize("kind", ["lru", "s5", "s4"])
def test_eval_eig_ssm_branch(A, tmp_path, monkeypatch, kind):
    import eigb200.ssm as S
    P, H, L = (16, 8, 2)
    cfg = dict(layer=kind, state_dim=P, hidden_dim=H, num_layers=L, seq_len=64, num_blocks=2, r_min=0.9, r_max=0.99)
    trained = S.get_init_layers_ssm(7, {}, {}, dict(cfg), 64, kind, 4)
    flat = {}
    for i, lay in enumerate(trained):
        for k, v in lay.items():
            flat["model/params/encoder/layers_%d/seq/%s" % (i, k)] = v
    ckpt = str(tmp_path / "ssm.npz")
    np.savez(ckpt, **flat)
    args = {"seed": 3, "model": cfg, "train": {"lr": 0.001}, "dataset": {"name": "ListOps"}}
    monkeypatch.chdir(tmp_path)
    eig, eig_init, pct, pct_init, pct_ph, pct_ph_init = A.eval_eig(args, {"batch_size": 4, "save_path": str(tmp_path) + "/"}, None,
                                                                   args["dataset"], [], ckpt, 0.1)
    n = P // 2 if kind == "s5" else P
    assert eig.shape == (n, L) and eig.dtype == np.complex64 and pct.shape == (7, L) and pct_ph.shape == (6, L)
    if kind == "lru":
        ref = np.stack([O.lru_lambda(l["nu_log"], l["theta_log"]) for l in trained], axis=1)
        np.testing.assert_allclose(eig, ref, rtol=1e-5, atol=1e-7)
    elif kind == "s5":
        ref = np.stack([O.s5_lambda(l["Lambda_re"], l["Lambda_im"], l["log_step"]) for l in trained], axis=1)
        np.testing.assert_allclose(eig, ref, rtol=1e-5, atol=1e-7)
    rad, ph = O.radius_phase(eig)
    np.testing.assert_array_equal(pct, O.threshold_analysis_ssm(rad, O.THRESHOLDS_RADIUS))
    np.testing.assert_array_equal(pct_ph, O.threshold_analysis_ssm(ph, O.THRESHOLDS_PHASE))
    assert open(tmp_path / "percentage_file.txt").read().startswith("threshold radius:")


def test_mamba_pass_graph_replay_matches_eager(A):
    """The CUDA-graph replay of a pass gives bit-identical eigenvalues and counts to the eager launches, also for a new batch."""
    import eigb200.layers as Ly
    sd, cfg, g = golden_model("model_mamba2")
    model = Ly.MambaDev(cfg, {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, "cuda")
    X = torch.from_numpy(g["X"]).cuda()
    gp = A.MambaPassGraph(model, X)
    for Xb in (X, torch.flip(X, dims=[0]), X.roll(3, dims=1)):
        ref = A.mamba_pass(model, Xb)
        res = gp.run(Xb)
        torch.cuda.synchronize()
        assert torch.equal(res.eig, ref.eig) and torch.equal(res.counts, ref.counts)
    assert gp.launches_per_run > 0


def test_full_size_pass_properties_c2(A):
    """BASELINE config C2 model (T=512, d_model=128, d_state=16, 4 layers) at 1024 sequences: properties of the WHOLE pass that need no CPU
    reference -- sequences are independent (permutation equivariance, which is what makes batch sharding across GPUs exact), bin counts are a
    checksum of the eigenvalue array, the graph replay equals the eager pass, and a sample of sequences matches the fp64 oracle."""
    import eigb200.layers as Ly
    import eigb200.ops as ops
    import eigb200.extractors as E
    cfg = dict(layer="mamba", version="mamba2", num_layers=4, num_heads=1, input_dim=1, output_dim=8192, hidden_dim=128, state_dim=16,
               conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none",
               token_embedding=True, vocab_size=8192)
    sd = Ly.init_mamba_state_dict(cfg, 1919)
    model = Ly.MambaDev(cfg, sd, "cuda")
    B, T, L = 1024, 512, 4
    g = torch.Generator().manual_seed(5)
    X = torch.randint(0, 8192, (B, T), generator=g).cuda()
    res = A.mamba_pass(model, X)
    eig, counts = res.eig.clone(), res.counts.clone()
    assert eig.shape == (B, T, 1, L) and counts.shape == (L, B, 1, ops.NSLOT)
    assert bool(((eig > 0) & (eig < 1)).all())                                        # lambda = exp(dt A), dt > 0, A < 0
    # (1) permutation equivariance, bitwise
    perm = torch.randperm(B, generator=g).cuda()
    res_p = A.mamba_pass(model, X[perm])
    assert torch.equal(res_p.eig, eig[perm]) and torch.equal(res_p.counts, counts[:, perm])
    # (2) the counts are the histogram of the eigenvalues the pass wrote
    assert bool((counts[..., 7] == T).all()) and bool((counts[..., :7].sum(-1) >= T).all())
    rad = torch.sqrt(eig * eig)                                                       # (B,T,1,L) float32 radius as eval_eig.py:605-606
    chk = E.threshold_counts_device(rad.reshape(B, T, L), O.THRESHOLDS_RADIUS)        # (B, L, 8)
    assert torch.equal(chk[..., :7].permute(1, 0, 2), counts[:, :, 0, :7])
    # (3) CUDA-graph replay == eager
    gp = A.MambaPassGraph(model, X)
    r2 = gp.run(X)
    torch.cuda.synchronize()
    assert torch.equal(r2.eig, eig) and torch.equal(r2.counts, counts)
    # (4) a sample of sequences against the fp64 oracle of the whole pass
    idx = [0, 511, 1023]
    ref, _ = O.mamba_eval_pass(X[idx].cpu().numpy(), {k: v.numpy() for k, v in sd.items()}, _mamba_ocfg(cfg), np.float64)
    assert_eig_close(eig[idx].cpu().numpy(), ref, rtol=3e-5)


def test_eval_eig_optional_quantiles(A, tmp_path, monkeypatch):
    """`quantiles:` in the analysis YAML (beyond the reference's keys): on-device log-spaced histogram of the radii per (head, layer) and its quantiles, saved next to
    the reference's 10 files; without the key nothing extra is written and the 6-tuple is unchanged."""
    sd, cfg, g = golden_model("model_mamba2")
    ckpt = str(tmp_path / "model.pth")
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, ckpt)
    X = torch.from_numpy(g["X"])
    loader = [(X, torch.zeros(X.shape[0]), None)]
    monkeypatch.chdir(tmp_path)
    qs = [0.1, 0.5, 0.9]
    outs = {}
    for tag, conf in (("plain", {"batch_size": X.shape[0], "save_path": str(tmp_path) + "/plain_"}),
                      ("q", {"batch_size": X.shape[0], "save_path": str(tmp_path) + "/q_", "quantiles": qs})):
        args = {"seed": 1919, "model": dict(cfg, layer="mamba", seq_len=X.shape[1]), "train": {"lr": 0.01}, "dataset": {"name": "MQAR"}}
        outs[tag] = A.eval_eig(args, conf, None, args["dataset"], loader, ckpt, 0.9)
    for a, b in zip(outs["plain"], outs["q"]):
        np.testing.assert_array_equal(a, b)
    dq = [d for d in os.listdir(tmp_path) if d.startswith("q_")][0]
    dp = [d for d in os.listdir(tmp_path) if d.startswith("plain_")][0]
    assert not os.path.exists(tmp_path / dp / "radius_quantiles.npy")
    quant = np.load(tmp_path / dq / "radius_quantiles.npy"); hist = np.load(tmp_path / dq / "radius_loghist.npy")
    eig = outs["q"][0]
    H, L = eig.shape[2], eig.shape[3]
    assert quant.shape == (3, H, L) and hist.shape == (H * L, 515) and hist.sum() == eig.size
    width = (1e2 / 1e-8) ** (1 / 512)
    n = eig.shape[0] * eig.shape[1]
    for h in range(H):
        for l in range(L):
            srt = np.sort(np.abs(eig[:, :, h, l].astype(np.float64)).ravel())
            for qi, qq in enumerate(qs):                           # between neighbouring order statistics, +- one bin
                k = int(np.ceil(qq * n))
                assert srt[max(k - 2, 0)] / width / 1.001 <= quant[qi, h, l] <= srt[min(k, n - 1)] * width * 1.001, (h, l, qq, quant[qi, h, l])
