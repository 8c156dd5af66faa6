"""The oracle against the vectors produced by the reference's own code (tests/golden/make_golden.py).
CPU only.  This is what pins the oracle; the GPU tests then compare the CUDA path with the oracle."""
import numpy as np
import pytest

import oracle as O
from conftest import load_golden, golden_model, assert_eig_close


def test_threshold_analysis_matches_reference():
    g = load_golden("thresholds")
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(O.threshold_analysis(g["v32"], O.THRESHOLDS_RADIUS), g["p_v32"])
        np.testing.assert_array_equal(O.threshold_analysis(g["v64"], O.THRESHOLDS_RADIUS), g["p_v64"])
        np.testing.assert_array_equal(O.threshold_analysis(g["ph"], O.THRESHOLDS_PHASE), g["p_ph"])
        np.testing.assert_array_equal(O.threshold_analysis_ssm(g["s"], O.THRESHOLDS_RADIUS), g["p_s"])
        np.testing.assert_array_equal(O.threshold_analysis_ssm(g["sp"], O.THRESHOLDS_PHASE), g["p_sp"])
    half = O.threshold_analysis(np.full((2, 8, 1, 1), 0.5), O.THRESHOLDS_RADIUS)
    np.testing.assert_array_equal(half, g["p_half"])
    assert half.sum(axis=0).ravel().tolist() == [200.0, 200.0]        # closed bins double count an edge value


def test_threshold_float32_compare_mode_differs_only_at_edges():
    g = load_golden("thresholds")
    with np.errstate(invalid="ignore"):
        c64 = O.threshold_counts(g["v32"], O.THRESHOLDS_RADIUS, axis=1, compare="float64")
        c32 = O.threshold_counts(g["v32"], O.THRESHOLDS_RADIUS, axis=1, compare="float32")
    # planted fp32(0.1): equal to the fp32 threshold but above the fp64 one
    assert (c64 != c32).sum() > 0
    assert np.abs(c64 - c32).max() <= 2


def test_mamba2_extractor():
    g = load_golden("mamba2_extractor")
    D, G, N, H = g["dims"]
    lam32 = O.mamba2_eig(g["x"], g["in_proj_weight"], g["dt_bias"], g["A_log"], D, G, N, H, np.float32)
    lam64 = O.mamba2_eig(g["x"], g["in_proj_weight"], g["dt_bias"], g["A_log"], D, G, N, H, np.float64)
    assert lam32.shape == g["lam"].shape and lam32.dtype == np.float32 and g["lam"].dtype == np.float32
    assert_eig_close(lam32, g["lam"])
    assert_eig_close(g["lam"], lam64)
    g2 = load_golden("mamba2_lti_extractor")
    B, T = g2["shape"]
    np.testing.assert_allclose(O.mamba2_lti_eig(B, T, g2["A"], g2["beta"], np.float32), g2["lam"], rtol=1e-6)


@pytest.mark.parametrize("fn", ["exp", "elu", "softplus", "sigmoid"])
@pytest.mark.parametrize("use_off", [0, 1])
def test_norm_extractor(fn, use_off):
    g = load_golden("norm_extractor")
    D, dqk, H = g["dims"]
    eta = O.normattn_eta(g["x"], g["weight"], g["bias"], g["offset"] if use_off else None, fn, D, dqk, H, np.float32)
    ref = g["eta_%s_%d" % (fn, use_off)]
    assert eta.shape == ref.shape and eta.dtype == np.float64 and ref.dtype == np.float64
    # n is fp32: a 1-ulp difference in n is 6e-8 relative on eta, amplified by exp(-exp(.)) sensitivity
    fin = np.isfinite(ref)
    assert (np.isfinite(eta) == fin).all()
    np.testing.assert_allclose(eta[fin], ref[fin], rtol=5e-4)
    if fn == "exp":
        assert (ref == 1.0).any() or (np.abs(np.log10(ref[fin])) > 15).any()      # the 2e-23 patch is exercised


def test_lin_softmax_extractor():
    g = load_golden("lin_softmax_extractor")
    D, dqk, H = g["dims"]
    q, k = O.linattn_qk(g["x"], g["weight"], g["bias"], dqk, H, np.float32)
    np.testing.assert_allclose(O.linattn_eta_quadratic(q, k), g["eta_lin"], rtol=1e-5)
    np.testing.assert_allclose(O.linattn_eta_prefix(q, k), g["eta_lin"], rtol=1e-5)
    np.testing.assert_allclose(O.softmax_eta_quadratic(q, k), g["eta_sm"], rtol=1e-4)
    np.testing.assert_allclose(O.softmax_eta_closed(q, k), g["eta_sm"], rtol=1e-4)


def test_lru_s5_eigs():
    g = load_golden("lru_s5_eigs")
    for i in range(3):
        lam = O.lru_lambda(g["lru_nu_%d" % i], g["lru_theta_%d" % i], np.complex64)
        np.testing.assert_allclose(lam[:, None], g["lru_eig_%d" % i], rtol=1e-5)
        np.testing.assert_allclose(np.abs(lam), np.exp(-np.exp(g["lru_nu_%d" % i].astype(np.float64))), rtol=1e-5)
        lam = O.s5_lambda(g["s5_Lambda_re_%d" % i], g["s5_Lambda_im_%d" % i], g["s5_log_step_%d" % i], np.complex64)
        np.testing.assert_allclose(lam[:, None], g["s5_eig_%d" % i], rtol=1e-5)


def test_hippo_and_dplr():
    g = load_golden("hippo_s4")
    for N in (8, 16, 64):
        np.testing.assert_allclose(O.make_hippo(N), g["hippo_%d" % N], rtol=1e-12)
        Lam, P, B, V, Bo = O.make_dplr_hippo(N)
        np.testing.assert_allclose(Lam.real, g["Lambda_%d" % N].real, rtol=1e-9)
        # eigenvector phases are LAPACK-run dependent; the invariants are |P| and |B| per eigenvalue
        np.testing.assert_allclose(np.sort(Lam.imag), np.sort(g["Lambda_%d" % N].imag), atol=1e-8)
    for N in (8, 16, 64):
        for l in (0, 1):
            layer = {k: g["s4_N%d_l%d_%s" % (N, l, k)] for k in ["Lambda_re", "Lambda_im", "P", "B", "C", "log_step"]}
            Ab, ev = O.s4_eigvals(layer, 1, np.complex64)
            ref = g["s4_N%d_l%d_Abar" % (N, l)]
            assert np.abs(Ab - ref).max() <= 2e-5 * np.abs(ref).max()
            Ab128, ev128 = O.s4_eigvals(layer, 1, np.complex128)
            assert np.abs(Ab128 - ref).max() <= 2e-5 * np.abs(ref).max()
    # exact spectrum at the HiPPO init (fp64 parameters; fp32-rounded parameters already move it by O(1e-2))
    for N in (8, 16):
        Lam, P, _, _, _ = O.make_dplr_hippo(N)
        Ab = O.discrete_dplr_abar(Lam, P, P, 0.01)
        ev = np.linalg.eigvals(Ab)
        np.testing.assert_allclose(np.sort(ev.real), np.sort(O.dplr_exact_spectrum(N, 0.01)), rtol=1e-4)
        assert np.abs(ev.imag).max() < 1e-4
    np.testing.assert_allclose(O.s5_discretize(g["disc_Lambda"], g["disc_B"], g["disc_Delta"], "zoh")[0], g["zoh_L"], rtol=1e-5)
    np.testing.assert_allclose(O.s5_discretize(g["disc_Lambda"], g["disc_B"], g["disc_Delta"], "zoh")[1], g["zoh_B"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(O.s5_discretize(g["disc_Lambda"], g["disc_B"], g["disc_Delta"], "bilinear")[0], g["bil_L"], rtol=1e-5)
    np.testing.assert_allclose(O.s5_discretize(g["disc_Lambda"], g["disc_B"], g["disc_Delta"], "bilinear")[1], g["bil_B"], rtol=1e-4, atol=1e-6)


def test_ssd_restatement_vs_third_party():
    g = load_golden("ssd_small")
    y = O.ssd_scan_sequential(g["x"], g["dt"], g["A"], g["Bm"], g["Cm"], g["D"])
    scale = np.abs(g["y_fla"]).max()
    assert np.abs(y - g["y_fla"]).max() <= 1e-5 * scale
    ych = O.ssd_scan_chunked(g["x"], g["dt"], g["A"], g["Bm"], g["Cm"], g["D"], chunk=16).numpy()
    assert np.abs(y - ych).max() <= 2e-5 * scale


def _mamba_cfg(cfg):
    D = cfg["hidden_dim"]
    hd = D // cfg["num_heads"]
    di = cfg["expansion"] * D
    return dict(num_layers=cfg["num_layers"], d_inner=di, ngroups=1, d_state=cfg["state_dim"], nheads=di // hd, headdim=hd,
                prenorm=cfg["prenorm"])


def test_mamba_model_pass():
    sd, cfg, g = golden_model("model_mamba2")
    eig, x = O.mamba_eval_pass(g["X"], sd, _mamba_cfg(cfg), np.float64)
    assert eig.shape == g["eig"].shape
    np.testing.assert_allclose(x, g["act_3"], rtol=0, atol=2e-5 * np.abs(g["act_3"]).max())
    assert_eig_close(g["eig"], eig, rtol=2e-5)
    with np.errstate(invalid="ignore"):
        p = O.threshold_analysis(eig.astype(np.float32), O.THRESHOLDS_RADIUS)
    assert np.abs(p - g["percentage"]).max() <= 100.0 / eig.shape[1] + 1e-9     # at most one value flips an edge
    assert len(np.unique(np.argmax(g["percentage"], axis=0))) > 1               # fixture spreads over several bins


def _tf_cfg(cfg):
    c = dict(cfg)
    c.update(d_model=cfg["hidden_dim"], d_qk=cfg["state_dim"])
    return c


@pytest.mark.parametrize("name", ["model_linattn", "model_linattn_glu_conv", "model_normattn", "model_normattn_exp", "model_smattn", "model_linattn_hybrid"])
def test_transformer_model_pass(name):
    sd, cfg, g = golden_model(name)
    eig, x = O.transformer_eval_pass(g["X"], sd, _tf_cfg(cfg), np.float64)
    np.testing.assert_allclose(x, g["act_2"], rtol=0, atol=3e-5 * np.abs(g["act_2"]).max())
    fin = np.isfinite(g["eig"])
    np.testing.assert_allclose(eig[fin], g["eig"][fin], rtol=2e-4)
    with np.errstate(invalid="ignore"):
        ph = O.threshold_analysis(0 * eig, O.THRESHOLDS_PHASE)
    np.testing.assert_array_equal(ph, g["percentage_phase"])


@pytest.mark.parametrize("N,L,step", [(8, 32, 0.02), (16, 64, 0.01), (16, 50, 0.08)])
def test_s4_kernel_identity_pins_kernel_dplr(N, L, step):
    """The S4 kernel identity: kernel_DPLR (generating function at the roots of unity, models/s4.py:50-69) equals the impulse response
    Re(Cbar Abar^t Bbar) of discrete_DPLR (:16-40) -- an independent route through the reference's own algebra, since JAX is not installed."""
    Lam, P, B, _, _ = O.make_dplr_hippo(N)
    rng = np.random.default_rng(N + L)
    C = (rng.normal(size=N) + 1j * rng.normal(size=N)) * 0.5 ** 0.5
    k1 = O.s4_kernel_dplr(Lam, P, P, B, C, step, L)
    k2 = O.s4_kernel_recurrent(Lam, P, P, B, C, step, L)
    np.testing.assert_allclose(k1, k2, rtol=0, atol=1e-12 * max(1.0, np.abs(k2).max()))


def _c1_fixture():
    import hashlib
    import json
    import torch
    import eigb200.layers as Ly
    g = load_golden("c1_linattn_mqar")
    cfg = json.loads(bytes(g["cfg_json"]).decode())
    hashes = json.loads(bytes(g["sd_sha256_json"]).decode())
    sd = Ly.init_transformer_state_dict(cfg, 1919)
    got = {k: hashlib.sha256(np.ascontiguousarray(v.numpy()).tobytes()).hexdigest() for k, v in sd.items()}
    return g, cfg, sd, hashes, got


def test_c1_model_init_is_the_reference_construction():
    """BASELINE configs[0] (linear attention, MQAR seq 64, d_model 64, 2 layers): every parameter drawn by eigb200.layers.init_transformer_state_dict
    under seed 1919 is bit-identical (SHA-256) to the reference's Transformer(cfg) -- so the 4 MB of weights need not be stored with the fixture."""
    g, cfg, sd, hashes, got = _c1_fixture()
    assert set(got) == set(hashes)
    assert all(got[k] == hashes[k] for k in hashes), [k for k in hashes if got[k] != hashes[k]]


def test_c1_oracle_pass_matches_reference():
    g, cfg, sd, hashes, got = _c1_fixture()
    eig, x = O.transformer_eval_pass(g["X"], {k: v.numpy() for k, v in sd.items()}, _tf_cfg(cfg), np.float64)
    np.testing.assert_allclose(x, g["act_2"], rtol=0, atol=3e-5 * np.abs(g["act_2"]).max())
    fin = np.isfinite(g["eig"])
    np.testing.assert_allclose(eig[fin], g["eig"][fin], rtol=2e-4)
