"""GPU parity: K1 family (fused gate projection -> eigenvalue -> bin counts) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden, assert_eig_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import eigb200.ops as ops
    return ops


def _dev(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def _radius32(lam):
    """float32 radius of a real float32 array the way the reference forms it (eval_eig.py:605-606)."""
    lam = np.asarray(lam, np.float32)
    return np.sqrt(np.power(lam, 2) + np.power(np.zeros_like(lam), 2))


def _counts_from_values(vals, thr, compare="float64"):
    """Oracle bin counts (B,inner,nthr+1) of values (B,N,inner)."""
    with np.errstate(invalid="ignore"):
        c = O.threshold_counts(vals, thr, axis=1, compare=compare)       # (nb,B,inner)
    return np.moveaxis(c, 0, -1)


def test_mamba2_eig_golden(ops):
    g = load_golden("mamba2_extractor")
    D, G, N, H = [int(v) for v in g["dims"]]
    W_dt = g["in_proj_weight"][O.mamba2_dt_rows(D, G, N, H)]
    lam, counts = ops.mamba2_eig(_dev(g["x"]), _dev(W_dt), _dev(g["dt_bias"]), _dev(g["A_log"]))
    lam = lam.cpu().numpy(); counts = counts.cpu().numpy()
    ref64 = O.mamba2_eig(g["x"], g["in_proj_weight"], g["dt_bias"], g["A_log"], D, G, N, H, np.float64)[..., 0]
    assert_eig_close(lam, ref64, rtol=1e-5, what="lambda vs fp64 oracle")
    assert_eig_close(lam, g["lam"][..., 0], rtol=2e-5, what="lambda vs reference fp32 output")
    # bin counts are EXACT for the values the kernel produced
    exp = _counts_from_values(_radius32(lam), O.THRESHOLDS_RADIUS)
    np.testing.assert_array_equal(counts[..., :7], exp)
    np.testing.assert_array_equal(counts[..., 7], lam.shape[1])


@pytest.mark.parametrize("B,T,D,H", [(3, 64, 128, 1), (2, 37, 512, 8), (5, 130, 64, 3), (2, 50, 96, 12), (1, 3, 32, 2), (4, 512, 128, 1)])
def test_mamba2_eig_shapes(ops, B, T, D, H):
    rng = np.random.default_rng(B * 1000 + T)
    x = rng.normal(0, 1.5, (B, T, D)).astype(np.float32)
    W = (rng.normal(0, 1, (H, D)) / np.sqrt(D) * 3).astype(np.float32)
    dt_bias = rng.normal(-2, 2, H).astype(np.float32)
    A_log = np.log(rng.uniform(1, 16, H)).astype(np.float32)
    lam, counts = ops.mamba2_eig(_dev(x), _dev(W), _dev(dt_bias), _dev(A_log))
    lam = lam.cpu().numpy(); counts = counts.cpu().numpy()
    z = x.astype(np.float64) @ W.astype(np.float64).T + dt_bias
    ref = np.exp(O.softplus(z) * -np.exp(A_log.astype(np.float64)))
    assert_eig_close(lam, ref, rtol=1e-5)
    np.testing.assert_array_equal(counts[..., :7], _counts_from_values(_radius32(lam), O.THRESHOLDS_RADIUS))
    np.testing.assert_array_equal(counts[..., 7], T)
    # statistics-only call (no eigenvalue array materialised) gives the same counts
    _, c2 = ops.mamba2_eig(_dev(x), _dev(W), _dev(dt_bias), _dev(A_log), want_lam=False)
    np.testing.assert_array_equal(c2.cpu().numpy(), counts)


def test_mamba2_eig_bf16_activations(ops):
    rng = np.random.default_rng(3)
    B, T, D, H = 4, 96, 256, 4
    x = rng.normal(0, 1, (B, T, D)).astype(np.float32)
    W = (rng.normal(0, 1, (H, D)) / np.sqrt(D)).astype(np.float32)
    dt_bias = rng.normal(-2, 1, H).astype(np.float32); A_log = np.log(rng.uniform(1, 16, H)).astype(np.float32)
    xb = _dev(x).to(torch.bfloat16)
    lam, _ = ops.mamba2_eig(xb, _dev(W), _dev(dt_bias), _dev(A_log))
    # exact in the bf16-rounded input ...
    xr = xb.float().cpu().numpy().astype(np.float64)
    ref_r = np.exp(O.softplus(xr @ W.astype(np.float64).T + dt_bias) * -np.exp(A_log.astype(np.float64)))
    assert_eig_close(lam.cpu().numpy(), ref_r, rtol=1e-5)
    # ... and within the north-star's 1e-2 of the fp32-input result
    ref = np.exp(O.softplus(x.astype(np.float64) @ W.astype(np.float64).T + dt_bias) * -np.exp(A_log.astype(np.float64)))
    assert_eig_close(lam.cpu().numpy(), ref, rtol=1e-2, what="bf16 activations vs fp32-input truth")


def test_edge_semantics_and_compare_modes(ops):
    g = load_golden("thresholds")
    for key, thr in (("v32", O.THRESHOLDS_RADIUS), ("v64", O.THRESHOLDS_RADIUS), ("ph", O.THRESHOLDS_PHASE)):
        v = g[key]                                                      # (B,N,H,L)
        B, N = v.shape[:2]
        for compare in ("float64", "float32"):
            _, counts = ops.ratio_hist(_dev(v), 0, thresholds=thr, compare=compare)
            exp = _counts_from_values(v.reshape(B, N, -1), thr, compare)
            nb = len(thr) + 1
            np.testing.assert_array_equal(counts.cpu().numpy()[..., :nb], exp)
        # percentages exactly as the reference printed them (numpy 2 promotion = float64 compare)
        _, counts = ops.ratio_hist(_dev(v), 0, thresholds=thr)
        pct = np.moveaxis(counts.cpu().numpy()[..., : len(thr) + 1], -1, 0).reshape((len(thr) + 1,) + v.shape[:1] + v.shape[2:]) / N * 100
        np.testing.assert_array_equal(pct, g["p_" + key])
    half = torch.full((2, 8, 1), 0.5, device="cuda", dtype=torch.float64)
    _, c = ops.ratio_hist(half, 0)
    assert c.cpu().numpy()[..., :7].sum(-1).ravel().tolist() == [16, 16]             # edge value counted twice -> 200 %


@pytest.mark.parametrize("fn", ["exp", "elu", "softplus", "sigmoid"])
@pytest.mark.parametrize("use_off", [0, 1])
def test_normattn_eta_golden(ops, fn, use_off):
    g = load_golden("norm_extractor")
    D, dqk, H = [int(v) for v in g["dims"]]
    rows = O.normattn_rows(D, dqk, H)
    n = ops.normattn_gate(_dev(g["x"]), _dev(g["weight"][rows]), _dev(g["bias"][rows]), _dev(g["offset"]) if use_off else None, fn)
    eta, counts = ops.ratio_hist(n, 1)
    eta = eta.cpu().numpy(); counts = counts.cpu().numpy()
    ref = g["eta_%s_%d" % (fn, use_off)][..., 0]
    fin = np.isfinite(ref)
    assert (np.isfinite(eta) == fin).all()
    np.testing.assert_allclose(eta[fin], ref[fin], rtol=2e-4 if fn == "exp" else 2e-5)   # vs the reference's own fp32 eta; vs fp64: test_parity_fullshape_gpu.py
    # n itself against the fp64 formula
    x64 = g["x"].astype(np.float64)
    raw = x64 @ g["weight"][rows].astype(np.float64).T + g["bias"][rows] + (g["offset"] if use_off else 0)
    with np.errstate(over="ignore", under="ignore"):
        n64 = np.exp(-O.norm_fn_apply(fn, raw))
    nn_ = n.cpu().numpy()
    ok = n64 > 1e-30
    assert_eig_close(nn_[ok], n64[ok], rtol=2e-5, what="n")
    np.testing.assert_array_equal(counts[..., :7], _counts_from_values(eta, O.THRESHOLDS_RADIUS))
    np.testing.assert_array_equal(counts[..., 7], np.isfinite(eta).sum(axis=1))
    with pytest.raises(RuntimeError):
        ops.normattn_gate(_dev(g["x"]), _dev(g["weight"][rows]), _dev(g["bias"][rows]), None, "tanh")


def test_ratio_orientations_and_zero_patch(ops):
    a = torch.tensor([[[2.0], [0.0], [4.0], [1.0]]], device="cuda", dtype=torch.float64)
    nxt, _ = ops.ratio_hist(a, 1)
    cur, _ = ops.ratio_hist(a, 2)
    np.testing.assert_allclose(nxt.cpu().numpy().ravel(), [2e-23 / 2.0, 4.0 / 2e-23, 0.25])
    np.testing.assert_allclose(cur.cpu().numpy().ravel(), [2.0 / 2e-23, 2e-23 / 4.0, 4.0])


def test_count_moments(ops):
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1000, (37, 6, 8)).astype(np.int32)
    s, s2 = ops.count_moments(_dev(c))
    np.testing.assert_array_equal(s.cpu().numpy(), c.astype(np.int64).sum(0))
    np.testing.assert_array_equal(s2.cpu().numpy(), (c.astype(np.int64) ** 2).sum(0))
    mean, std = O.batch_mean_std_from_counts(s.cpu().numpy(), s2.cpu().numpy(), 512, 37)
    pct = c / 512 * 100
    np.testing.assert_allclose(mean, pct.mean(0), rtol=1e-12)
    np.testing.assert_allclose(std, pct.std(0), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("L,B,inner", [(4, 4096, 1), (3, 37, 6), (12, 1000, 8), (2, 5, 40), (1, 513, 3)])
def test_count_moments_layers(ops, L, B, inner):
    """The per-pass moments buffer (2,L,inner,8) -- what a rank contributes to the path's single all-reduce -- in one launch, exact in int64."""
    rng = np.random.default_rng(L * 1000 + B)
    c = rng.integers(0, 2049, (L, B, inner, 8)).astype(np.int32)
    m = ops.count_moments_layers(_dev(c)).cpu().numpy()
    assert m.shape == (2, L, inner, 8) and m.dtype == np.int64
    np.testing.assert_array_equal(m[0], c.astype(np.int64).sum(1))
    np.testing.assert_array_equal(m[1], (c.astype(np.int64) ** 2).sum(1))
    for l in range(L):                                            # agrees with the single-layer entry point
        s, s2 = ops.count_moments(_dev(c[l]))
        np.testing.assert_array_equal(s.cpu().numpy(), m[0, l]); np.testing.assert_array_equal(s2.cpu().numpy(), m[1, l])


def test_full_size_properties_c2(ops):
    """BASELINE config C2 K1 shape (4096 x 512 x 128, H=1): properties that need no CPU pass over 1 GB."""
    torch.manual_seed(0)
    B, T, D, H = 4096, 512, 128, 1
    x = torch.randn(B, T, D, device="cuda")
    W = torch.randn(H, D, device="cuda") * 0.3
    dt_bias = torch.full((H,), -1.0, device="cuda"); A_log = torch.log(torch.full((H,), 4.0, device="cuda"))
    lam, counts = ops.mamba2_eig(x, W, dt_bias, A_log)
    assert (counts[..., 7] == T).all()
    assert (counts[..., :7].sum(-1) >= T).all() and (counts[..., :7].sum(-1) <= T + 8).all()
    thr = torch.tensor(O.THRESHOLDS_RADIUS, device="cuda", dtype=torch.float64)
    l64 = torch.sqrt(lam * lam).double()
    first = ((l64 >= 0) & (l64 <= thr[0])).sum(1)
    mid = ((l64 >= thr[1]) & (l64 <= thr[2])).sum(1)
    assert torch.equal(first.int(), counts[..., 0]) and torch.equal(mid.int(), counts[..., 2])
    # checksum of checksums: a spot sample of rows against the fp64 formula
    idx = torch.randint(0, B, (16,), device="cuda")
    z = x[idx].double() @ W.double().T + dt_bias.double()
    ref = torch.exp(torch.nn.functional.softplus(z) * -torch.exp(A_log.double()))
    assert_eig_close(lam[idx].cpu().numpy(), ref.cpu().numpy(), rtol=1e-5)


def test_stats_allreduce_c_abi(ops):
    """The exchange step through the C ABI (eigb200_stats_comm_init_all / eigb200_stats_allreduce, NCCL dlopen'ed at run time): every visible GPU (1 on the
    driver's test box, 2+ under `gpurun --gpus N`) contributes the moments of its slice; the in-place int64 sum equals the moments of the whole batch."""
    import eigb200.dist as D
    ndev = min(torch.cuda.device_count(), 4)
    rng = np.random.default_rng(0)
    L_, B, H = 3, 16 * ndev, 2
    c = rng.integers(0, 513, (L_, B, H, 8)).astype(np.int32)
    comm = D.StatsComm(list(range(ndev)))
    try:
        parts = []
        for d in range(ndev):
            with torch.cuda.device(d):
                sl = torch.from_numpy(np.ascontiguousarray(c[:, d * 16:(d + 1) * 16])).cuda(d)
                parts.append(ops.count_moments_layers(sl))
        comm.allreduce(parts)
        for d in range(ndev):
            torch.cuda.synchronize(d)
        ref = np.stack([c.astype(np.int64).sum(1), (c.astype(np.int64) ** 2).sum(1)])
        for d in range(ndev):
            np.testing.assert_array_equal(parts[d].cpu().numpy(), ref)
    finally:
        comm.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("B,N,inner", [(7, 300, 5), (64, 512, 4), (3, 50, 70)])
def test_log_hist_and_quantiles(ops, dtype, B, N, inner):
    """On-device log-spaced histogram + quantiles per (layer, head) column: exact bin counts against NumPy on the same fp32 bin coordinate, quantiles within one
    bin width of np.quantile; special values (0, negative, NaN, beyond the range) land in their slots; a second call accumulates."""
    rng = np.random.default_rng(B + inner)
    v = np.exp(rng.normal(-3, 3, (B, N, inner))).astype(dtype)
    v[0, 0, 0] = 0.0; v[0, 1, 0] = -1.0; v[0, 2, 0] = np.nan; v[0, 3, 0] = 1e9; v[0, 4, 0] = 1e-12
    lo, hi, nb = 1e-8, 1e2, 512
    h = ops.log_hist(_dev(v), lo, hi, nb)
    hh = h.cpu().numpy()
    assert hh.shape == (inner, nb + 3) and hh.sum() == v.size
    x = v.astype(np.float32).reshape(-1, inner)
    with np.errstate(invalid="ignore", divide="ignore"):
        t = (np.log2(x) - np.float32(np.log2(lo))) * np.float32(nb / (np.log2(hi) - np.log2(lo)))
    slot = np.where(np.isnan(x), nb + 2, np.where(~(x > 0), 0, np.where(t < 0, 0, np.where(t >= nb, nb + 1, 1 + np.floor(np.nan_to_num(t, nan=0.0)).astype(np.int64)))))
    ref = np.stack([np.bincount(slot[:, c], minlength=nb + 3) for c in range(inner)])
    assert np.abs(hh - ref).sum() <= max(4, v.size // 20000)          # __log2f vs np.log2 may move a value sitting on a bin edge
    assert hh[0, 0] >= 3 and hh[0, nb + 1] >= 1 and hh[0, nb + 2] == 1
    qs = [0.05, 0.25, 0.5, 0.75, 0.95]
    q = ops.hist_quantiles(h, qs, lo, hi).cpu().numpy()
    width = (hi / lo) ** (1.0 / nb)
    n = B * N
    for c in range(1, inner):                                        # column 0 holds the planted specials
        srt = np.sort(v[:, :, c].astype(np.float64).ravel())
        for qi, qq in enumerate(qs):                                 # the value where the cumulative count reaches q n: between neighbouring order statistics, +- a bin
            k = int(np.ceil(qq * n))
            lo_v, hi_v = srt[max(k - 2, 0)] / width / 1.001, srt[min(k, n - 1)] * width * 1.001
            assert lo_v <= q[c, qi] <= hi_v, (c, qq, q[c, qi], lo_v, hi_v)
    h2 = ops.log_hist(_dev(v), lo, hi, nb, hist=h)
    np.testing.assert_array_equal(h2.cpu().numpy(), 2 * hh)
