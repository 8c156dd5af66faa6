"""GPU parity: block glue (embedding, LayerNorm, conv+SiLU, elementwise), the FFMA GEMM with its epilogues, linear attention."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import eigb200.ops as ops
    return ops


dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_embedding(ops):
    rng = np.random.default_rng(0)
    V, D, B, T = 100, 64, 5, 17
    w = rng.normal(size=(V, D)).astype(np.float32); pos = rng.normal(size=(32, D)).astype(np.float32)
    ids = rng.integers(0, V, (B, T))
    np.testing.assert_array_equal(ops.embedding(dev(ids), dev(w)).cpu().numpy(), w[ids])
    np.testing.assert_array_equal(ops.embedding(dev(ids), dev(w), dev(pos)).cpu().numpy(), w[ids] + pos[:T][None])


@pytest.mark.parametrize("D", [32, 128, 512, 1000])
def test_layernorm(ops, D):
    rng = np.random.default_rng(D)
    x = (rng.normal(size=(7, 19, D)) * 3 + 1).astype(np.float32)
    w = rng.normal(size=D).astype(np.float32); b = rng.normal(size=D).astype(np.float32)
    out = ops.layernorm(dev(x), dev(w), dev(b)).cpu().numpy()
    ref = O.layer_norm(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64))
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("k", [1, 3, 4, 8])
def test_conv_silu(ops, k):
    rng = np.random.default_rng(k)
    B, T, C, ld = 3, 150, 70, 96
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    w = rng.normal(size=(C, k)).astype(np.float32); b = rng.normal(size=C).astype(np.float32)
    out = ops.conv_silu(dev(buf), ld, dev(w), dev(b), B, T, C).cpu().numpy()
    ref = O.causal_depthwise_conv_silu(buf.reshape(B, T, ld)[..., :C].astype(np.float64), w.astype(np.float64), b.astype(np.float64))
    np.testing.assert_allclose(out.reshape(B, T, C), ref, rtol=1e-5, atol=1e-6)


def test_elementwise(ops):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(1031,)).astype(np.float32) * 3; b = rng.normal(size=(1031,)).astype(np.float32) * 3
    np.testing.assert_array_equal(ops.add(dev(a), dev(b)).cpu().numpy(), a + b)
    np.testing.assert_allclose(ops.mul_silu(dev(a), dev(b)).cpu().numpy(), a * (b / (1 + np.exp(-b.astype(np.float64)))), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ops.gelu(dev(a)).cpu().numpy(), O.gelu_erf(a.astype(np.float64)), rtol=1e-5, atol=1e-6)
    s = rng.normal(size=(8,)).astype(np.float32); m = rng.normal(size=(13, 8)).astype(np.float32)
    np.testing.assert_array_equal(ops.scale_cols(dev(m), dev(s)).cpu().numpy(), m * s)


@pytest.mark.parametrize("M,N,K", [(300, 161, 128), (1000, 128, 128), (257, 50, 32), (64, 96, 33), (5, 7, 3), (2048, 256, 128)])
@pytest.mark.parametrize("epi", ["none", "gelu", "residual", "glu_residual"])
def test_linear_simt(ops, M, N, K, epi):
    if epi == "glu_residual" and N % 2:
        N += 1
    rng = np.random.default_rng(M + N + K)
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    nout = N // 2 if epi == "glu_residual" else N
    r = rng.normal(size=(M, nout)).astype(np.float32)
    out = ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r) if "residual" in epi else None, mode="simt").cpu().numpy()
    z = a.astype(np.float64) @ w.astype(np.float64).T + bias
    if epi == "gelu":
        ref = O.gelu_erf(z)
    elif epi == "residual":
        ref = z + r
    elif epi == "glu_residual":
        ref = z[:, :nout] * O.sigmoid(z[:, nout:]) + r
    else:
        ref = z
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_linear_strided_output(ops):
    rng = np.random.default_rng(5)
    M, N, K = 100, 161, 128
    a = rng.normal(size=(M, K)).astype(np.float32); w = rng.normal(size=(N, K)).astype(np.float32)
    out = ops.linear(dev(a), dev(w), None, ldc=164, mode="simt").cpu().numpy()
    assert out.shape == (M, 164)
    np.testing.assert_allclose(out[:, :N], a.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("H,d,dv,normalise", [(2, 16, 16, True), (1, 64, 64, True), (4, 32, 64, False), (1, 128, 128, True), (8, 64, 64, False)])
def test_linattn_forward(ops, H, d, dv, normalise):
    rng = np.random.default_rng(d + dv)
    B, T = 2, 45
    ld = 2 * H * d + H * dv + 3
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    gate = rng.uniform(0.1, 1.0, (B, T, H)).astype(np.float32)
    out = ops.linattn_forward(dev(buf), ld, 0, H * d, 2 * H * d, B, T, H, d, dv, gate=None if normalise else dev(gate),
                              phi_elu=True, normalise=normalise, kscale=1.0 if normalise else 0.25).cpu().numpy()
    b3 = buf.reshape(B, T, ld).astype(np.float64)
    q = O.elu(b3[..., :H * d].reshape(B, T, H, d)) + 1; k = O.elu(b3[..., H * d:2 * H * d].reshape(B, T, H, d)) + 1
    v = b3[..., 2 * H * d:2 * H * d + H * dv].reshape(B, T, H, dv)
    kv = np.cumsum(np.einsum("bthd,bthe->bthde", k * (1.0 if normalise else 0.25), v), axis=1)
    num = np.einsum("bthd,bthde->bthe", q, kv)
    if normalise:
        ref = num / np.einsum("bthd,bthd->bth", q, np.cumsum(k, axis=1))[..., None]
    else:
        ref = num * gate[..., None]
    np.testing.assert_allclose(out.reshape(B, T, H, dv), ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("T", [24, 45, 64, 7])
@pytest.mark.parametrize("H,d,dv,normalise", [(1, 64, 64, True), (4, 32, 64, False), (2, 16, 32, True), (1, 128, 128, False)])
def test_linattn_forward_column_owner_tail(ops, T, H, d, dv, normalise):
    """The column-owner kernel (d in {16,32,64,128}, dv % 32 == 0, ld % 4 == 0: the C1 / C5 shapes) with T % 16 != 0: phi must reach the k rows of a
    partial last chunk (`tc * 2 * D_` covered q rows only)."""
    rng = np.random.default_rng(1000 * T + d + dv)
    B = 3
    ld = 2 * H * d + H * dv + 4                                   # multiple of 4: 16-byte cp.async pieces => column-owner kernel
    assert ld % 4 == 0
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    gate = rng.uniform(0.1, 1.0, (B, T, H)).astype(np.float32)
    out = ops.linattn_forward(dev(buf), ld, 0, H * d, 2 * H * d, B, T, H, d, dv, gate=None if normalise else dev(gate),
                              phi_elu=True, normalise=normalise, kscale=1.0 if normalise else 0.25).cpu().numpy()
    b3 = buf.reshape(B, T, ld).astype(np.float64)
    q = O.elu(b3[..., :H * d].reshape(B, T, H, d)) + 1; k = O.elu(b3[..., H * d:2 * H * d].reshape(B, T, H, d)) + 1
    v = b3[..., 2 * H * d:2 * H * d + H * dv].reshape(B, T, H, dv)
    kv = np.cumsum(np.einsum("bthd,bthe->bthde", k * (1.0 if normalise else 0.25), v), axis=1)
    num = np.einsum("bthd,bthde->bthe", q, kv)
    ref = num / np.einsum("bthd,bthd->bth", q, np.cumsum(k, axis=1))[..., None] if normalise else num * gate[..., None]
    np.testing.assert_allclose(out.reshape(B, T, H, dv), ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max())


def _linattn_ref(buf, B, T, H, d, dv, q_off, k_off, v_off, normalise, gate, kscale, phi=True):
    b3 = buf.reshape(B, T, -1).astype(np.float64)
    q = b3[..., q_off:q_off + H * d].reshape(B, T, H, d); k = b3[..., k_off:k_off + H * d].reshape(B, T, H, d)
    if phi:
        q = O.elu(q) + 1; k = O.elu(k) + 1
    v = b3[..., v_off:v_off + H * dv].reshape(B, T, H, dv)
    kv = np.cumsum(np.einsum("bthd,bthe->bthde", k * kscale, v), axis=1)
    num = np.einsum("bthd,bthde->bthe", q, kv)
    if normalise:
        return num / np.einsum("bthd,bthd->bth", q, np.cumsum(k, axis=1))[..., None]
    return num * gate[..., None]


@pytest.mark.parametrize("T", [200, 64, 130, 1])
@pytest.mark.parametrize("H,normalise,phi", [(3, True, True), (2, False, True), (1, False, False)])
def test_linattn_forward_chunked_mma(ops, T, H, normalise, phi, monkeypatch):
    """The chunked tensor-core form (k9_linattn_mma.cu: d = dv = 64, the C1 / C5 heads): several chunks, a partial last chunk, both scalings, against
    float64 and against the recurrent column-owner kernel (EIGB200_LINATTN_FORM=col)."""
    rng = np.random.default_rng(77 * T + H)
    B, d, dv = 2, 64, 64
    ld = 2 * H * d + H * dv + 4
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    if not phi:
        buf = np.abs(buf) + 0.05                                   # raw positive features
    gate = rng.uniform(0.1, 1.0, (B, T, H)).astype(np.float32)
    kw = dict(gate=None if normalise else dev(gate), phi_elu=phi, normalise=normalise, kscale=1.0 if normalise else 0.125)
    out = ops.linattn_forward(dev(buf), ld, 0, H * d, 2 * H * d, B, T, H, d, dv, **kw).cpu().numpy()
    ref = _linattn_ref(buf, B, T, H, d, dv, 0, H * d, 2 * H * d, normalise, gate, kw["kscale"], phi)
    np.testing.assert_allclose(out.reshape(B, T, H, dv), ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max())
    monkeypatch.setenv("EIGB200_LINATTN_FORM", "col")
    col = ops.linattn_forward(dev(buf), ld, 0, H * d, 2 * H * d, B, T, H, d, dv, **kw).cpu().numpy()
    np.testing.assert_allclose(out, col, rtol=2e-5, atol=1e-5 * np.abs(ref).max())
    e_mma = np.abs(out.reshape(ref.shape) - ref).max() / np.abs(ref).max(); e_col = np.abs(col.reshape(ref.shape) - ref).max() / np.abs(ref).max()
    print("linattn T=%d: max err / max |ref|: chunked mma %.2e, recurrent fp32 %.2e" % (T, e_mma, e_col))


@pytest.mark.parametrize("T", [150, 64, 5])
@pytest.mark.parametrize("kconv,full,normalise", [(4, True, False), (3, False, False), (4, True, True), (2, False, True)])
def test_linattn_forward_conv_fused(ops, T, kconv, full, normalise):
    """conv1d + SiLU fused into the chunked kernel's loader == eigb200_conv_silu followed by eigb200_linattn_forward (same FMA order, approximate-unit SiLU), for the norm-attention column order [v | q | k | n] and the linear-attention order [q | k | v], conv over everything or over q, k only."""
    rng = np.random.default_rng(5 * T + kconv)
    B, H, d, dv = 2, 2, 64, 64
    D = dqk = H * d
    if normalise:                                                  # lin-attention: [q | k | v]
        ld = 2 * dqk + D
        q_off, k_off, v_off = 0, dqk, 2 * dqk
        ncv = ld if full else 2 * dqk; col0 = 0
        chq, chk, chv = 0, dqk, (2 * dqk if full else -1)
    else:                                                          # norm-attention: [v | q | k | n]
        ld = D + 2 * dqk + 4
        q_off, k_off, v_off = D, D + dqk, 0
        ncv = D + 2 * dqk if full else 2 * dqk; col0 = 0 if full else D
        chq, chk, chv = (D, D + dqk, 0) if full else (0, dqk, -1)
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    cw = (rng.normal(size=(ncv, kconv)) * 0.5).astype(np.float32); cb = (rng.normal(size=ncv) * 0.2).astype(np.float32)
    gate = rng.uniform(0.1, 1.0, (B, T, H)).astype(np.float32)
    kw = dict(gate=None if normalise else dev(gate), phi_elu=True, normalise=normalise, kscale=1.0 if normalise else 0.125)
    assert ops.linattn_conv_fusable(dev(buf), ld, q_off, k_off, v_off, H, d, dv, kconv)
    fused = ops.linattn_forward_conv(dev(buf), ld, q_off, k_off, v_off, B, T, H, d, dv, dev(cw), dev(cb), chq, chk, chv, **kw).cpu().numpy()
    cbuf = dev(buf).clone()
    ops.conv_silu(dev(buf)[:, col0:], ld, dev(cw), dev(cb), B, T, ncv, out=cbuf[:, col0:], ldo=ld)
    split = ops.linattn_forward(cbuf, ld, q_off, k_off, v_off, B, T, H, d, dv, **kw).cpu().numpy()
    ref = _linattn_ref(cbuf.cpu().numpy(), B, T, H, d, dv, q_off, k_off, v_off, normalise, gate, kw["kscale"])
    # the same attention kernel; the fused loader's SiLU uses ex2.approx / rcp.approx (~3 ulp) where eigb200_conv_silu divides
    np.testing.assert_allclose(fused, split, rtol=5e-6, atol=2e-6 * np.abs(ref).max())
    np.testing.assert_allclose(fused.reshape(B, T, H, dv), ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max())


def test_linattn_nu_and_eta(ops):
    from conftest import load_golden
    g = load_golden("lin_softmax_extractor")
    D, dqk, H = [int(v) for v in g["dims"]]
    B, T, _ = g["x"].shape
    qk = ops.linear(dev(g["x"]), dev(g["weight"][:2 * dqk]), dev(g["bias"][:2 * dqk]), mode="simt")
    nu = ops.linattn_nu(qk, 2 * dqk, B, T, H, dqk // H, dqk)
    eta, counts = ops.ratio_hist(nu, 2)
    np.testing.assert_allclose(eta.cpu().numpy(), g["eta_lin"][..., 0], rtol=1e-5)
    with np.errstate(invalid="ignore"):
        exp = np.moveaxis(O.threshold_counts(eta.cpu().numpy(), O.THRESHOLDS_RADIUS, axis=1), 0, -1)
    np.testing.assert_array_equal(counts.cpu().numpy()[..., :7], exp)


def test_softmax_nu_and_eta_golden(ops):
    """get_eig_att_softmax (eval_eig.py:43-95) against the vector produced by executing the reference function."""
    from conftest import load_golden
    g = load_golden("lin_softmax_extractor")
    D, dqk, H = [int(v) for v in g["dims"]]
    B, T, _ = g["x"].shape
    qk = ops.linear(dev(g["x"]), dev(g["weight"][:2 * dqk]), dev(g["bias"][:2 * dqk]), mode="simt")
    nu, m = ops.softmax_nu(qk, 2 * dqk, B, T, H, dqk // H, dqk)
    eta, counts = ops.softmax_eta(nu, m)
    np.testing.assert_allclose(eta.cpu().numpy(), g["eta_sm"][..., 0], rtol=1e-4)     # fp32 score rounding order differs from cuBLAS / einsum
    with np.errstate(invalid="ignore"):
        exp = np.moveaxis(O.threshold_counts(eta.cpu().numpy(), O.THRESHOLDS_RADIUS, axis=1), 0, -1)
    np.testing.assert_array_equal(counts.cpu().numpy()[..., :7], exp)


@pytest.mark.parametrize("B,T,H,d,scale", [(2, 70, 2, 16, 1.0), (1, 33, 1, 64, 3.0), (3, 129, 4, 8, 0.3), (2, 2, 1, 4, 1.0)])
def test_softmax_nu_matches_quadratic_and_closed_forms(ops, B, T, H, d, scale):
    """Ragged T (tile tails), several heads, large logits (scale 3: row maxima far from 0) and the T = 2 minimum."""
    rng = np.random.default_rng(T * 7 + d)
    q = (rng.normal(size=(B, T, H, d)) * scale).astype(np.float32); k = (rng.normal(size=(B, T, H, d)) * scale).astype(np.float32)
    buf = np.concatenate([q.reshape(B * T, H * d), k.reshape(B * T, H * d)], axis=1)
    nu, m = ops.softmax_nu(dev(buf), 2 * H * d, B, T, H, d, H * d)
    eta, _ = ops.softmax_eta(nu, m)
    ref_q = O.softmax_eta_quadratic(q, k)[..., 0]; ref_c = O.softmax_eta_closed(q, k)[..., 0]
    np.testing.assert_allclose(ref_c, ref_q, rtol=1e-4)
    np.testing.assert_allclose(eta.cpu().numpy(), ref_q, rtol=2e-4)
    # the row maximum is an actual score (or the masked zero), so it is reproduced to fp32 rounding of the dot product
    s = np.einsum("bthd,bshd->btsh", q.astype(np.float64), k.astype(np.float64))
    mask = np.tril(np.ones((T, T)))[None, :, :, None]
    mref = (s * mask).max(axis=2)
    np.testing.assert_allclose(m.cpu().numpy(), mref, rtol=1e-5, atol=1e-5 * np.abs(s).max())
    assert (nu.cpu().numpy() >= 1.0 - 1e-6).all()


@pytest.mark.parametrize("B,T,H,d,dv", [(2, 70, 2, 16, 16), (1, 129, 1, 64, 64), (2, 33, 4, 8, 32)])
def test_softmax_attn_forward(ops, B, T, H, d, dv):
    """SelfAttention.forward (models/attention.py:14-35): k scaled before the product, additive -10000 causal mask, softmax, P V."""
    rng = np.random.default_rng(T + dv)
    ld = 2 * H * d + H * dv
    buf = rng.normal(size=(B * T, ld)).astype(np.float32)
    scale = 1.0 / np.sqrt(d)
    out = ops.softmax_attn_forward(dev(buf), ld, 0, H * d, 2 * H * d, B, T, H, d, dv, scale).cpu().numpy().reshape(B, T, H, dv)
    b3 = buf.reshape(B, T, ld).astype(np.float64)
    q = b3[..., :H * d].reshape(B, T, H, d); k = b3[..., H * d:2 * H * d].reshape(B, T, H, d); v = b3[..., 2 * H * d:].reshape(B, T, H, dv)
    sc = np.einsum("bthd,bshd->bhts", q, k * scale) + np.triu(np.full((T, T), -10000.0), 1)
    sc = sc - sc.max(axis=-1, keepdims=True)
    pr = np.exp(sc); pr /= pr.sum(axis=-1, keepdims=True)
    ref = np.einsum("bhts,bshd->bthd", pr, v)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max() + 1e-6)


# ---- K4 on the tensor cores ------------------------------------------------------------------------------------------
def _lin_ref(a, w, bias, epi, r):
    z = a.astype(np.float64) @ w.astype(np.float64).T + (bias if bias is not None else 0)
    nout = w.shape[0] // 2 if epi == "glu_residual" else w.shape[0]
    if epi == "gelu":
        return O.gelu_erf(z)
    if epi == "residual":
        return z + r
    if epi == "glu_residual":
        return z[:, :nout] * O.sigmoid(z[:, nout:]) + (r if r is not None else 0)
    return z


TCMODE = {"tf32x3": "tc3", "f16x3": "f16x3"}


@pytest.fixture(params=["tf32x3", "f16x3"])
def prec(request, ops):
    """Both error-compensated operand splits of the tensor-core GEMMs (3 kind::tf32 MMAs / 3 kind::f16 MMAs per product) are held to the SAME fp32-level
    bounds.  The fixture also makes the split the process default, which is what linear_ln / linear_glu_extract / linear_prepare use."""
    ops.set_gemm_precision(request.param)
    yield request.param
    ops.set_gemm_precision(None)
    assert not ops.gemm_overflow()                                 # no test below may leave the sticky overflow flag raised


@pytest.mark.parametrize("M,N,K", [(4096, 161, 128), (5000, 128, 128), (3000, 256, 128), (1000, 64, 64), (129, 96, 32), (777, 200, 256),
                                    (2048, 48, 100), (300000, 128, 128), (1300, 72, 96), (2000, 136, 160)])
@pytest.mark.parametrize("epi", ["none", "gelu", "residual", "glu_residual"])
def test_linear_tcgen05_3xtf32(ops, prec, M, N, K, epi):
    if epi == "glu_residual" and N % 2:
        N += 1
    rng = np.random.default_rng(M + N + K)
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    nout = N // 2 if epi == "glu_residual" else N
    ldc = (nout + 3) // 4 * 4
    r = rng.normal(size=(M, ldc)).astype(np.float32)
    out = ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r)[:, :nout] if "residual" in epi else None, mode=TCMODE[prec], ldc=ldc).cpu().numpy()
    ref = _lin_ref(a, w, bias, epi, r[:, :nout].astype(np.float64))
    # fp32-level accuracy: same bound the FFMA path is held to
    np.testing.assert_allclose(out[:, :nout], ref, rtol=1e-5, atol=1e-5)
    simt = ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r)[:, :nout] if "residual" in epi else None, mode="simt", ldc=ldc).cpu().numpy()
    assert np.abs(out[:, :nout] - simt[:, :nout]).max() <= 2e-5


@pytest.mark.parametrize("N", [129, 130, 144, 145, 160, 161, 176, 177, 191, 192])
@pytest.mark.parametrize("K,epi", [(128, "none"), (96, "gelu"), (64, "residual"), (32, "none"), (128, "residual")])
def test_linear_tcgen05_wide_plan(ops, prec, N, K, epi):
    """129..192 output columns: the single-CTA wide plan (all columns resident, 2 TMEM operand stages, partial last 32-column group, ragged last row tile)."""
    M = 1500
    rng = np.random.default_rng(N * 7 + K)
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    ldc = (N + 7) // 8 * 8
    r = rng.normal(size=(M, ldc)).astype(np.float32)
    out = torch.full((M, ldc), 7.0, device="cuda")
    ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r)[:, :N] if epi == "residual" else None, mode=TCMODE[prec], out=out)
    out = out.cpu().numpy()
    ref = _lin_ref(a, w, bias, epi, r[:, :N].astype(np.float64))
    np.testing.assert_allclose(out[:, :N], ref, rtol=1e-5, atol=1e-5)
    assert (out[:, N:] == 7.0).all()                               # the pad columns between N and ldc are never written


@pytest.mark.parametrize("M,N,K", [(4096, 552, 512), (3000, 512, 512), (2500, 1024, 512), (1000, 96, 384), (4100, 40, 260), (2048, 1544, 512),
                                    (70000, 256, 320)])
@pytest.mark.parametrize("epi", ["none", "gelu", "residual", "glu_residual"])
@pytest.mark.parametrize("mode", ["tc3", "f16x3"])
def test_linear_tcgen05_streamed_operands(ops, M, N, K, epi, mode):
    """Shapes whose weight slice cannot stay resident in shared memory (K > 256): the streamed-operand kernel (A pre-split into tf32 hi / lo, or into scaled
    fp16 hi / lo: half the operand bytes per product, kind::f16)."""
    rng = np.random.default_rng(M + N + K)
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    nout = N // 2 if epi == "glu_residual" else N
    ldc = (nout + 7) // 8 * 8
    r = rng.normal(size=(M, ldc)).astype(np.float32)
    out = ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r)[:, :nout] if "residual" in epi else None, mode=mode, ldc=ldc).cpu().numpy()
    assert not ops.gemm_overflow()
    ref = _lin_ref(a, w, bias, epi, r[:, :nout].astype(np.float64))
    # The tensor core adds every MMA into the fp32 accumulator with truncation, 3 * K / 8 times per output: the error grows linearly with K
    # (measured 3e-6 of the output scale at K = 512, against 1e-6 at K = 128) where the FFMA path's round-to-nearest sum grows like sqrt(K).
    np.testing.assert_allclose(out[:, :nout], ref, rtol=1e-5, atol=1e-5 * max(1.0, K / 256))
    assert np.abs(out[:, :nout] - ref).max() <= 6e-6 * np.abs(ref).max() * max(1.0, K / 256)


@pytest.mark.parametrize("mode", ["tc3", "f16x3"])
def test_linear_streamed_strided_input(ops, mode):
    """K > 256 on a column slice of a wider buffer (row stride 600): the fp16-split form reads the raw fp32 rows by TMA with that stride and converts them on the SM
    (gemm_tc_stream_kernel<EPI, true, true>); rows beyond M and columns beyond K are zero-filled by the tensor map."""
    rng = np.random.default_rng(11)
    M, N, K = 3001, 200, 324
    big = rng.normal(size=(M, 600)).astype(np.float32)
    w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32); bias = rng.normal(size=N).astype(np.float32)
    r = rng.normal(size=(M, N)).astype(np.float32)
    a = dev(big)[:, 8:8 + K]
    out = ops.linear(a, dev(w), dev(bias), epilogue="residual", residual=dev(r), mode=mode).cpu().numpy()
    assert not ops.gemm_overflow()
    ref = _lin_ref(big[:, 8:8 + K], w, bias, "residual", r.astype(np.float64))
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3"])
def test_linear_ln_streamed_operands(ops, prec):
    ops.set_gemm_precision(prec)
    try:
        _linear_ln_streamed(ops)
        assert not ops.gemm_overflow()
    finally:
        ops.set_gemm_precision(None)


def test_streamed_f16_overflow_raises_the_flag(ops):
    """An activation beyond 65504 / S_a cannot be represented in the fp16 split: the sticky flag must come up (eval_eig then reruns the pass with 3xTF32)."""
    a = torch.ones(2048, 512, device="cuda"); a[5, 7] = 1.0e4                        # 1e4 * 16 > 65504
    w = torch.randn(256, 512, device="cuda") / 512 ** 0.5
    ops.gemm_overflow()                                                               # clear
    ops.linear(a, w, None, mode="f16x3")
    assert ops.gemm_overflow()
    ops.linear(a, w, None, mode="tc3")
    assert not ops.gemm_overflow()


def _linear_ln_streamed(ops):
    rng = np.random.default_rng(3)
    M, N, K = 3000, 552, 512
    a = (rng.normal(size=(M, K)) * 2 + rng.normal(0, 1, (M, 1))).astype(np.float32)
    w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    gamma = rng.normal(1, 0.3, K).astype(np.float32); beta = rng.normal(0, 0.3, K).astype(np.float32)
    st = ops.rowstats(dev(a))
    out = ops.linear_ln(dev(a), st, dev(gamma), dev(beta), dev(w), None, ldc=N).cpu().numpy()
    ref = O.layer_norm(a.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64)) @ w.astype(np.float64).T
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-5)


def test_linear_tcgen05_plain_tf32_is_coarser(ops):
    rng = np.random.default_rng(9)
    M, N, K = 2048, 128, 128
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    e1 = np.abs(ops.linear(dev(a), dev(w), None, mode="tc1").cpu().numpy() - ref).max()
    e3 = np.abs(ops.linear(dev(a), dev(w), None, mode="tc3").cpu().numpy() - ref).max()
    assert e1 < 5e-3 and e3 < 2e-5 and e3 < e1 / 20


# ---- LayerNorm fused into the GEMM A operand: statistics from the producer kernels ------------------------------------------
def _ln_stats_ref(x, eps=1e-5):
    x = x.astype(np.float64)
    mu = x.mean(-1); var = ((x - mu[..., None]) ** 2).mean(-1)
    return mu, 1.0 / np.sqrt(var + eps)


@pytest.mark.parametrize("D", [32, 64, 128, 512])
def test_rowstats_sources_agree(ops, D):
    rng = np.random.default_rng(D)
    B, T, H = 3, 50, 2
    x = (rng.normal(size=(B, T, D)) * rng.uniform(0.1, 3, (B, T, 1)) + rng.normal(0, 5, (B, T, 1))).astype(np.float32)   # large means: cancellation test
    mu, rstd = _ln_stats_ref(x)
    st = ops.rowstats(dev(x)).cpu().numpy()
    np.testing.assert_allclose(st[..., 0], mu, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st[..., 1], rstd, rtol=2e-5)
    # the K1 extractor emits the same statistics while it streams x
    W = rng.normal(size=(H, D)).astype(np.float32) * 0.1
    out = torch.empty(B, T, 2, device="cuda")
    ops.mamba2_eig(dev(x), dev(W), dev(np.zeros(H, np.float32)), dev(np.zeros(H, np.float32)), rowstats_out=out)
    np.testing.assert_allclose(out.cpu().numpy()[..., 0], mu, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out.cpu().numpy()[..., 1], rstd, rtol=2e-5)
    # ... and so does the embedding
    V = 40
    word = (rng.normal(size=(V, D)) + 3).astype(np.float32); pos = rng.normal(size=(T, D)).astype(np.float32)
    ids = rng.integers(0, V, (B, T))
    est = torch.empty(B, T, 2, device="cuda")
    e = ops.embedding(dev(ids), dev(word), dev(pos), rowstats_out=est).cpu().numpy()
    mu2, rstd2 = _ln_stats_ref(e)
    np.testing.assert_allclose(est.cpu().numpy()[..., 0], mu2, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(est.cpu().numpy()[..., 1], rstd2, rtol=2e-5)


@pytest.mark.parametrize("M,N,K", [(4096, 161, 128), (1000, 50, 32), (3000, 384, 128), (700, 96, 64)])
def test_linear_ln_fused(ops, prec, M, N, K):
    rng = np.random.default_rng(M + N)
    a = (rng.normal(size=(M, K)) * 2 + rng.normal(0, 1, (M, 1))).astype(np.float32)
    w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    gamma = rng.normal(1, 0.3, K).astype(np.float32); beta = rng.normal(0, 0.3, K).astype(np.float32)
    st = ops.rowstats(dev(a))
    ldc = (N + 3) // 4 * 4
    out = ops.linear_ln(dev(a), st, dev(gamma), dev(beta), dev(w), None, ldc=ldc).cpu().numpy()
    ref = O.layer_norm(a.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64)) @ w.astype(np.float64).T
    np.testing.assert_allclose(out[:, :N], ref, rtol=2e-5, atol=2e-5)
    unfused = ops.linear(ops.layernorm(dev(a), dev(gamma), dev(beta)), dev(w), None, mode=TCMODE[prec], ldc=ldc).cpu().numpy()
    assert np.abs(out[:, :N] - unfused[:, :N]).max() <= 2e-5


@pytest.mark.parametrize("M,D", [(4096, 128), (3000, 64), (2049, 256), (1500, 32)])
def test_glu_extract_partials_match_k1(ops, prec, M, D):
    """The GLU + residual GEMM's extractor partials, combined by eigb200_mamba2_eig_partials, against K1 on the GEMM's own output: lambda within fp32
    reassociation, LayerNorm statistics, and bin counts exact for the lambdas each path wrote."""
    rng = np.random.default_rng(M + D)
    B, T = 1, M
    a = rng.normal(size=(M, D)).astype(np.float32); w = (rng.normal(size=(2 * D, D)) / np.sqrt(D)).astype(np.float32)
    bias = rng.normal(size=2 * D).astype(np.float32); r = (rng.normal(size=(M, D)) * 1.5 + rng.normal(0, 1, (M, 1))).astype(np.float32)
    wg = (rng.normal(size=D) / np.sqrt(D) * 3).astype(np.float32)
    dtb = np.array([-1.5], np.float32); Al = np.log(np.array([3.0], np.float32))
    out, part = ops.linear_glu_extract(dev(a), dev(w), dev(bias), dev(r), dev(wg))
    plain = ops.linear(dev(a), dev(w), dev(bias), epilogue="glu_residual", residual=dev(r), mode=TCMODE[prec])
    assert torch.equal(out, plain)                                                     # the extra epilogue work does not touch the GEMM result
    st = torch.empty(B, T, 2, device="cuda")
    lam, counts = ops.mamba2_eig_partials(part, B, T, dev(dtb), dev(Al), rowstats_out=st)
    st_ref = torch.empty(B, T, 2, device="cuda")
    lam_ref, counts_ref = ops.mamba2_eig(out.reshape(B, T, D), dev(wg[None]), dev(dtb), dev(Al), rowstats_out=st_ref)
    l1, l0 = lam.cpu().numpy(), lam_ref.cpu().numpy()
    err = np.abs(l1 - l0) / (l0 * (1 + np.abs(np.log(l0))) + 1e-37)
    assert err.max() < 2e-6, err.max()
    np.testing.assert_allclose(st.cpu().numpy()[..., 0], st_ref.cpu().numpy()[..., 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.cpu().numpy()[..., 1], st_ref.cpu().numpy()[..., 1], rtol=2e-5)
    with np.errstate(invalid="ignore"):
        exp = np.moveaxis(O.threshold_counts(np.sqrt(np.power(l1, 2)), O.THRESHOLDS_RADIUS, axis=1), 0, -1)
    np.testing.assert_array_equal(counts.cpu().numpy()[..., :7], exp)
    assert (counts.cpu().numpy()[..., 7] == T).all()


@pytest.mark.parametrize("M,N,K,epi", [(4096, 161, 128, "none"), (3000, 128, 128, "gelu"), (2048, 256, 128, "glu_residual"), (1500, 96, 64, "residual")])
def test_prepared_weights_bit_identical(ops, prec, M, N, K, epi):
    """eigb200_linear_prepare once + d_W = NULL calls == per-call preparation, bit for bit (same kernels, same operands)."""
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).cuda(); w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda(); b = torch.randn(N, generator=g).cuda()
    nout = N // 2 if epi == "glu_residual" else N
    r = torch.randn(M, nout, generator=g).cuda() if "residual" in epi else None
    ldc = (nout + 3) // 4 * 4                                      # the tensor-core path wants 16-byte aligned output rows
    ref = ops.linear(a, w, b, epilogue=epi, residual=r, mode=TCMODE[prec], ldc=ldc)
    ws = ops.linear_prepare(w, b, epi)
    assert ws is not None
    for _ in range(2):                                             # the workspace is read-only for the GEMM: a second call sees the same operands
        out = ops.linear(a, w, b, epilogue=epi, residual=r, mode=TCMODE[prec], ldc=ldc, prepared=ws)
        assert torch.equal(out[:, :nout], ref[:, :nout])
    other = "f16x3" if prec == "tf32x3" else "tc3"                 # a prepared workspace holds the operands of ONE precision: asking for the other is refused
    with pytest.raises(Exception):
        ops.linear(a, w, b, epilogue=epi, residual=r, mode=other, ldc=ldc, prepared=ws)
    with pytest.raises(Exception):
        ops.linear(a, w, b, epilogue=epi, residual=r, mode="simt", ldc=ldc, prepared=ws)


def test_prepared_weights_layernorm_and_extract(ops, prec):
    g = torch.Generator().manual_seed(5)
    M, K = 4096, 128
    a = (torch.randn(M, K, generator=g) * 2 + 0.5).cuda()
    gamma = (1 + 0.1 * torch.randn(K, generator=g)).cuda(); beta = (0.1 * torch.randn(K, generator=g)).cuda()
    stats = torch.stack([a.mean(-1), torch.rsqrt(a.var(-1, unbiased=False) + 1e-5)], -1).contiguous()
    w = (torch.randn(161, K, generator=g) / K ** 0.5).cuda()
    ref = ops.linear_ln(a, stats, gamma, beta, w, None, ldc=168)
    ws = ops.linear_prepare(w, None, "none", gamma, beta)
    out = ops.linear_ln(a, stats, gamma, beta, w, None, ldc=168, prepared=ws)
    assert torch.equal(out[:, :161], ref[:, :161])
    wg = (torch.randn(256, K, generator=g) / K ** 0.5).cuda(); bg = torch.randn(256, generator=g).cuda()
    res = torch.randn(M, 128, generator=g).cuda(); wdt = torch.randn(128, generator=g).cuda()
    o1, p1 = ops.linear_glu_extract(a, wg, bg, res, wdt)
    ws2 = ops.linear_prepare(wg, bg, "glu_residual")
    o2, p2 = ops.linear_glu_extract(a, wg, bg, res, wdt, prepared=ws2)
    assert torch.equal(o1, o2) and torch.equal(p1, p2)


def test_linear_prepare_no_resident_plan(ops):
    w = torch.randn(512, 512).cuda()
    assert ops.linear_prepare(w, None, "none") is None             # K > 256: streamed-operand kernel, prepared per call


@pytest.mark.parametrize("scale,wscale", [(1e-3, 1.0), (300.0, 1.0), (1.0, 1e-4), (1.0, 50.0), (1e-2, 1e-3)])
@pytest.mark.parametrize("epi,N", [("none", 161), ("gelu", 128), ("glu_residual", 256)])
def test_linear_f16_split_operand_scales(ops, scale, wscale, epi, N):
    """The fp16 split keeps fp32-level accuracy off the unit scale: S_w is chosen from max |w| at preparation (weights of 1e-4 or 50 are as good as
    0.1), activations carry S_a = 16.  Bound: 2e-6 of sum_k |a||w| per output (the 3xTF32 path passes the same bound; a plain fp32 SGEMM sits at ~3e-7)."""
    M, K = 3000, 128
    rng = np.random.default_rng(int(scale * 1000) + N)
    a = (rng.normal(size=(M, K)) * scale).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K) * wscale).astype(np.float32)
    nout = N // 2 if epi == "glu_residual" else N
    ldc = (nout + 7) // 8 * 8
    bound = 2e-6 * (np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T)
    for mode in ("f16x3", "tc3"):
        out = ops.linear(dev(a), dev(w), None, epilogue="none", mode=mode, ldc=(N + 7) // 8 * 8).cpu().numpy()[:, :N]
        err = np.abs(out - a.astype(np.float64) @ w.astype(np.float64).T)
        assert (err <= bound + 1e-30).all(), (mode, float((err / (bound + 1e-300)).max()))
    if epi != "none" and scale * wscale <= 1.0:                    # and through the fused epilogues (large pre-activations make GLU / GELU ill-conditioned: value x d sigmoid)
        bias = rng.normal(size=N).astype(np.float32); r = rng.normal(size=(M, ldc)).astype(np.float32)
        out = ops.linear(dev(a), dev(w), dev(bias), epilogue=epi, residual=dev(r)[:, :nout] if "residual" in epi else None, mode="f16x3", ldc=ldc).cpu().numpy()
        ref = _lin_ref(a, w, bias, epi, r[:, :nout].astype(np.float64))
        # an fp32 GEMM's error scales with sum_k |a||w| of the pre-activation, not with the (possibly cancelled) output: the absolute term follows it
        np.testing.assert_allclose(out[:, :nout], ref, rtol=1e-5, atol=1e-5 + float(bound.max()))
    assert not ops.gemm_overflow()


def test_linear_f16_split_overflow_is_loud(ops):
    """An activation beyond 65504 / S_a cannot be represented by the fp16 split: the result is non-finite there AND the sticky flag is raised, so the
    caller can rerun with 3xTF32 (which has fp32's exponent range) -- never a silently wrong finite number."""
    M, N, K = 2048, 128, 128
    rng = np.random.default_rng(0)
    a = rng.normal(size=(M, K)).astype(np.float32); w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    a[1000, 5] = 1e5
    assert not ops.gemm_overflow()
    out = ops.linear(dev(a), dev(w), None, mode="f16x3").cpu().numpy()
    assert ops.gemm_overflow(reset=False) and ops.gemm_overflow()                      # sticky until reset ...
    assert not ops.gemm_overflow()                                                     # ... and cleared by it
    assert not np.isfinite(out[1000]).all()
    good = np.ones(M, bool); good[1000] = False
    np.testing.assert_allclose(out[good], (a.astype(np.float64) @ w.astype(np.float64).T)[good], rtol=1e-5, atol=1e-5)   # other rows are unaffected
    out3 = ops.linear(dev(a), dev(w), None, mode="tc3").cpu().numpy()
    np.testing.assert_allclose(out3, a.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-5, atol=2e-3)
    assert not ops.gemm_overflow()


@pytest.mark.parametrize("M,K1", [(4096, 128), (3001, 128), (129, 128), (2048, 64), (1500, 96), (40000, 32)])
@pytest.mark.parametrize("extract", [True, False])
@pytest.mark.parametrize("form", ["n", "t"])
def test_out_glu_fused_matches_two_kernels(ops, M, K1, extract, form, monkeypatch):
    """eigb200_out_glu_fused (GELU(out_proj) never leaves the SM) against the two-kernel form on the same prepared fp16-split operands, and against fp64.
    form "t" = the transposed kernel (k8_tail_fused_t.cu) where its shape conditions hold (K1 = 128), "n" = gemm_out_glu_kernel."""
    monkeypatch.setenv("EIGB200_TAIL_FORM", form)
    ops.set_gemm_precision("f16x3")
    try:
        D = 128
        rng = np.random.default_rng(M + K1)
        y = rng.normal(size=(M, K1)).astype(np.float32)
        w1 = (rng.normal(size=(D, K1)) / np.sqrt(K1)).astype(np.float32); b1 = (rng.normal(size=D) * 0.3).astype(np.float32)
        w2 = (rng.normal(size=(2 * D, D)) / np.sqrt(D)).astype(np.float32); b2 = rng.normal(size=2 * D).astype(np.float32)
        r = (rng.normal(size=(M, D)) * 1.5).astype(np.float32); wg = (rng.normal(size=D) / np.sqrt(D) * 3).astype(np.float32)
        assert ops.out_glu_fused_supported(D, K1)
        ws1 = ops.linear_prepare(dev(w1), dev(b1), "gelu"); ws2 = ops.linear_prepare(dev(w2), dev(b2), "glu_residual")
        out, part = ops.out_glu_fused(dev(y), ws1, dev(b1), ws2, dev(b2), dev(r), dev(wg) if extract else None)
        o = ops.linear(dev(y), dev(w1), dev(b1), epilogue="gelu", mode="f16x3")
        if extract:
            ref2, part2 = ops.linear_glu_extract(o, dev(w2), dev(b2), dev(r), dev(wg))
        else:
            ref2 = ops.linear(o, dev(w2), dev(b2), epilogue="glu_residual", residual=dev(r), mode="f16x3"); part2 = None
        # same operands, same split, same order of the accumulation; GEMM 2 reads its A operand from shared memory instead of TMEM and the residual add
        # contracts differently, so the two agree to 2 ulp (measured: <= 4.8e-7 on O(1) outputs, 20 % of the elements differ), not bit for bit
        np.testing.assert_allclose(out.cpu().numpy(), ref2.cpu().numpy(), rtol=1e-6, atol=1e-6)
        if extract:
            np.testing.assert_allclose(part.cpu().numpy(), part2.cpu().numpy(), rtol=1e-5, atol=2e-5)
        else:
            assert part is None
        o64 = _lin_ref(y, w1, b1, "gelu", None)
        z = o64 @ w2.astype(np.float64).T + b2
        ref = z[:, :D] / (1 + np.exp(-z[:, D:])) + r
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
        assert not ops.gemm_overflow()
    finally:
        ops.set_gemm_precision(None)
