"""GPU parity: K2a diagonal complex scan and K2b SSD selective scan vs the fp64 oracle.
Scan tolerance: |h - h_ref| <= 1e-5 * max_t |h_ref| per (sequence, channel)."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import eigb200.ops as ops
    return ops


def _scan_close(h, ref, rtol=1e-5, axis=1):
    h = np.asarray(h); ref = np.asarray(ref)
    scale = np.abs(ref).max(axis=axis, keepdims=True)
    err = np.abs(h - ref)
    assert (err <= rtol * scale + 1e-30).all(), "worst normalised error %.3g" % (err / (scale + 1e-30)).max()


def _lru_like_lambda(rng, P, r_min=0.9, r_max=0.999):
    u1, u2 = rng.uniform(size=P), rng.uniform(size=P)
    mag = np.sqrt(u1 * (r_max ** 2 - r_min ** 2) + r_min ** 2)
    return (mag * np.exp(1j * 6.28 * u2)).astype(np.complex64)


@pytest.mark.parametrize("B,T,P,reverse", [(4, 300, 64, False), (2, 2048, 256, False), (3, 77, 10, True), (1, 1, 32, False),
                                            (640, 64, 256, False), (2, 513, 33, True)])
def test_diag_scan(ops, B, T, P, reverse):
    rng = np.random.default_rng(B + T + P)
    lam = _lru_like_lambda(rng, P)
    Bu = (rng.normal(size=(B, T, P)) + 1j * rng.normal(size=(B, T, P))).astype(np.complex64)
    h = ops.diag_scan(torch.from_numpy(lam).cuda(), torch.from_numpy(Bu).cuda(), reverse=reverse).cpu().numpy()
    ref = O.diag_scan(lam.astype(np.complex128), Bu.astype(np.complex128), reverse=reverse)
    _scan_close(h, ref)


def test_diag_scan_linearity_and_roundtrip_full_size(ops):
    """C3-sized properties (no CPU pass over GBs): linearity in Bu and  Bu_t = h_t - lam h_{t-1}  (scan inverse)."""
    torch.manual_seed(0)
    B, T, P = 128, 2048, 256
    lam = torch.polar(torch.rand(P, device="cuda") * 0.09 + 0.9, torch.rand(P, device="cuda") * 6.28)
    Bu = torch.randn(B, T, P, dtype=torch.complex64, device="cuda")
    h = ops.diag_scan(lam, Bu)
    rec = h.clone()
    rec[:, 1:] -= lam * h[:, :-1]
    scale = h.abs().amax(dim=1, keepdim=True)
    assert ((rec - Bu).abs() <= 2e-5 * scale).all()
    h2 = ops.diag_scan(lam, 2.5 * Bu)
    assert ((h2 - 2.5 * h).abs() <= 2e-5 * scale * 2.5).all()


def test_lru_s5_lambda(ops):
    g = load_golden("lru_s5_eigs")
    for i in range(3):
        lam = ops.ssm_lambda("lru", torch.from_numpy(g["lru_nu_%d" % i]).cuda(), torch.from_numpy(g["lru_theta_%d" % i]).cuda()).cpu().numpy()
        np.testing.assert_allclose(lam[:, None], g["lru_eig_%d" % i], rtol=1e-5, atol=1e-7)
        lam = ops.ssm_lambda("s5_zoh", torch.from_numpy(g["s5_Lambda_re_%d" % i]).cuda(), torch.from_numpy(g["s5_Lambda_im_%d" % i]).cuda(),
                             torch.from_numpy(g["s5_log_step_%d" % i]).cuda()).cpu().numpy()
        np.testing.assert_allclose(lam[:, None], g["s5_eig_%d" % i], rtol=1e-5, atol=1e-7)
    h = load_golden("hippo_s4")
    lam = ops.ssm_lambda("s5_bilinear", torch.from_numpy(h["disc_Lambda"].real.copy()).cuda(), torch.from_numpy(h["disc_Lambda"].imag.copy()).cuda(),
                         torch.log(torch.from_numpy(h["disc_Delta"])).cuda()).cpu().numpy()
    np.testing.assert_allclose(lam, h["bil_L"], rtol=2e-5, atol=1e-7)


@pytest.fixture(params=["scan", "mma", "tc"])
def ssd_form(request, monkeypatch):
    """All forms of K2b: the recurrent scan kernels, the chunked mma.sync form (csrc/k2_ssd_mma.cuh) and the chunked tcgen05 form (csrc/k2_ssd_tc.cu: fused
    conv + SSD with head dim 128, d_state 16), each taken where its shape conditions hold."""
    monkeypatch.setenv("EIGB200_SSD_FORM", request.param)
    return request.param


def test_ssd_scan_golden(ops):
    g = load_golden("ssd_small")
    y = ops.ssd_scan(*(torch.from_numpy(g[k]).cuda() for k in ("x", "dt", "A", "Bm", "Cm", "D"))).cpu().numpy()
    ref = O.ssd_scan_sequential(g["x"], g["dt"], g["A"], g["Bm"], g["Cm"], g["D"])
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.abs(y - g["y_fla"]).max() <= 2e-5 * np.abs(ref).max()          # third-party (fla) output


@pytest.mark.parametrize("B,T,H,P,G,N", [(3, 70, 1, 128, 1, 16), (2, 33, 4, 16, 2, 8), (2, 40, 2, 64, 1, 128), (1, 96, 8, 8, 1, 64), (2, 20, 3, 20, 1, 4),
                                          (2, 512, 2, 128, 1, 16), (1, 37, 1, 64, 1, 8), (2, 3, 1, 256, 1, 16), (1, 64, 2, 64, 2, 4), (1, 100, 1, 128, 1, 16),
                                          (2, 50, 4, 64, 2, 16), (1, 16, 1, 128, 1, 16), (1, 17, 2, 192, 1, 16)])
def test_ssd_scan_shapes(ops, ssd_form, B, T, H, P, G, N):
    rng = np.random.default_rng(T + N)
    x = rng.normal(size=(B, T, H, P)).astype(np.float32)
    dt = O.softplus(rng.normal(-1, 1, (B, T, H))).astype(np.float32)
    A = -rng.uniform(1, 16, H).astype(np.float32)
    Bm = rng.normal(size=(B, T, G, N)).astype(np.float32); Cm = rng.normal(size=(B, T, G, N)).astype(np.float32)
    Dv = rng.normal(size=H).astype(np.float32)
    y, fs = ops.ssd_scan(*(torch.from_numpy(a).cuda() for a in (x, dt, A, Bm, Cm, Dv)), return_final_state=True)
    ref, st = O.ssd_scan_sequential(x, dt, A, Bm, Cm, Dv, return_state=True)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(y.cpu().numpy() - ref) <= 1e-5 * scale + 1e-7).all()
    assert np.abs(fs.cpu().numpy() - st).max() <= 1e-5 * np.abs(st).max() + 1e-7


@pytest.mark.parametrize("P,T", [(32, 75), (128, 75), (128, 256), (64, 33)])
@pytest.mark.parametrize("kconv", [4, 2, 0])
def test_mamba_conv_ssd_fused(ops, ssd_form, kconv, P, T):
    rng = np.random.default_rng(kconv)
    B, H, G, N = 3, 2, 1, 16
    C_ = H * P + 2 * G * N
    ldz = (C_ + H + 3) // 4 * 4
    z = rng.normal(size=(B, T, ldz)).astype(np.float32)
    cw = rng.normal(size=(C_, max(kconv, 1))).astype(np.float32) * 0.5; cb = rng.normal(size=C_).astype(np.float32) * 0.1
    dtb = rng.normal(-1, 1, H).astype(np.float32); Al = np.log(rng.uniform(1, 16, H)).astype(np.float32); Dv = rng.normal(size=H).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    y = ops.mamba_conv_ssd(dev(z), ldz, dev(cw) if kconv else None, dev(cb) if kconv else None, dev(dtb), dev(Al), dev(Dv), B, T, H, P, G, N).cpu().numpy()
    z64 = z.astype(np.float64)
    xBC = z64[..., :C_]
    if kconv:
        xBC = O.causal_depthwise_conv_silu(xBC, cw.astype(np.float64), cb.astype(np.float64))
    dt = O.softplus(z64[..., C_:C_ + H] + dtb)
    ref = O.ssd_scan_sequential(xBC[..., :H * P].reshape(B, T, H, P), dt, -np.exp(Al.astype(np.float64)),
                                xBC[..., H * P:H * P + G * N].reshape(B, T, G, N), xBC[..., H * P + G * N:].reshape(B, T, G, N), Dv.astype(np.float64))
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(y.reshape(B, T, H, P) - ref) <= 1e-5 * scale + 1e-6).all()


@pytest.mark.parametrize("B,T,H,G,kconv", [(2, 64, 1, 1, 4), (3, 512, 1, 1, 4), (2, 130, 2, 1, 4), (5, 1, 1, 1, 4), (4, 63, 1, 1, 3), (2, 65, 4, 2, 4), (310, 96, 1, 1, 4),
                                            (1, 1000, 1, 1, 1)])
def test_mamba_conv_ssd_tcgen05_form(ops, monkeypatch, B, T, H, G, kconv):
    """The chunked tcgen05 form (k2_ssd_tc.cu) on its own shapes: chunk-aligned and ragged sequence lengths (T % 64 in {0, 1, 2, 63, ...}), T < 64, several chunks
    (state carried through TMEM), heads sharing a B / C group, more (sequence, head) items than SMs (CTAs walk several items: state reset), short conv kernels.
    Same bound as the recurrent form: 1e-5 of the per-(sequence, channel) output scale."""
    monkeypatch.setenv("EIGB200_SSD_FORM", "tc")
    rng = np.random.default_rng(B * 1000 + T)
    P, N = 128, 16
    C_ = H * P + 2 * G * N
    ldz = (C_ + H + 7) // 8 * 8
    z = rng.normal(size=(B, T, ldz)).astype(np.float32)
    cw = rng.normal(size=(C_, kconv)).astype(np.float32) * 0.5; cb = rng.normal(size=C_).astype(np.float32) * 0.1
    dtb = rng.normal(-1, 1, H).astype(np.float32); Al = np.log(rng.uniform(1, 16, H)).astype(np.float32); Dv = rng.normal(size=H).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    y = ops.mamba_conv_ssd(dev(z), ldz, dev(cw), dev(cb), dev(dtb), dev(Al), dev(Dv), B, T, H, P, G, N)
    monkeypatch.setenv("EIGB200_SSD_FORM", "scan")
    y_scan = ops.mamba_conv_ssd(dev(z), ldz, dev(cw), dev(cb), dev(dtb), dev(Al), dev(Dv), B, T, H, P, G, N).cpu().numpy()
    y = y.cpu().numpy()
    nb = min(B, 6)                                                 # fp64 oracle on a few sequences (first and last), the recurrent kernel on all of them
    sel = np.unique(np.r_[np.arange(nb // 2), B - 1 - np.arange(nb - nb // 2)])
    z64 = z[sel].astype(np.float64)
    xBC = O.causal_depthwise_conv_silu(z64[..., :C_], cw.astype(np.float64), cb.astype(np.float64))
    dt = O.softplus(z64[..., C_:C_ + H] + dtb)
    ref = O.ssd_scan_sequential(xBC[..., :H * P].reshape(len(sel), T, H, P), dt, -np.exp(Al.astype(np.float64)),
                                xBC[..., H * P:H * P + G * N].reshape(len(sel), T, G, N), xBC[..., H * P + G * N:].reshape(len(sel), T, G, N), Dv.astype(np.float64))
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(y[sel].reshape(len(sel), T, H, P) - ref) <= 1e-5 * scale + 1e-6).all()
    sc2 = np.abs(y_scan).reshape(B, T, H * P).max(axis=1, keepdims=True)
    assert (np.abs(y - y_scan) <= 2e-5 * sc2 + 1e-6).all()


@pytest.mark.gpu
def test_pass_graph_lru_stack_equals_eager():
    """analysis.PassGraph: two LRU layer calls captured as one CUDA graph return exactly the eager result, also for a new input."""
    import torch
    import eigb200.analysis as A
    import eigb200.ssm as S
    rng = np.random.default_rng(5)
    P, Hd, B, T = 64, 32, 16, 96
    def mk():
        return {"nu_log": torch.from_numpy(np.log(-np.log(rng.uniform(0.8, 0.99, P))).astype(np.float32)).cuda(),
                "theta_log": torch.from_numpy(np.log(rng.uniform(0.01, 3.0, P)).astype(np.float32)).cuda(),
                "gamma_log": torch.from_numpy(rng.normal(-1, 0.2, P).astype(np.float32)).cuda(),
                "B_re": torch.from_numpy((rng.normal(size=(P, Hd)) / np.sqrt(2 * Hd)).astype(np.float32)).cuda(),
                "B_im": torch.from_numpy((rng.normal(size=(P, Hd)) / np.sqrt(2 * Hd)).astype(np.float32)).cuda(),
                "C_re": torch.from_numpy((rng.normal(size=(Hd, P)) / np.sqrt(P)).astype(np.float32)).cuda(),
                "C_im": torch.from_numpy((rng.normal(size=(Hd, P)) / np.sqrt(P)).astype(np.float32)).cuda(),
                "D": torch.from_numpy(rng.normal(size=Hd).astype(np.float32)).cuda()}
    layers = [mk(), mk()]
    def fn(u):
        for p in layers:
            u = S.lru_forward(p, u)
        return u
    u0 = torch.randn(B, T, Hd, device="cuda"); u1 = torch.randn(B, T, Hd, device="cuda")
    pg = A.PassGraph(fn, u0)
    for u in (u0, u1, u0):
        ref = fn(u)
        out = pg.run(u)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)

