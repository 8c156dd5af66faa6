#!/usr/bin/env python
"""Per-kernel micro-benchmarks (CUDA events on the launching stream, inputs larger than L2 or L2 flushed).
Usage: python tools/kbench.py <case> [--iters N] [--json out]   cases: k1_c2 k1_c5 normgate_c5 ratio ...
Prints one JSON line per case with achieved algorithmic GB/s against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigb200.ops as ops  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_fn(fn, iters, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def case_k1(B, T, D, H, dtype=torch.float32, want_lam=True):
    x = torch.randn(B, T, D, device="cuda").to(dtype)
    W = torch.randn(H, D, device="cuda") * 0.3
    dtb = torch.full((H,), -1.0, device="cuda"); Al = torch.zeros(H, device="cuda")
    counts = ops.new_counts(B, H, "cuda")
    lam = torch.empty(B, T, H, device="cuda")
    def fn():
        counts.zero_()
        ops.mamba2_eig(x, W, dtb, Al, counts=counts, want_lam=want_lam, lam_out=lam)
    nbytes = x.numel() * x.element_size() + (B * T * H * 4 if want_lam else 0)
    return fn, nbytes, B * T * H


def case_k1_stats(B, T, D, H):
    x = torch.randn(B, T, D, device="cuda")
    W = torch.randn(H, D, device="cuda") * 0.3
    dtb = torch.full((H,), -1.0, device="cuda"); Al = torch.zeros(H, device="cuda")
    counts = ops.new_counts(B, H, "cuda")
    lam = torch.empty(B, T, H, device="cuda"); st = torch.empty(B, T, 2, device="cuda")
    def fn():
        ops.mamba2_eig(x, W, dtb, Al, counts=counts, lam_out=lam, rowstats_out=st)
    return fn, x.numel() * 4 + B * T * H * 4 + B * T * 8, B * T * H


def case_embed(B, T, D, V, stats=True):
    ids = torch.randint(0, V, (B, T), device="cuda")
    word = torch.randn(V, D, device="cuda"); pos = torch.randn(T, D, device="cuda")
    st = torch.empty(B, T, 2, device="cuda") if stats else None
    def fn():
        ops.embedding(ids, word, pos, rowstats_out=st)
    return fn, B * T * (8 + D * 4 + (8 if stats else 0)), B * T


def case_diag(B, T, P):
    """C3 (LRU / S5 on ListOps-shaped input): h_t = lam h_{t-1} + Bu_t over complex64, 16 B per state update."""
    lam = torch.polar(torch.empty(P, device="cuda").uniform_(0.9, 0.999), torch.empty(P, device="cuda").uniform_(0, 6.28))
    Bu = torch.view_as_complex(torch.randn(B, T, P, 2, device="cuda"))
    def fn():
        ops.diag_scan(lam, Bu)
    return fn, B * T * P * 16, B * T * P


def case_k3(nmat, N):
    """C4 (S4 DPLR, N = 64): discretise + dense nonsymmetric eigenvalues, one warp per matrix.  Reported as matrices/s (latency / FP64 bound)."""
    import numpy as np
    rng = np.random.default_rng(0)
    Lam = torch.from_numpy((-0.5 + 1j * np.pi * np.arange(N))[None].repeat(nmat, 0).astype(np.complex64)).cuda()
    Lam = Lam + torch.view_as_complex(torch.randn(nmat, N, 2, device="cuda") * 0.05)
    Pv = torch.view_as_complex(torch.randn(nmat, N, 2, device="cuda") * 0.5)
    step = torch.exp(torch.empty(nmat, device="cuda").uniform_(-6.9, -2.3))
    def fn():
        Ab = ops.dplr_abar(Lam, Pv, Pv, step)
        ops.eigvals_c64(Ab)
    return fn, nmat * (4 * N * 8 + N * 8), nmat


def case_s4_conv(B, T, H):
    """S4 CNN-mode layer call: causal convolution with the DPLR kernel (direct form, B H T^2 / 2 FMA)."""
    u = torch.randn(B, T, H, device="cuda"); Kt = torch.randn(T, H, device="cuda") * 0.05; D = torch.ones(H, device="cuda")
    def fn():
        ops.s4_causal_conv(u, Kt, D)
    return fn, 2 * B * T * H * 4, B * H * T * (T + 1) // 2


def case_s4_kernel(H, N, L):
    Lam = torch.complex(-0.5 * torch.ones(H, N), 3.14159 * torch.arange(N).float().repeat(H, 1)).cuda()
    P = torch.view_as_complex(torch.randn(H, N, 2, device="cuda") * 0.5); Bv = torch.view_as_complex(torch.randn(H, N, 2, device="cuda"))
    Cv = torch.view_as_complex(torch.randn(H, N, 2, device="cuda") * 0.7); step = torch.exp(torch.empty(H, device="cuda").uniform_(-6.9, -2.3))
    def fn():
        ops.s4_kernel(Lam, P, P, Bv, Cv, step, L)
    return fn, H * L * 4, H


def case_normgate(B, T, D, H):
    x = torch.randn(B, T, D, device="cuda")
    W = torch.randn(H, D, device="cuda") * 0.3
    b = torch.zeros(H, device="cuda"); off = torch.linspace(4, 9, H, device="cuda")
    def fn():
        n = ops.normattn_gate(x, W, b, off, "softplus")
        ops.ratio_hist(n, 1)
    nbytes = x.numel() * 4 + B * T * H * (4 + 4 + 8)
    return fn, nbytes, B * (T - 1) * H


def case_ssd(B, T=512, H=1, P=128, G=1, N=16):
    Cn = H * P + 2 * G * N
    ldz = (Cn + H + 3) // 4 * 4
    z = torch.randn(B * T, ldz, device="cuda")
    cw = torch.randn(Cn, 4, device="cuda") * 0.3; cb = torch.zeros(Cn, device="cuda")
    dtb = torch.full((H,), -1.0, device="cuda"); Al = torch.zeros(H, device="cuda"); Dv = torch.ones(H, device="cuda")
    y = torch.empty(B, T, H * P, device="cuda")
    def fn():
        ops.mamba_conv_ssd(z, ldz, cw, cb, dtb, Al, Dv, B, T, H, P, G, N, out=y)
    return fn, B * T * (Cn + H + H * P) * 4, B * T * H


def case_front(B, T=512, D=128, P=128, N=16):
    """LayerNorm + in_proj + conv + SiLU + SSD scan as one kernel (k7_front_fused.cu): reads x once, writes y once."""
    ops.set_gemm_precision("f16x3")
    x = torch.randn(B, T, D, device="cuda")
    stats = ops.rowstats(x)
    w = torch.randn(P + 2 * N + 1, D, device="cuda") / D ** 0.5
    ws = ops.linear_prepare(w, None, "none", torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"))
    Cn = P + 2 * N
    cw = torch.randn(Cn, 4, device="cuda") * 0.3; cb = torch.zeros(Cn, device="cuda")
    dtb = torch.full((1,), -1.0, device="cuda"); Al = torch.zeros(1, device="cuda"); Dv = torch.ones(1, device="cuda")
    y = torch.empty(B, T, P, device="cuda")
    def fn():
        ops.mamba_front_fused(x, stats, ws, cw, cb, dtb, Al, Dv, P, N, out=y)
    return fn, B * T * (D + P + 2) * 4, B * T


def case_tail(M, D=128, K1=128):
    """out_proj + GELU -> GLU + residual + extractor partials as one kernel (k4_gemm_fused.cu)."""
    ops.set_gemm_precision("f16x3")
    y = torch.randn(M, K1, device="cuda"); r = torch.randn(M, D, device="cuda")
    w1 = torch.randn(D, K1, device="cuda") / K1 ** 0.5; b1 = torch.randn(D, device="cuda") * 0.3
    w2 = torch.randn(2 * D, D, device="cuda") / D ** 0.5; b2 = torch.randn(2 * D, device="cuda")
    wg = torch.randn(D, device="cuda") / D ** 0.5
    ws1 = ops.linear_prepare(w1, b1, "gelu"); ws2 = ops.linear_prepare(w2, b2, "glu_residual")
    out = torch.empty(M, D, device="cuda"); part = torch.empty(D // 16, 3, M, device="cuda")
    def fn():
        ops.out_glu_fused(y, ws1, b1, ws2, b2, r, wg, part, out=out)
    return fn, M * 3 * D * 4, 2.0 * M * D * 3 * D


def case_linear(M, N, K, epi, mode):
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    nout = N // 2 if epi == "glu_residual" else N
    ldc = (nout + 7) // 8 * 8
    r = torch.randn(M, nout, device="cuda") if ("residual" in epi and not epi.endswith("_nores")) else None
    epi = epi.replace("_nores", "")
    out = torch.empty(M, ldc, device="cuda")
    ws, _ = ops.linear_workspace(N, K, "cuda", M)
    def fn():
        ops.linear(a, w, b, epilogue=epi, residual=r, mode=mode, out=out, workspace=ws)
    nbytes = M * (K + nout + (nout if r is not None else 0)) * 4
    return fn, nbytes, 2.0 * M * N * K


def case_linear_ln(M, N, K, ldc):
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    stats = torch.stack([a.mean(-1), torch.rsqrt(a.var(-1, unbiased=False) + 1e-5)], -1).contiguous()
    g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
    out = torch.empty(M, ldc, device="cuda")
    ws, _ = ops.linear_workspace(N, K, "cuda", M)
    def fn():
        ops.linear_ln(a, stats, g, be, w, b, out=out, workspace=ws)
    return fn, M * (K + N) * 4 + M * 8, 2.0 * M * N * K


def case_ln(M, D):
    x = torch.randn(M, D, device="cuda"); w = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
    def fn():
        ops.layernorm(x, w, b)
    return fn, 2 * M * D * 4, M


def case_linattn(B, T, H, form, conv=False):
    """Norm-attention layer core at the C5 head shape (d = dv = 64): projection buffer [v | q | k | n] -> gated causal linear attention."""
    if form:
        os.environ["EIGB200_LINATTN_FORM"] = form
    else:
        os.environ.pop("EIGB200_LINATTN_FORM", None)
    d = dv = 64; D = H * d
    ld = 3 * D + 8
    buf = torch.randn(B * T, ld, device="cuda")
    gate = torch.rand(B, T, H, device="cuda")
    cw = torch.randn(3 * D, 4, device="cuda") * 0.5; cb = torch.randn(3 * D, device="cuda") * 0.2
    if conv:
        def fn():
            ops.linattn_forward_conv(buf, ld, D, 2 * D, 0, B, T, H, d, dv, cw, cb, D, 2 * D, 0, gate=gate, phi_elu=True, normalise=False, kscale=0.125)
    else:
        def fn():
            ops.linattn_forward(buf, ld, D, 2 * D, 0, B, T, H, d, dv, gate=gate, phi_elu=True, normalise=False, kscale=0.125)
    return fn, B * T * 4 * D * 4, B * T * H


def case_conv(B, T, Cn):
    ld = Cn + 8
    buf = torch.randn(B * T, ld, device="cuda"); out = torch.empty_like(buf)
    cw = torch.randn(Cn, 4, device="cuda") * 0.5; cb = torch.randn(Cn, device="cuda") * 0.2
    def fn():
        ops.conv_silu(buf, ld, cw, cb, B, T, Cn, out=out, ldo=ld)
    return fn, 2 * B * T * Cn * 4, B * T


M_C2 = 4096 * 512
CASES = {
    "linattn_c5": lambda: case_linattn(1024, 1024, 8, None),
    "linattn_c5_col": lambda: case_linattn(1024, 1024, 8, "col"),
    "linattn_c5_conv": lambda: case_linattn(1024, 1024, 8, None, conv=True),
    "linattn_small": lambda: case_linattn(128, 1024, 8, None),
    "conv_c5": lambda: case_conv(1024, 1024, 1536),
    "ssd_c2": lambda: case_ssd(4096),
    "ssd_small": lambda: case_ssd(512),
    "ssd_c5": lambda: case_ssd(1024, T=1024, H=8, P=64),
    "ssd_c2_tc": lambda: (os.environ.__setitem__("EIGB200_SSD_FORM", "tc"), case_ssd(4096))[1],
    "ssd_small_tc": lambda: (os.environ.__setitem__("EIGB200_SSD_FORM", "tc"), case_ssd(512))[1],
    "ssd_c2_scan": lambda: (os.environ.__setitem__("EIGB200_SSD_FORM", "scan"), case_ssd(4096))[1],
    "ssd_small_scan": lambda: (os.environ.__setitem__("EIGB200_SSD_FORM", "scan"), case_ssd(512))[1],
    "tail_c2": lambda: case_tail(M_C2),
    "tail_small": lambda: case_tail(592 * 512),
    "front_c2": lambda: case_front(4096),
    "front_small": lambda: case_front(592),
    "ln_c2": lambda: case_ln(M_C2, 128),
    "lin_in_tc3": lambda: case_linear(M_C2, 161, 128, "none", "tc3"),
    "lin_in_ln_tc3": lambda: case_linear_ln(M_C2, 161, 128, 168),
    "lin_out_tc3": lambda: case_linear(M_C2, 128, 128, "gelu", "tc3"),
    "lin_glu_tc3": lambda: case_linear(M_C2, 256, 128, "glu_residual", "tc3"),
    "lin_out_tc1": lambda: case_linear(M_C2, 128, 128, "gelu", "tc1"),
    "lin_glu_tc1": lambda: case_linear(M_C2, 256, 128, "glu_residual", "tc1"),
    "lin_in_tc1": lambda: case_linear(M_C2, 161, 128, "none", "tc1"),
    "lin_in_f16": lambda: case_linear(M_C2, 161, 128, "none", "f16x3"),
    "lin_out_f16": lambda: case_linear(M_C2, 128, 128, "gelu", "f16x3"),
    "lin_glu_f16": lambda: case_linear(M_C2, 256, 128, "glu_residual", "f16x3"),
    "lin_out_none_f16": lambda: case_linear(M_C2, 128, 128, "none", "f16x3"),
    "lin_glu_nores_f16": lambda: case_linear(M_C2, 256, 128, "glu_residual_nores", "f16x3"),
    "lin_c5_qkv_f16": lambda: case_linear(262144, 1544, 512, "none", "f16x3"),
    "lin_c5_out_f16": lambda: case_linear(262144, 512, 512, "residual", "f16x3"),
    "lin_c3_out": lambda: case_linear(262144, 128, 512, "residual", "auto"),
    "lin_c5_glu_tc3": lambda: case_linear(262144, 1024, 512, "glu_residual", "tc3"),
    "lin_c5_glu_f16": lambda: case_linear(262144, 1024, 512, "glu_residual", "f16x3"),
    "lin_c5_in_tc3": lambda: case_linear(262144, 552, 512, "none", "tc3"),
    "lin_c5_in_f16": lambda: case_linear(262144, 552, 512, "none", "f16x3"),
    "lin_out_simt": lambda: case_linear(M_C2, 128, 128, "gelu", "simt"),
    "lin_small_tc3": lambda: case_linear(65536, 128, 128, "gelu", "tc3"),
    "lin_glu_mid_tc3": lambda: case_linear(524288, 256, 128, "glu_residual", "tc3"),
    "lin_glu_nores_tc3": lambda: case_linear(M_C2, 256, 128, "glu_residual_nores", "tc3"),
    "lin_out_none_tc3": lambda: case_linear(M_C2, 128, 128, "none", "tc3"),
    "lin_n96_none_tc3": lambda: case_linear(M_C2, 96, 128, "none", "tc3"),
    "lin_n96_gelu_tc3": lambda: case_linear(M_C2, 96, 128, "gelu", "tc3"),
    "lin_n64_none_tc3": lambda: case_linear(M_C2, 64, 128, "none", "tc3"),
    "diag_c3": lambda: case_diag(128, 2048, 256),
    "k3_c4": lambda: case_k3(3072, 64),
    "s4_conv_c4": lambda: case_s4_conv(64, 2048, 128),
    "s4_kernel_c4": lambda: case_s4_kernel(128, 64, 2048),
    "k1_c2": lambda: case_k1(4096, 512, 128, 1),
    "k1_c2_stats": lambda: case_k1_stats(4096, 512, 128, 1),
    "emb_c2": lambda: case_embed(4096, 512, 128, 8192),
    "emb_c2_nostats": lambda: case_embed(4096, 512, 128, 8192, stats=False),
    "k1_c2_nolam": lambda: case_k1(4096, 512, 128, 1, want_lam=False),
    "k1_c2_bf16": lambda: case_k1(4096, 512, 128, 1, torch.bfloat16),
    "k1_c5": lambda: case_k1(1024, 1024, 512, 8),
    "k1_small": lambda: case_k1(64, 512, 128, 1),
    "normgate_c5": lambda: case_normgate(1024, 1024, 512, 8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=["k1_c2"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--ssd-variants", default=None, help="comma list of EIGB200_SSD_VARIANT values to sweep for ssd_* cases")
    a = ap.parse_args()
    peak, kind = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    sweep = []
    for name in a.cases:
        if a.ssd_variants and name.startswith("ssd_"):
            sweep += [(name, v) for v in a.ssd_variants.split(",")]
        else:
            sweep.append((name, None))
    for name, variant in sweep:
        if variant is not None:
            os.environ["EIGB200_SSD_VARIANT"] = variant
        fn, nbytes, units = CASES[name]()
        med, best = time_fn(fn, a.iters, flush=flush if nbytes < (512 << 20) else None)
        rec = {"case": name if variant is None else "%s[v%s]" % (name, variant), "ms_median": med, "ms_best": best, "alg_bytes": nbytes, "GBps": nbytes / med / 1e6,
               "frac_of_%s_peak" % kind: nbytes / med / 1e6 / peak, "units_per_s": units / med * 1e3}
        if name.startswith("lin_"):
            rec["TFLOPs_fp32_equiv"] = units / med / 1e9
        print(json.dumps(rec)); out.append(rec)
        del fn
        torch.cuda.empty_cache()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
