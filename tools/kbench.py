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


def case_normgate(B, T, D, H):
    x = torch.randn(B, T, D, device="cuda")
    W = torch.randn(H, D, device="cuda") * 0.3
    b = torch.zeros(H, device="cuda"); off = torch.linspace(4, 9, H, device="cuda")
    def fn():
        n = ops.normattn_gate(x, W, b, off, "softplus")
        ops.ratio_hist(n, 1)
    nbytes = x.numel() * 4 + B * T * H * (4 + 4 + 8)
    return fn, nbytes, B * (T - 1) * H


CASES = {
    "k1_c2": lambda: case_k1(4096, 512, 128, 1),
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
    a = ap.parse_args()
    peak, kind = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for name in a.cases:
        fn, nbytes, units = CASES[name]()
        med, best = time_fn(fn, a.iters, flush=flush if nbytes < (512 << 20) else None)
        rec = {"case": name, "ms_median": med, "ms_best": best, "alg_bytes": nbytes, "GBps": nbytes / med / 1e6,
               "frac_of_%s_peak" % kind: nbytes / med / 1e6 / peak, "units_per_s": units / med * 1e3}
        print(json.dumps(rec)); out.append(rec)
        del fn
        torch.cuda.empty_cache()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
