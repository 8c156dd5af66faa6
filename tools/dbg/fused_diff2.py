import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigb200.ops as ops
ops.set_gemm_precision("f16x3")
def dev(a): return torch.from_numpy(a).cuda()
def trial(M, K1, bias1, yscale, rscale, seed=0):
    D = 128
    rng = np.random.default_rng(seed)
    y = (rng.normal(size=(M, K1)) * yscale).astype(np.float32)
    w1 = (rng.normal(size=(D, K1)) / np.sqrt(K1)).astype(np.float32); b1 = (rng.normal(size=D) * 0.3).astype(np.float32) if bias1 else None
    w2 = (rng.normal(size=(2 * D, D)) / np.sqrt(D)).astype(np.float32); b2 = rng.normal(size=2 * D).astype(np.float32)
    r = (rng.normal(size=(M, D)) * rscale).astype(np.float32)
    ws1 = ops.linear_prepare(dev(w1), dev(b1) if bias1 else None, "gelu"); ws2 = ops.linear_prepare(dev(w2), dev(b2), "glu_residual")
    out, _ = ops.out_glu_fused(dev(y), ws1, dev(b1) if bias1 else None, ws2, dev(b2), dev(r))
    o = ops.linear(dev(y), dev(w1), dev(b1) if bias1 else None, epilogue="gelu", mode="f16x3")
    ref2 = ops.linear(o, dev(w2), dev(b2), epilogue="glu_residual", residual=dev(r), mode="f16x3")
    # GEMM2 alone on the same o: fused-kernel GEMM 2 vs TS kernel -- emulate by feeding fused kernel? not possible; report diff stats
    d = (out - ref2).abs()
    print("M %d K1 %d bias1 %s yscale %g rscale %g: maxdiff %.3e n %d / %d" % (M, K1, bias1, yscale, rscale, d.max().item(), (d > 0).sum().item(), d.numel()))
for args in [(4096, 128, True, 1.0, 1.5), (4096, 128, False, 1.0, 1.5), (4096, 128, False, 0.2, 1.5), (4096, 128, False, 0.2, 0.3), (1536, 128, False, 0.05, 1.0), (4096, 128, True, 0.05, 1.0)]:
    trial(*args)
