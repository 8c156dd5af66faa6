import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigb200.ops as ops, oracle as O
def run(B, T, H, G, kconv, P=128, N=16, seed=None):
    rng = np.random.default_rng(kconv if seed is None else seed)
    C_ = H * P + 2 * G * N
    ldz = (C_ + H + 3) // 4 * 4
    z = rng.normal(size=(B, T, ldz)).astype(np.float32)
    cw = rng.normal(size=(C_, kconv)).astype(np.float32) * 0.5; cb = rng.normal(size=C_).astype(np.float32) * 0.1
    dtb = rng.normal(-1, 1, H).astype(np.float32); Al = np.log(rng.uniform(1, 16, H)).astype(np.float32); Dv = rng.normal(size=H).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    out = {}
    for form in ("tc", "scan"):
        os.environ["EIGB200_SSD_FORM"] = form
        out[form] = ops.mamba_conv_ssd(dev(z), ldz, dev(cw), dev(cb), dev(dtb), dev(Al), dev(Dv), B, T, H, P, G, N).cpu().numpy().reshape(B, T, H, P)
    z64 = z.astype(np.float64)
    xBC = O.causal_depthwise_conv_silu(z64[..., :C_], cw.astype(np.float64), cb.astype(np.float64))
    dt = O.softplus(z64[..., C_:C_ + H] + dtb)
    ref = O.ssd_scan_sequential(xBC[..., :H * P].reshape(B, T, H, P), dt, -np.exp(Al.astype(np.float64)),
                                xBC[..., H * P:H * P + G * N].reshape(B, T, G, N), xBC[..., H * P + G * N:].reshape(B, T, G, N), Dv.astype(np.float64))
    scale = np.abs(ref).max(axis=1, keepdims=True)
    for form in ("tc", "scan"):
        r = np.abs(out[form] - ref) / (1e-5 * scale + 1e-6)
        idx = np.unravel_index(np.argmax(r), r.shape)
        bad = np.argwhere(r > 1)
        print("B%d T%d H%d k%d %-4s max ratio %.3f at (b,t,h,p)=%s nbad %d  bad t hist %s" % (B, T, H, kconv, form, r.max(), idx, len(bad),
              np.bincount(bad[:, 1] // 16, minlength=(T + 15) // 16).tolist() if len(bad) else []))
run(3, 256, 2, 1, 4); run(3, 75, 2, 1, 4); run(2, 512, 1, 1, 4, seed=7); run(2, 64, 1, 1, 4, seed=9)
