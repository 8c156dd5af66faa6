"""Error study of the operand splits for fp32-parity GEMMs on the tensor cores: 3xTF32 against the scaled fp16 split
(kind::f16 runs at twice the kind::tf32 rate).  Emulation in NumPy: operands are split exactly as the kernels do, every product is exact in
fp32 for both schemes (11-bit x 11-bit significands), accumulation in float64 here to isolate the REPRESENTATION error of the split (the tensor
core's own fp32 accumulation, truncating, is common to both schemes).  Usage: python tools/split_error_study.py"""
import json
import numpy as np


def tf32_round(x):
    b = x.astype(np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(a):                       # converter: hi = rna(a), lo = a - hi truncated by the tensor core; weights: lo = rna(w - hi)
    hi = tf32_round(a); lo = tf32_trunc(a - hi)
    return hi, lo


def split_f16(a, scale):                 # hi = fp16(a S), lo = fp16(a S - hi)
    s = (a * np.float32(scale)).astype(np.float32)
    hi = s.astype(np.float16)
    lo = (s - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def pow2_scale(maxabs, target):
    return 2.0 ** np.floor(np.log2(target / maxabs))


def study(M=2048, N=256, K=128, a_scale=1.0, w_bound=None, seed=0, sa=16.0):
    rng = np.random.default_rng(seed)
    A = (rng.normal(size=(M, K)) * a_scale).astype(np.float32)
    wb = w_bound if w_bound else 1.0 / np.sqrt(K)
    W = rng.uniform(-wb, wb, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T          # sum |a||w|: what an fp32 SGEMM's error is relative to
    out = {}
    ah, al = split_tf32(A); wh, wl = split_tf32(W)
    r = ah.astype(np.float64) @ wh.astype(np.float64).T + ah.astype(np.float64) @ wl.astype(np.float64).T + al.astype(np.float64) @ wh.astype(np.float64).T
    out["3xTF32"] = dict(max_rel_to_sumabs=float(np.max(np.abs(r - ref) / scale)), max_rel_to_outmax=float(np.max(np.abs(r - ref)) / np.abs(ref).max()))
    sw = pow2_scale(np.abs(W).max(), 2.0 ** 14)
    ah, al = split_f16(A, sa); wh, wl = split_f16(W, sw)
    r = (ah.astype(np.float64) @ wh.astype(np.float64).T + ah.astype(np.float64) @ wl.astype(np.float64).T + al.astype(np.float64) @ wh.astype(np.float64).T) / (sa * sw)
    out["f16x3 (S_a=%g, S_w=2^%d)" % (sa, int(np.log2(sw)))] = dict(max_rel_to_sumabs=float(np.max(np.abs(r - ref) / scale)),
                                                                      max_rel_to_outmax=float(np.max(np.abs(r - ref)) / np.abs(ref).max()),
                                                                      overflow=bool(np.isinf(ah).any()))
    ah, al = split_f16(A, 1.0); wh, wl = split_f16(W, 1.0)
    r = ah.astype(np.float64) @ wh.astype(np.float64).T + ah.astype(np.float64) @ wl.astype(np.float64).T + al.astype(np.float64) @ wh.astype(np.float64).T
    out["f16x3 unscaled"] = dict(max_rel_to_sumabs=float(np.max(np.abs(r - ref) / scale)), max_rel_to_outmax=float(np.max(np.abs(r - ref)) / np.abs(ref).max()))
    r32 = A @ W.T                                                                    # fp32 SGEMM (what the reference runs)
    out["fp32 sgemm (numpy)"] = dict(max_rel_to_sumabs=float(np.max(np.abs(r32 - ref) / scale)), max_rel_to_outmax=float(np.max(np.abs(r32 - ref)) / np.abs(ref).max()))
    return out


if __name__ == "__main__":
    res = {}
    for name, kw in [("unit-scale activations, K=128", dict()), ("activations x 1e-3", dict(a_scale=1e-3)), ("activations x 100", dict(a_scale=100.0)),
                     ("K=512, N=1024", dict(K=512, N=1024, M=512)), ("tiny weights 1e-3", dict(w_bound=1e-3))]:
        res[name] = study(**kw)
        print(name)
        for k, v in res[name].items():
            print("   %-28s %s" % (k, json.dumps(v)))
    json.dump(res, open("profiles/r2_split_error_study.json", "w"), indent=1)
