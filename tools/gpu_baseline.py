"""On-GPU comparator for bench.py / kbench.py: the reference's Mamba analysis pass (analysis/eval_eig.py:575-618 with MambaBlock.forward,
models/mamba.py:111-154, :328-340) the way it runs on a GPU with stock libraries -- eager PyTorch (cuBLAS SGEMM with allow_tf32 = False as in
the reference, cuDNN / ATen conv1d, elementwise ATen kernels) and, for the SSD scan that the reference takes from mamba_ssm's Triton kernel
(not installable here), `fla.ops.simple_gla.chunk_simple_gla(scale=1.0)` (Triton) when it imports, else a chunked SSD in eager torch.

This is a BASELINE, not the product and not the oracle: nothing under eigb200/ imports it, it shares no kernel with libeigb200.so, and its numbers
only say what the same pass costs without the hand-written sm_100a path.  It deliberately keeps the reference's structure: the extractor re-runs
the WHOLE in_proj on the block output (get_eig_mamba2, eval_eig.py:176-190), eigenvalues go to the host per layer and are concatenated there,
the threshold statistics are NumPy on the host.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn.functional as F

THRESHOLDS_RADIUS = (0.1, 0.5, 0.9, 1.0, 10.0, 100.0)


def _ssd_torch_chunked(x, dt, A, Bm, Cm, D, chunk=64):
    """Chunked SSD ("ssd_minimal" form of the Mamba-2 paper) in eager torch.  x (B,T,H,P), dt (B,T,H), A (H), Bm/Cm (B,T,G,N) with G = 1."""
    Bsz, T, H, P = x.shape
    N = Bm.shape[-1]
    pad = (-T) % chunk
    if pad:
        x = F.pad(x, (0, 0, 0, 0, 0, pad)); dt = F.pad(dt, (0, 0, 0, pad)); Bm = F.pad(Bm, (0, 0, 0, 0, 0, pad)); Cm = F.pad(Cm, (0, 0, 0, 0, 0, pad))
    nc = (T + pad) // chunk
    xd = (x * dt[..., None]).reshape(Bsz, nc, chunk, H, P)
    a = (dt * A).reshape(Bsz, nc, chunk, H).permute(0, 3, 1, 2)                    # (B,H,c,l) log decays
    Bc = Bm[:, :, 0].reshape(Bsz, nc, chunk, N); Cc = Cm[:, :, 0].reshape(Bsz, nc, chunk, N)
    acs = torch.cumsum(a, dim=-1)
    seg = acs[..., :, None] - acs[..., None, :]
    Lm = torch.exp(seg.masked_fill(~torch.tril(torch.ones(chunk, chunk, dtype=torch.bool, device=x.device)), float("-inf")))
    y_diag = torch.einsum("bcln,bcsn,bhcls,bcshp->bclhp", Cc, Bc, Lm, xd)
    decay_states = torch.exp(acs[..., -1:] - acs)
    states = torch.einsum("bcln,bhcl,bclhp->bchpn", Bc, decay_states, xd)
    chunk_decay = torch.exp(acs[..., -1])                                          # (B,H,c)
    prev = torch.zeros(Bsz, H, P, N, device=x.device, dtype=x.dtype)
    outs = []
    for c in range(nc):                                                            # inter-chunk recurrence
        outs.append(prev)
        prev = prev * chunk_decay[:, :, c, None, None] + states[:, c]
    st = torch.stack(outs, dim=1)                                                  # (B,c,H,P,N) state entering each chunk
    y_off = torch.einsum("bcln,bchpn,bhcl->bclhp", Cc, st, torch.exp(acs))
    y = (y_diag + y_off).reshape(Bsz, nc * chunk, H, P)[:, :T]
    return y + D[None, None, :, None] * x[:, :T]


def _ssd(x, dt, A, Bm, Cm, D, impl):
    if impl == "fla":
        from fla.ops.simple_gla import chunk_simple_gla
        H = x.shape[2]
        q = Cm.expand(-1, -1, H, -1).contiguous(); k = Bm.expand(-1, -1, H, -1).contiguous()
        y, _ = chunk_simple_gla(q=q, k=k, v=(x * dt[..., None]).contiguous(), g=(dt * A).contiguous(), scale=1.0)
        return y + D[None, None, :, None] * x
    return _ssd_torch_chunked(x, dt, A, Bm, Cm, D)


def pick_ssd_impl():
    try:
        from fla.ops.simple_gla import chunk_simple_gla  # noqa: F401
        return "fla"
    except Exception:
        return "torch"


@torch.no_grad()
def mamba_pass_eager(ids, sd, cfg, ssd_impl="fla"):
    """-> (eig (B,T,H,L) float32 numpy, percentage (7,B,H,L)).  sd: reference-format state dict of CUDA tensors."""
    D = cfg["hidden_dim"]; H = cfg["num_heads"]; N = cfg["state_dim"]; hd = D // H; di = D; G = 1
    x = F.embedding(ids, sd["encoder.word_embeddings.weight"])
    Bsz, T, _ = x.shape
    eig = None
    for i in range(cfg["num_layers"]):
        p = "blocks.%d." % i
        skip = x
        xn = F.layer_norm(x, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"])
        z = F.linear(xn, sd[p + "mamba.in_proj.weight"])
        xBC, dt = z[..., : di + 2 * G * N], z[..., di + 2 * G * N:]
        dt = F.softplus(dt + sd[p + "mamba.dt_bias"])
        w = sd[p + "mamba.conv1d.weight"]
        xBC = F.silu(F.conv1d(xBC.transpose(1, 2), w, sd[p + "mamba.conv1d.bias"], padding=w.shape[-1] - 1, groups=w.shape[0]).transpose(1, 2))[:, :T]
        xs, Bm, Cm = xBC[..., :di], xBC[..., di:di + G * N], xBC[..., di + G * N:]
        y = _ssd(xs.reshape(Bsz, T, H, hd), dt, -torch.exp(sd[p + "mamba.A_log"]), Bm.reshape(Bsz, T, G, N), Cm.reshape(Bsz, T, G, N),
                 sd[p + "mamba.D"], ssd_impl).reshape(Bsz, T, di)
        o = F.gelu(F.linear(y, sd[p + "mamba.out_proj.weight"]))
        g = F.linear(o, sd[p + "glu.linear.weight"], sd[p + "glu.linear.bias"])
        x = g[..., :D] * torch.sigmoid(g[..., D:]) + skip
        zz = F.linear(x, sd[p + "mamba.in_proj.weight"])                            # get_eig_mamba2: the whole in_proj again
        lam = torch.exp(F.softplus(zz[..., di + 2 * G * N:] + sd[p + "mamba.dt_bias"]) * -torch.exp(sd[p + "mamba.A_log"]))
        lam = np.expand_dims(lam.cpu().numpy(), axis=-1)
        eig = lam if eig is None else np.concatenate((eig, lam), axis=-1)
    rad = np.sqrt(np.power(eig.real, 2) + np.power(eig.imag, 2))
    edges = (0.0,) + THRESHOLDS_RADIUS
    pct = [((rad >= lo) & (rad <= hi)).sum(axis=1) * 100.0 / T for lo, hi in zip(edges[:-1], edges[1:])] + [(rad >= edges[-1]).sum(axis=1) * 100.0 / T]
    return eig, np.stack(pct)


def time_mamba_pass_eager(cfg, sd_cpu, X_dev, steps=3, warmup=2):
    """-> dict(eig_per_s, ms_per_pass, ssd_impl, sample) for one GPU.  Wall-clock around the whole call (it ends on the host by construction)."""
    torch.backends.cuda.matmul.allow_tf32 = False                                   # the reference's default: full fp32 GEMMs
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(X_dev.device) for k, v in sd_cpu.items()}
    impl = pick_ssd_impl()
    try:
        for _ in range(warmup):
            mamba_pass_eager(X_dev, sd, cfg, impl)
    except Exception:
        if impl != "fla":
            raise
        impl = "torch"
        for _ in range(warmup):
            mamba_pass_eager(X_dev, sd, cfg, impl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eig, _ = mamba_pass_eager(X_dev, sd, cfg, impl)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    return {"value": eig.size / sec, "unit": "eigenvalues/s", "ms_per_pass": sec * 1e3,
            "kind": "eager PyTorch on the same B200 (cuBLAS fp32 GEMMs, ATen conv / elementwise, SSD scan: %s), reference structure incl. per-layer "
                    "D2H of the eigenvalues and NumPy statistics" % ("fla.ops.simple_gla.chunk_simple_gla (Triton, scale=1.0)" if impl == "fla" else "chunked SSD in eager torch"),
            "sample": "%d sequences x T=%d x %d layers per pass, mean of %d passes after %d warm-ups" % (X_dev.shape[0], X_dev.shape[1], cfg["num_layers"], steps, warmup)}
