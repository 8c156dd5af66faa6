import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.ops as ops
torch.manual_seed(0)
M, N, K = 300000, 128, 128
a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
ref = (a.double() @ w.double().T)
R = torch.randn(M, N, device="cuda")
EPI = os.environ.get("EPI", "none")
if EPI == "residual":
    ref = ref + R.double()
for trial in range(3):
    out = ops.linear(a, w, None, mode=os.environ.get("MODE", "tc3"), epilogue=EPI, residual=R if EPI == "residual" else None)
    err = (out.double() - ref).abs()
    bad_rows = (err.max(dim=1).values > (1e-3 if os.environ.get("MODE", "tc3") == "tc3" else 5e-2)).nonzero().flatten()
    tiles = torch.unique(bad_rows // 128)
    print("trial", trial, "max err %.3g" % err.max().item(), "bad rows", bad_rows.numel(), "bad tiles", tiles.numel(), tiles[:20].tolist())
    if bad_rows.numel():
        t0 = tiles[0].item()
        e_t = err[t0 * 128:(t0 + 1) * 128]
        br = (e_t.max(dim=1).values > 1e-3).nonzero().flatten().tolist()
        bc = (e_t.max(dim=0).values > 1e-3).nonzero().flatten().tolist()
        print("  first bad tile", t0, "rows-in-tile", br, "cols", bc[:8], "...", bc[-4:], "ncols", len(bc))
        t1 = tiles[min(3, tiles.numel() - 1)].item()
        e_t = err[t1 * 128:(t1 + 1) * 128]
        print("  another bad tile", t1, "rows-in-tile", (e_t.max(dim=1).values > 1e-3).nonzero().flatten().tolist(), "ncols", int((e_t.max(dim=0).values > 1e-3).sum()))
        # does the wrong value equal the right GEMM result plus ANOTHER row's residual?
        r = bad_rows[0].item()
        g = (a[r].double() @ w.double().T)
        resid_used = out[r].double() - g
        dist = (R.double() - resid_used[None]).abs().max(dim=1).values
        jbest = dist.argmin().item()
        print("  row", r, "used the residual of row", jbest, "(dist %.3g)" % dist[jbest].item(), "delta rows", jbest - r)
        r = bad_rows[0].item()
        # which K chunk is wrong? recompute per-chunk partial products for that row
        row_out = out[r].double() - (R[r].double() if EPI == 'residual' else 0); parts = [(a[r, 32*c:32*c+32].double() @ w[:, 32*c:32*c+32].double().T) for c in range(4)]
        full = sum(parts)
        d = row_out - full
        # test hypotheses: a chunk missing / doubled / replaced by another row's chunk
        for c in range(4):
            print("  row", r, "tile", r // 128, "worker", (r // 128) % 148, "chunk", c, "resid if chunk missing %.3g  doubled %.3g" % ((d + parts[c]).abs().max().item(), (d - parts[c]).abs().max().item()))
        # is it another tile's data? find the row r2 with same in-tile offset whose product matches
        cand = torch.arange(r % 128, M, 128, device="cuda")
        for c in range(4):
            alt = a[cand, 32*c:32*c+32].double() @ w[:, 32*c:32*c+32].double().T          # (ntiles, N)
            res = (d[None] + parts[c][None] - alt).abs().max(dim=1).values
            j = res.argmin().item()
            print("  chunk", c, "best replacement tile", cand[j].item() // 128, "resid %.3g" % res[j].item())
