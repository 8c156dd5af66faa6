"""TS-variant GLU epilogue race hunt: gate forced to 1 (bias +30, zero gate weights), value weights select columns, A encodes (row, col)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.ops as ops
M, K = 300000, 128
m_idx = torch.arange(M, device="cuda") % 1024
a = (m_idx[:, None] * 128 + torch.arange(K, device="cuda")[None, :]).float()
use_r = os.environ.get("RES", "1") == "1"
R = torch.randint(0, 512, (M, 64), device="cuda").float() if use_r else None
for sel in (0, 1):
    Wv = torch.zeros(64, K, device="cuda"); Wv[torch.arange(64), torch.arange(64) + 64 * sel] = 1
    W = torch.cat([Wv, torch.zeros(64, K, device="cuda")]); bias = torch.cat([torch.zeros(64), 30 * torch.ones(64)]).cuda()
    ref = a[:, 64 * sel:64 * sel + 64] + (R if use_r else 0)
    for trial in range(2):
        out = ops.linear(a, W, bias, epilogue="glu_residual", residual=R, mode="tc3")
        torch.cuda.synchronize()
        bad = (out != ref)
        bad_rows = bad.any(dim=1).nonzero().flatten()
        tiles = torch.unique(bad_rows // 128)
        print("sel", sel, "trial", trial, "bad rows", bad_rows.numel(), "bad tiles", tiles.numel(), tiles[:12].tolist(), flush=True)
        for t0 in tiles[:3].tolist():
            rows = bad_rows[(bad_rows // 128) == t0]
            rin = (rows % 128).tolist()
            print("  tile", t0, "worker", t0 % 148, "nth", t0 // 148, "rows-in-tile", rin[:48], "n", len(rin))
            r = rows[0].item()
            cols = bad[r].nonzero().flatten().tolist()
            got = (out[r, cols[:4]] - (R[r, cols[:4]] if use_r else 0)).tolist()
            src = [(int(v) // 128, int(v) % 128) for v in got]
            print("    row", r, "(mod 1024 = %d)" % (r % 1024), "bad cols", cols[:6], "..", cols[-2:], "ncols", len(cols), "value decodes to (row',col')", src,
                  "raw", [out[r, c].item() for c in cols[:3]], "expected", [ref[r, c].item() for c in cols[:3]])
