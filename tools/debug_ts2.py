"""TS-variant GEMM race hunt: A encodes (row, col), W = I, so every wrong output names the element that was actually multiplied."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.ops as ops
M, N, K = 300000, 128, 128
m_idx = torch.arange(M, device="cuda") % 8192
a = (m_idx[:, None] * 128 + torch.arange(K, device="cuda")[None, :]).float()
w = torch.eye(N, K, device="cuda")
mode = os.environ.get("MODE", "tc3")
for trial in range(2):
    out = ops.linear(a, w, None, mode=mode)
    torch.cuda.synchronize()
    bad = (out != a)
    bad_rows = bad.any(dim=1).nonzero().flatten()
    tiles = torch.unique(bad_rows // 128)
    print("trial", trial, "bad rows", bad_rows.numel(), "bad tiles", tiles.numel(), tiles[:12].tolist(), flush=True)
    for t0 in tiles[:4].tolist():
        rows = bad_rows[(bad_rows // 128) == t0]
        rin = (rows % 128).tolist()
        print("  tile", t0, "worker", t0 % 148, "nth tile of worker", t0 // 148, "rows-in-tile", rin[:40], "n", len(rin))
        r = rows[0].item()
        cols = bad[r].nonzero().flatten().tolist()
        vals = out[r, cols[:4]].tolist()
        src = [(int(v) // 128, int(v) % 128) for v in vals]
        print("    row", r, "(mod 8192 = %d)" % (r % 8192), "bad cols", cols[:6], "..", cols[-2:], "ncols", len(cols), "got (row',col')", src,
              "delta rows", [s[0] - r % 8192 for s in src])
