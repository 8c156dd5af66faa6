import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigb200.analysis as A, eigb200.layers as Ly, eigb200.ops as ops
ops.set_gemm_precision("f16x3")
cfg = dict(layer="mamba", version="mamba2", num_layers=2, num_heads=1, input_dim=1, output_dim=32, hidden_dim=128, state_dim=16, conv_dim=4,
           expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none", token_embedding=True, vocab_size=97)
sd = Ly.init_mamba_state_dict(cfg, 11)
model = Ly.MambaDev(cfg, sd, "cuda")
X = torch.randint(0, 97, (16, 96), generator=torch.Generator().manual_seed(4)).cuda()
def run(fuse, prep=True):
    for b in model.blocks: b.fuse_tail = fuse; b.prepare_weights = prep
    r = A.mamba_pass(model, X); torch.cuda.synchronize()
    return r.eig.clone(), r.x_last.clone()
e_f, x_f = run(True); e_f2, x_f2 = run(True)
e_u, x_u = run(False); e_p, x_p = run(False, False)
print("fused vs fused again   eig equal", torch.equal(e_f, e_f2), "x equal", torch.equal(x_f, x_f2))
print("fused vs unfused       eig maxdiff %.3e x maxdiff %.3e nx %d" % ((e_f - e_u).abs().max().item(), (x_f - x_u).abs().max().item(), (x_f != x_u).sum().item()))
print("unfused prep vs percall eig equal", torch.equal(e_u, e_p), "x maxdiff %.3e" % (x_u - x_p).abs().max().item())
# layer by layer
x = model.encoder(X)
for i, blk in enumerate(model.blocks):
    blk.fuse_tail = True; blk.prepare_weights = True
    st = ops.rowstats(x)
    part = torch.empty(8, 3, x.shape[0] * x.shape[1], device="cuda")
    a = blk(x, st, extract_partials=part).clone(); pa = part.clone()
    blk.fuse_tail = False
    b = blk(x, st, extract_partials=part).clone(); pb = part.clone()
    d = (a - b).abs()
    idx = (d > 0).nonzero()
    print("layer", i, "x maxdiff %.3e n %d" % (d.max().item(), idx.shape[0]), "partials maxdiff %.3e" % (pa - pb).abs().max().item(), idx[:5].tolist())
    x = b
