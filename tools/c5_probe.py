"""Per-kernel timing of one analysis pass at a BASELINE-C5-like shape (Mamba-2, d_model 512, 8 heads, d_state 16, T 1024): which kernels dominate
outside the C2 regime.  Usage: python tools/c5_probe.py [batch] [layers]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.analysis as A, eigb200.layers as Ly, eigb200.ops as ops
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(layer="mamba", version="mamba2", num_layers=nl, num_heads=8, input_dim=1, output_dim=50257, hidden_dim=512, state_dim=16,
           conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none", token_embedding=True, vocab_size=50257)
sd = Ly.init_mamba_state_dict(cfg, 1919)
model = Ly.MambaDev(cfg, sd, "cuda")
X = torch.randint(0, 50257, (Bsz, 1024)).cuda()
for _ in range(2): A.mamba_pass(model, X)
torch.cuda.synchronize()
ops.PROFILE = []
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); res = A.mamba_pass(model, X); e1.record(); torch.cuda.synchronize()
per = {}
for name, s0, s1 in ops.PROFILE: per.setdefault(name, []).append(s0.elapsed_time(s1))
ops.PROFILE = None
tot = e0.elapsed_time(e1)
print(json.dumps({"batch": Bsz, "layers": nl, "ms": tot, "eig_per_s": res.eig.numel() / tot * 1e3,
                  "kernels": {k: [len(v), round(sum(v), 3)] for k, v in per.items()}}))
