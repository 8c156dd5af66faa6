"""Per-kernel timing of one analysis pass at the BASELINE-C5 normalised-attention shape (d_model 512, 8 heads, d_qk 512, conv 4, T 1024).
Usage: python tools/c5_probe_attn.py [batch] [layers] [attention_fn]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.analysis as A, eigb200.layers as Ly, eigb200.ops as ops
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fn = sys.argv[3] if len(sys.argv) > 3 else "norm-attention"
cfg = dict(layer="transformer", input_dim=1, output_dim=50257, num_layers=nl, hidden_dim=512, embedding=True, vocab_size=50257, max_pos_embed=1024,
           pooling="none", dual=False, classifier=False, mixer_dim=2048, norm="layer", dropout=0.0, state_dim=512, num_heads=8, att_dropout=0.0,
           use_flash=False, attention_fn=fn, mixer="glu", mode="attention", norm_fn="softplus", approx_fn="elu", scale_B=False, offset=True,
           offset_init="exp", learn_A=False, dim_conv=4)
sd = Ly.init_transformer_state_dict(cfg, 1919)
model = Ly.TransformerDev(cfg, sd, "cuda")
X = torch.randint(0, 50257, (Bsz, 1024)).cuda()
for _ in range(2): A.transformer_pass(model, X, cfg)
torch.cuda.synchronize()
ops.PROFILE = []
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); res = A.transformer_pass(model, X, cfg); e1.record(); torch.cuda.synchronize()
per = {}
for name, s0, s1 in ops.PROFILE: per.setdefault(name, []).append(s0.elapsed_time(s1))
ops.PROFILE = None
tot = e0.elapsed_time(e1)
print(json.dumps({"attention_fn": fn, "batch": Bsz, "layers": nl, "ms": tot, "eig_per_s": res.eig.numel() / tot * 1e3,
                  "kernels": {k: [len(v), round(sum(v), 3)] for k, v in per.items()}}))
