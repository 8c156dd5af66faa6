"""Per-kernel timing of one LRU layer call at the BASELINE-C3 shape (ListOps-like: T 2048, d_model 128, P 256 complex states), batch per GPU 128.
Usage: python tools/c3_probe.py [batch]"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigb200.ssm as S, eigb200.ops as ops
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T, Hd, P = 2048, 128, 256
rng = np.random.default_rng(0)
lam = np.sqrt(rng.uniform(0.9 ** 2, 0.99 ** 2, P))
prm = dict(nu_log=np.log(-np.log(lam)).astype(np.float32), theta_log=np.log(6.28 * rng.uniform(size=P)).astype(np.float32),
           gamma_log=np.log(np.sqrt(1 - lam ** 2)).astype(np.float32),
           B_re=(rng.normal(size=(P, Hd)) / np.sqrt(2 * Hd)).astype(np.float32), B_im=(rng.normal(size=(P, Hd)) / np.sqrt(2 * Hd)).astype(np.float32),
           C_re=(rng.normal(size=(Hd, P)) / np.sqrt(P)).astype(np.float32), C_im=(rng.normal(size=(Hd, P)) / np.sqrt(P)).astype(np.float32),
           D=rng.normal(size=Hd).astype(np.float32))
prm = {k: torch.from_numpy(v).cuda() for k, v in prm.items()}
u = torch.randn(Bsz, T, Hd, device="cuda")
for _ in range(2): S.lru_forward(prm, u)
torch.cuda.synchronize()
ops.PROFILE = []
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); y = S.lru_forward(prm, u); e1.record(); torch.cuda.synchronize()
per = {}
for name, s0, s1 in ops.PROFILE: per.setdefault(name, []).append(s0.elapsed_time(s1))
ops.PROFILE = None
tot = e0.elapsed_time(e1)
print(json.dumps({"batch": Bsz, "ms_per_layer": tot, "state_updates_per_s": Bsz * T * P / tot * 1e3, "kernels": {k: [len(v), round(sum(v), 3)] for k, v in per.items()}}))
