#!/usr/bin/env python
"""Wall time of the WHOLE reference-facing call eval_eig(args, conf_args, wandb_config, data_config, loader, path_file, perf) at the BASELINE C2
model (4-layer Mamba-2, d_model 128, T 512) for a batch of B sequences: checkpoint read, init model, two device passes (trained + init), the
6-tuple on the host, 10 .npy + yaml + percentage_file.txt.  usage: python tools/eval_eig_probe.py [B] [repeats]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import eigb200.analysis as A
import eigb200.layers as Ly

C2 = dict(version="mamba2", num_layers=4, num_heads=1, input_dim=1, output_dim=8192, hidden_dim=128, state_dim=16,
          conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none",
          token_embedding=True, vocab_size=8192)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 512
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    sd = Ly.init_mamba_state_dict(dict(C2, layer="mamba"), 7)
    ckpt = os.path.join(tmp, "model-perf0.900.pth")
    torch.save(sd, ckpt)
    X = torch.randint(0, C2["vocab_size"], (B, T), generator=torch.Generator().manual_seed(1))
    loader = [(X, torch.zeros(B), None)]
    times = []
    for r in range(reps):
        args = {"seed": 1919, "model": dict(C2, layer="mamba", seq_len=T), "train": {"lr": 0.01}, "dataset": {"name": "MQAR"}}
        conf = {"batch_size": B, "save_path": os.path.join(tmp, "out%d" % r) + "/"}
        os.makedirs(conf["save_path"], exist_ok=True)                # like the reference, eval_eig creates only the run's own sub-directory
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = A.eval_eig(args, conf, None, args["dataset"], loader, ckpt, 0.9)
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    model = Ly.MambaDev(dict(C2, layer="mamba"), sd, "cuda")
    Xd = X.cuda()
    A.mamba_pass(model, Xd); torch.cuda.synchronize()
    t0 = time.perf_counter(); res = A.mamba_pass(model, Xd); torch.cuda.synchronize(); t_pass = time.perf_counter() - t0
    t0 = time.perf_counter(); e = res.eig_host(); c = res.counts.cpu(); t_d2h = time.perf_counter() - t0
    n_eig = 2 * B * T * 1 * 4                                       # trained + init pass
    print(json.dumps({"probe": "eval_eig whole call", "B": B, "T": T, "wall_s": times, "best_s": min(times), "eigenvalues": n_eig,
                      "eig_per_s_whole_call": n_eig / min(times), "one_device_pass_s": t_pass, "d2h_pageable_s": t_d2h,
                      "eig_shape": list(out[0].shape)}))
