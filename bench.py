#!/usr/bin/env python
"""bench.py -- eigenvalues/sec of the eval_eig hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference algorithm on the host cores (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE config C2 -- Mamba-2 on MQAR-shaped synthetic tokens, T=512, d_model=128, 1 head, d_state=16,
conv 4, GLU, prenorm, 4 layers, vocab 8192, 4096 sequences PER GPU (weak scaling: the batch shards by sequence).
One step = one analysis pass over the batch: token embedding, then for each layer the block forward (LayerNorm, in_proj, causal
conv + SiLU, SSD scan, out_proj + GELU, GLU + residual) followed by the fused eigenvalue extractor + radius/phase bin counts on the
block's output (analysis/eval_eig.py:575-618).  One "eigenvalue" = one element of the returned `eig` array: B*T*H*L per step.

value   : whole-job eigenvalues/s with the token ids already resident in HBM (CUDA events, max over ranks).
e2e     : the same through the public API with HOST buffers: pinned token ids -> device, the pass, eigenvalue array + bin counts ->
          pinned host memory, inside the timed region, every step.  Two batches are in flight (two captured passes, uploads and downloads on
          their own streams), so the PCIe transfers of one step overlap the kernels of the next; `--no-graph` runs it strictly serially.
roofline: dominant kernel of the step, algorithmic bytes / its event-timed duration, against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "eigenvalues/sec in eval_eig"
UNIT = "eigenvalues/s"

C2 = dict(layer="mamba", version="mamba2", num_layers=4, num_heads=1, input_dim=1, output_dim=8192, hidden_dim=128, state_dim=16,
          conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none",
          token_embedding=True, vocab_size=8192)
SEQ_LEN = 512
SEED = 1919


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), kind="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, kind="fallback")


# ----------------------------------------------------------------------------------------------------------------------
# algorithmic bytes per (token, layer) of each kernel of the step (DESIGN.md "Kernels"), fp32 activations
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_token(cfg):
    D = cfg["hidden_dim"]; H = cfg["num_heads"]; N = cfg["state_dim"]; G = 1
    d_in = D + 2 * G * N + H
    return {
        "eigb200_mamba2_eig": D * 4 + 4 * H,                      # read x once, write lambda
        "eigb200_mamba_conv_ssd": (d_in + D) * 4,                 # read [x|B|C|dt], write y
        "eigb200_layernorm": 2 * D * 4,
        "eigb200_linear": None,                                   # per call, see linear_bytes
        "eigb200_embedding": 8 + D * 4,
    }


def clocks_sampler_start(path):
    try:
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        f = open(path, "w")
        return subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL), f
    except Exception:
        return None, None


def clocks_summary(proc, f, path, dev_index):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        proc.wait(timeout=5)
    except Exception:
        proc.kill()
    f.close()
    sm, mx, reasons = [], [], set()
    for line in open(path):
        parts = [p.strip() for p in line.split(",")]
        if len(parts) < 9 or parts[0] != str(dev_index):
            continue
        try:
            sm.append(float(parts[1])); mx.append(float(parts[2]))
        except ValueError:
            continue
        for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
            if val.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    hi = sorted(sm)[len(sm) // 2:]                                # samples under load dominate the upper half
    return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_pass_eigs_per_s(cfg, sample_b, steps, warmup, threads):
    import oracle as O
    import eigb200.layers as Ly
    torch.set_num_threads(threads)
    sd = Ly.init_mamba_state_dict(cfg, SEED)
    D = cfg["hidden_dim"]; hd = D // cfg["num_heads"]
    ocfg = dict(num_layers=cfg["num_layers"], d_inner=D, ngroups=1, d_state=cfg["state_dim"], nheads=D // hd, headdim=hd, prenorm=cfg["prenorm"])
    g = torch.Generator().manual_seed(42)
    X = torch.randint(0, cfg["vocab_size"], (sample_b, SEQ_LEN), generator=g)
    for _ in range(warmup):
        O.mamba_eval_pass_torch_cpu(X, sd, ocfg)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        eig, _, _ = O.mamba_eval_pass_torch_cpu(X, sd, ocfg)
        ts.append(time.perf_counter() - t0)
    return eig.size / float(np.mean(ts)), float(np.mean(ts))


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (oracle port; JAX / mamba_ssm are not installable here)."""
    if rank != 0:
        return
    cfg = dict(C2)
    threads = len(os.sched_getaffinity(0))
    sample_b = args.cpu_sample
    eps, sec = cpu_pass_eigs_per_s(cfg, sample_b, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": eps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, world), launch="host threads (torch CPU)"),
            "cpu_baseline": {"value": eps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d sequences x T=%d x %d layers per step (throughput is batch-linear)" % (sample_b, SEQ_LEN, cfg["num_layers"])},
            "e2e": {"value": eps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, world):
    return {"workload": "C2 mamba2-mqar: T=512 d_model=128 heads=1 d_state=16 conv=4 glu prenorm layers=4 vocab=8192",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "seq_len": SEQ_LEN, "parallelism": "batch-sharded x%d" % world,
            "l2": "inputs larger than L2 (activations 1.07 GB per layer per GPU)", "gemm": args.gemm,
            "launch": "cuda-graph replay" if args.graph else "eager"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_eigb200(args, rank, local, world):
    import torch.distributed as dist
    import eigb200.analysis as A
    import eigb200.dist as D
    import eigb200.layers as Ly
    import eigb200.ops as ops

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = dict(C2)
    cfg["_gemm_mode"] = args.gemm
    Bsz, T, nl, H = args.batch, SEQ_LEN, cfg["num_layers"], cfg["num_heads"]
    sd = Ly.init_mamba_state_dict(cfg, SEED)
    model = Ly.MambaDev(cfg, sd, dev)
    g = torch.Generator().manual_seed(42 + rank)
    X_host = torch.randint(0, cfg["vocab_size"], (Bsz, T), generator=g).pin_memory()
    X = X_host.to(dev)
    eig_host = torch.empty(Bsz, T, H, nl, dtype=torch.float32).pin_memory()
    counts_host = torch.empty(nl, Bsz, H, ops.NSLOT, dtype=torch.int32).pin_memory()
    n_eig = Bsz * T * H * nl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph = A.MambaPassGraph(model, X, want_eig=True) if args.graph else None

    def run_pass(Xd):
        return graph.run(Xd) if graph is not None else A.mamba_pass(model, Xd, want_eig=True)

    def step_resident():
        res = run_pass(None) if graph is not None else A.mamba_pass(model, X, want_eig=True)
        if world > 1:
            D.allreduce_moments(*ops.count_moments(res.counts.reshape(nl * Bsz, H, ops.NSLOT)))     # the one exchange step (statistics only)
        return res

    def step_e2e():
        Xd = X_host.to(dev, non_blocking=True)
        res = run_pass(Xd)
        eig_host.copy_(res.eig, non_blocking=True)
        counts_host.copy_(res.counts, non_blocking=True)
        if world > 1:
            D.allreduce_moments(*ops.count_moments(res.counts.reshape(nl * Bsz, H, ops.NSLOT)))
        torch.cuda.synchronize()

    # ---- device-resident throughput ----------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clk_path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % rank)
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, f = clocks_sampler_start(clk_path) if rank == 0 else (None, None)
    ops.LAUNCHES["n"] = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.LAUNCHES["n"]

    # ---- end to end with host buffers ----------------------------------------------------------------------------------
    # Every step copies ITS token ids from pinned host memory and brings ITS eigenvalue array + bin counts back to pinned host memory.  With the
    # graph path two batches are in flight (two captured passes with their own static buffers, copies on a second stream), so the PCIe
    # transfers of step i overlap the kernels of step i+1; --no-graph runs the strictly serial copy -> pass -> copy -> sync loop.
    if graph is not None:
        graphs = [graph, A.MambaPassGraph(model, X, want_eig=True)]
        eig_hosts = [eig_host, torch.empty_like(eig_host).pin_memory()]
        cnt_hosts = [counts_host, torch.empty_like(counts_host).pin_memory()]
        h2d_stream = torch.cuda.Stream(); d2h_stream = torch.cuda.Stream()   # separate queues: an upload must not wait behind the previous download
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]; ev_done = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]
        started = [False, False]

        def e2e_pipelined(nsteps):
            for i in range(nsteps):
                k = i & 1
                gk = graphs[k]
                with torch.cuda.stream(h2d_stream):
                    if started[k]:
                        h2d_stream.wait_event(ev_done[k])        # the previous pass on these buffers has read its input
                    gk.X.copy_(X_host, non_blocking=True)        # H2D of this step's token ids
                    ev_in[k].record(h2d_stream)
                main.wait_event(ev_in[k])
                if started[k]:
                    main.wait_event(ev_out[k])                   # the previous results of these buffers are on the host
                res = gk.run(None)
                if world > 1:
                    D.allreduce_moments(*ops.count_moments(res.counts.reshape(nl * Bsz, H, ops.NSLOT)))
                ev_done[k].record(main)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(ev_done[k])
                    eig_hosts[k].copy_(res.eig, non_blocking=True)   # D2H of this step's results
                    cnt_hosts[k].copy_(res.counts, non_blocking=True)
                    ev_out[k].record(d2h_stream)
                started[k] = True
            torch.cuda.synchronize()

        e2e_pipelined(max(2, args.warmup // 2))
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
    clocks = clocks_summary(proc, f, clk_path, local) if rank == 0 else None

    # ---- per-kernel durations (CUDA events around every C-ABI call of one more step) -> roofline of the dominant kernel --
    ops.PROFILE = []
    A.mamba_pass(model, X, want_eig=True)                          # eager (not the graph): one event pair per C-ABI call
    torch.cuda.synchronize()
    per = {}
    for name, s0, s1 in ops.PROFILE:
        per.setdefault(name, []).append(s0.elapsed_time(s1))
    ops.PROFILE = None

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return

    pk = peaks()
    D_, N_ = cfg["hidden_dim"], cfg["state_dim"]
    d_in = D_ + 2 * N_ + H
    ldz = (d_in + 7) // 8 * 8                                    # in_proj rows are padded to 32 bytes (layers._pad8)
    tokens = Bsz * T
    # ALGORITHMIC bytes per launch of every C-ABI call of the step: operands read once, results written once, fp32.
    # GEMMs are recorded per shape: all three are memory-bound at K = 128 (36-64 fp32 flop per byte, below the tensor ridge), so their
    # roofline is HBM; the fp32-equivalent tensor rate is reported next to it.
    alg = {"eigb200_mamba2_eig": tokens * (D_ * 4 + 4 * H + 8), "eigb200_mamba_conv_ssd": tokens * (d_in + D_) * 4,
           "eigb200_layernorm": tokens * 2 * D_ * 4, "eigb200_embedding": tokens * (8 + D_ * 4), "eigb200_embedding_stats": tokens * (8 + D_ * 4 + 8),
           "eigb200_linear_ln[N%d K%d none]" % (d_in, D_): tokens * (D_ + d_in) * 4 + tokens * 8,
           "eigb200_linear[N%d K%d none]" % (d_in, D_): tokens * (D_ + d_in) * 4,
           "eigb200_linear[N%d K%d gelu]" % (D_, D_): tokens * 2 * D_ * 4,
           "eigb200_linear[N%d K%d glu_residual]" % (2 * D_, D_): tokens * 3 * D_ * 4,
           # GLU + residual with the extractor partials of the output rows (D/16 groups x 3 floats per row), and the kernel that finishes them
           "eigb200_linear_glu_extract[N%d K%d glu_residual+extract]" % (2 * D_, D_): tokens * (3 * D_ * 4 + (D_ // 16) * 12),
           "eigb200_mamba2_eig_partials": tokens * ((D_ // 16) * 12 + 4 * H + 8)}
    flops = {"eigb200_linear_ln[N%d K%d none]" % (d_in, D_): 2.0 * tokens * D_ * d_in, "eigb200_linear[N%d K%d none]" % (d_in, D_): 2.0 * tokens * D_ * d_in,
             "eigb200_linear[N%d K%d gelu]" % (D_, D_): 2.0 * tokens * D_ * D_, "eigb200_linear[N%d K%d glu_residual]" % (2 * D_, D_): 2.0 * tokens * D_ * 2 * D_,
             "eigb200_linear_glu_extract[N%d K%d glu_residual+extract]" % (2 * D_, D_): 2.0 * tokens * D_ * 2 * D_}
    totals = {k: sum(v) for k, v in per.items()}
    step_total = sum(totals.values())
    dom = max(totals, key=totals.get)
    dom_avg_ms = totals[dom] / len(per[dom])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per launch from the committed ncu --set full captures
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    achieved = alg.get(dom, 0) / (dom_avg_ms * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
            "traffic": traffic, "peak_kind": pk["kind"] + " copy (MEASURED_PEAKS.json hbm_gbs)", "share_of_step": totals[dom] / step_total,
            "algorithmic_bytes_per_launch": alg.get(dom), "avg_launch_ms": dom_avg_ms}
    if dom in flops:
        roof["fp32_equiv_TFLOPs"] = flops[dom] / (dom_avg_ms * 1e-3) / 1e12
    kernels = {k: {"launches_per_step": len(v), "ms_per_step": sum(v)} for k, v in per.items()}
    for k in kernels:
        if k in alg:
            kernels[k]["GBps_alg"] = alg[k] * len(per[k]) / (kernels[k]["ms_per_step"] * 1e-3) / 1e9
            kernels[k]["frac_of_hbm_peak"] = kernels[k]["GBps_alg"] / pk["hbm"]

    value = n_eig * world * args.steps / (ms * 1e-3)
    e2e_val = n_eig * world * args.steps / (e2e_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": X_host.numel() * 8,
                    "d2h_bytes_per_step": eig_host.numel() * 4 + counts_host.numel() * 4,
                    "mode": "2 batches in flight (uploads and downloads on their own streams)" if args.graph else "serial copy-pass-copy"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels,
            "alg_bytes_per_eig": 2696, "e2e_path_frac_of_hbm": (2696.0 * n_eig / (ms / args.steps * 1e-3) / 1e9) / pk["hbm"]}
    if not args.no_cpu_baseline and world == 1:
        threads = len(os.sched_getaffinity(0))
        eps, sec = cpu_pass_eigs_per_s(dict(C2), args.cpu_sample, 2, 1, threads)
        line["cpu_baseline"] = {"value": eps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d sequences x T=%d x %d layers, mean of 2 passes after 1 warm-up" % (args.cpu_sample, SEQ_LEN, nl)}
    else:
        line["cpu_baseline"] = None
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract goes to the process' original stdout; everything libraries print (NCCL's version banner goes to
    stdout at any NCCL_DEBUG level >= VERSION) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                                  # fd 1 -> stderr for the rest of the run
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="eigb200", choices=["eigb200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="sequences per GPU")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tc3", "tc1"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="sequences in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "eigb200":
        args.warmup = 3                                            # timing rule: W >= 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the eigb200 path has no CPU fallback (use --impl reference for the host baseline)")
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("EIGB200_NCCL_DEBUG", "WARN")   # NCCL_DEBUG=VERSION prints to stdout: keep it to the ONE JSON line
        import eigb200.dist as D
        D.init_from_env("nccl")
    try:
        run_eigb200(args, rank, local, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
