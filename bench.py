#!/usr/bin/env python
"""bench.py -- eigenvalues/sec of the eval_eig hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path, BASELINE configs[1] (C2)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference algorithm on the host cores (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --config c5-mamba                              # another BASELINE config: c1 c3-lru c3-s5 c4 c5-mamba c5-normattn

Workload (config.workload), default: BASELINE configs[1] as written -- Mamba-2 on MQAR-shaped synthetic tokens, T=512, d_model=128, 1 head,
d_state=16, conv 4, GLU, prenorm, 4 layers, vocab 8192, ONE analysis batch of 4096 sequences sharded by sequence over the N GPUs
(`"scaling": "strong"`: 4096/N sequences per GPU; `--scaling weak` keeps 4096 per GPU, and at N > 1 the default line carries that curve too
under "weak").  One step = one analysis pass over the batch: token embedding, then per layer the block forward (LayerNorm, in_proj, causal
conv + SiLU, SSD scan, out_proj + GELU, GLU + residual) and the eigenvalue extractor + radius / phase bin counts on the block's output
(analysis/eval_eig.py:575-618).  One "eigenvalue" = one element of the returned `eig` array: B*T*H*L per step.

value   : whole-job eigenvalues/s with the token ids already resident in HBM (CUDA events, max over ranks).
e2e     : the same through the public API with HOST buffers: pinned token ids -> device, the pass, eigenvalue array + bin counts -> pinned host
          memory, inside the timed region, every step.  Two batches are in flight (two captured passes, uploads and downloads on their own
          streams), so the PCIe transfers of one step overlap the kernels of the next; `--no-graph` runs it strictly serially.
exchange: at N > 1 the only cross-rank step of the path (the int64 moments of the per-sample bin counts, eval_eig.py:620-623) is ONE NCCL
          all-reduce per pass.  It is issued on a side stream behind an event: nothing in pass i+1 depends on it, so it overlaps pass i+1.
roofline: dominant kernel of the step, STRICT algorithmic bytes (operands read once + results written once) / its event-timed duration,
          against MEASURED_PEAKS.json (and the nominal 8 TB/s beside it).
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "eigenvalues/sec in eval_eig"
UNIT = "eigenvalues/s"
NOMINAL_HBM_GBS = 8000.0
SCAN_FMA = {}

C2 = dict(layer="mamba", version="mamba2", num_layers=4, num_heads=1, input_dim=1, output_dim=8192, hidden_dim=128, state_dim=16,
          conv_dim=4, expansion=1, dropout=0.0, glu=True, norm="layer", dual=False, prenorm=True, pooling="none",
          token_embedding=True, vocab_size=8192)
SEQ_LEN = 512
SEED = 1919
OTHER_CONFIGS = ["c1", "c3-lru", "c3-s5", "c4", "c5-mamba", "c5-normattn"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), kind="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, kind="fallback")


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


def batch_plan(args, world):
    """-> (sequences per GPU, global batch).  strong: --batch is the GLOBAL analysis batch (BASELINE configs[1]: 4096), sharded like
    eigb200.dist.shard_bounds (the first batch % world ranks carry one more sequence; bench uses the largest shard on every rank so that the
    per-rank work is the job's critical path); weak: --batch sequences on every GPU."""
    if args.scaling == "weak":
        return args.batch, args.batch * world
    return (args.batch + world - 1) // world, args.batch


def workload_config(args, world):
    """Identical for the eigb200 arm and the --impl reference arm (the driver compares the two `config` objects)."""
    if args.config != "c2":
        import workloads as WL
        b = args.batch_other if args.batch_other else WL.DEFAULT_BATCH[args.config]
        return {"workload": "BASELINE %s (tools/workloads.py)" % args.config, "batch_per_gpu": b, "global_batch": b * world,
                "parallelism": "batch-sharded x%d" % world, "scaling": "weak", "l2": "L2 flushed between timed steps when the working set is below 512 MiB"}
    per_gpu, glob = batch_plan(args, world)
    return {"workload": "C2 mamba2-mqar: T=512 d_model=128 heads=1 d_state=16 conv=4 glu prenorm layers=4 vocab=8192",
            "batch_per_gpu": per_gpu, "global_batch": glob, "seq_len": SEQ_LEN, "parallelism": "batch-sharded x%d" % world,
            "scaling": args.scaling,
            "l2": "inputs larger than L2 (activations %.2f GB per layer per GPU)" % (per_gpu * SEQ_LEN * 128 * 4 / 1e9)}


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference's own code needs CUDA + mamba_ssm + JAX and
    cannot run on a CPU at all).  Each step is a bounded SAMPLE of the workload named in `config` (cpu_baseline.sample says which)."""
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    if args.config != "c2":
        import workloads as WL
        wl = WL.make(args.config, WL.CPU_SAMPLE[args.config], None)
        for _ in range(max(0, min(args.warmup, 1))):
            wl.cpu(WL.CPU_SAMPLE[args.config], threads)
        us, ts = 0, 0.0
        for _ in range(max(1, min(args.steps, 3))):
            u, s = wl.cpu(WL.CPU_SAMPLE[args.config], threads)
            us += u; ts += s
        eps, sec = us / ts, ts / max(1, min(args.steps, 3))
        unit, sample = wl.unit, "%d of the workload's sequences / matrices per step" % WL.CPU_SAMPLE[args.config]
    else:
        eps, sec = cpu_pass_eigs_per_s(dict(C2), args.cpu_sample, args.steps, args.warmup, threads)
        unit = UNIT
        sample = "%d sequences x T=%d x %d layers per step (throughput is batch-linear)" % (args.cpu_sample, SEQ_LEN, C2["num_layers"])
    metric = METRIC if args.config == "c2" else "%s in eval_eig [%s]" % (unit.replace("/s", "/sec"), args.config)
    line = {"impl": "reference", "metric": metric, "value": eps, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling if args.config == "c2" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "runtime": "host threads (torch CPU / NumPy)",
            "cpu_baseline": {"value": eps, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": eps, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm, C2
# ----------------------------------------------------------------------------------------------------------------------
class C2Run:
    """One (batch per GPU) measurement of the C2 pass on this rank: device-resident loop, end-to-end loop, per-kernel profile."""

    def __init__(self, args, rank, local, world, Bsz):
        import eigb200.analysis as A
        import eigb200.layers as Ly
        import eigb200.ops as ops
        self.A, self.ops = A, ops
        self.args, self.rank, self.world, self.Bsz = args, rank, world, Bsz
        self.dev = torch.device("cuda", local)
        cfg = dict(C2)
        cfg["_gemm_mode"] = args.gemm
        self.cfg = cfg
        self.T, self.nl, self.H = SEQ_LEN, cfg["num_layers"], cfg["num_heads"]
        self.sd = Ly.init_mamba_state_dict(cfg, SEED)
        self.model = Ly.MambaDev(cfg, self.sd, self.dev)
        g = torch.Generator().manual_seed(42 + rank)
        self.X_host = torch.randint(0, cfg["vocab_size"], (Bsz, self.T), generator=g).pin_memory()
        self.X = self.X_host.to(self.dev)
        self.n_eig = Bsz * self.T * self.H * self.nl
        self.graph = A.MambaPassGraph(self.model, self.X, want_eig=True, moments=world > 1) if args.graph else None
        # the exchange step: int64 (sum c, sum c^2) of the per-sample bin counts, all-reduced on a side stream (two staging buffers in flight)
        self.side = torch.cuda.Stream() if world > 1 else None
        self.mom_stage = [torch.zeros(2, self.nl, self.H, ops.NSLOT, dtype=torch.int64, device=self.dev) for _ in range(2)] if world > 1 else None
        self.ev_ready = [torch.cuda.Event() for _ in range(2)]
        self.ev_sent = [torch.cuda.Event() for _ in range(2)]
        self.sent = [False, False]
        self.k = 0

    def exchange(self, res, moments):
        """counts of this rank's pass -> global moments.  Issued behind the pass on `side`; the next pass does not wait for it."""
        if self.world == 1:
            return
        import torch.distributed as dist
        k = self.k; self.k ^= 1
        main = torch.cuda.current_stream()
        if self.sent[k]:
            main.wait_event(self.ev_sent[k])                       # staging buffer k was all-reduced two passes ago
        if moments is None:
            self.ops.count_moments_layers(res.counts, out=self.mom_stage[k])
        else:
            self.mom_stage[k].copy_(moments)
        self.ev_ready[k].record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_ready[k])
            dist.all_reduce(self.mom_stage[k], op=dist.ReduceOp.SUM)
            self.ev_sent[k].record(self.side)
        self.sent[k] = True

    def step_resident(self):
        if self.graph is not None:
            res = self.graph.run(None)
            self.exchange(res, self.graph.moments)
        else:
            res = self.A.mamba_pass(self.model, self.X, want_eig=True)
            self.exchange(res, None)
        return res

    def finish(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------------------------
    def time_resident(self, steps, warmup):
        ops = self.ops
        for _ in range(warmup):
            self.step_resident()
        self.finish(); self.barrier()
        ops.LAUNCHES["n"] = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            self.step_resident()
        self.finish()
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1), ops.LAUNCHES["n"]

    # ---- end to end with host buffers ------------------------------------------------------------------------------------
    # Every step copies ITS token ids from pinned host memory and brings ITS eigenvalue array + bin counts back to pinned host memory.  With the
    # graph path two batches are in flight (two captured passes with their own static buffers, copies on their own streams), so the PCIe
    # transfers of step i overlap the kernels of step i+1; --no-graph runs the strictly serial copy -> pass -> copy -> sync loop.
    def time_e2e(self, steps, warmup):
        A, ops = self.A, self.ops
        Bsz, T, H, nl = self.Bsz, self.T, self.H, self.nl
        eig_host = torch.empty(Bsz, T, H, nl, dtype=torch.float32).pin_memory()
        counts_host = torch.empty(nl, Bsz, H, ops.NSLOT, dtype=torch.int32).pin_memory()
        self.h2d = self.X_host.numel() * 8
        self.d2h = eig_host.numel() * 4 + counts_host.numel() * 4
        if self.graph is not None:
            graphs = [self.graph, A.MambaPassGraph(self.model, self.X, want_eig=True, moments=self.world > 1)]
            eig_hosts = [eig_host, torch.empty_like(eig_host).pin_memory()]
            cnt_hosts = [counts_host, torch.empty_like(counts_host).pin_memory()]
            h2d_stream = torch.cuda.Stream(); d2h_stream = torch.cuda.Stream()   # separate queues: an upload must not wait behind the previous download
            main = torch.cuda.current_stream()
            ev_in = [torch.cuda.Event() for _ in range(2)]; ev_done = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]
            started = [False, False]

            def run(nsteps):
                for i in range(nsteps):
                    k = i & 1
                    gk = graphs[k]
                    with torch.cuda.stream(h2d_stream):
                        if started[k]:
                            h2d_stream.wait_event(ev_done[k])        # the previous pass on these buffers has read its input
                        gk.X.copy_(self.X_host, non_blocking=True)   # H2D of this step's token ids
                        ev_in[k].record(h2d_stream)
                    main.wait_event(ev_in[k])
                    if started[k]:
                        main.wait_event(ev_out[k])                   # the previous results of these buffers are on the host
                    res = gk.run(None)
                    self.exchange(res, gk.moments)
                    ev_done[k].record(main)
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(ev_done[k])
                        eig_hosts[k].copy_(res.eig, non_blocking=True)   # D2H of this step's results
                        cnt_hosts[k].copy_(res.counts, non_blocking=True)
                        ev_out[k].record(d2h_stream)
                    started[k] = True
                self.finish()
                torch.cuda.synchronize()
            mode = "2 batches in flight (uploads and downloads on their own streams)"
        else:
            def run(nsteps):
                for _ in range(nsteps):
                    Xd = self.X_host.to(self.dev, non_blocking=True)
                    res = A.mamba_pass(self.model, Xd, want_eig=True)
                    eig_host.copy_(res.eig, non_blocking=True)
                    counts_host.copy_(res.counts, non_blocking=True)
                    self.exchange(res, None)
                    self.finish()
                    torch.cuda.synchronize()
            mode = "serial copy-pass-copy"
        run(max(2, warmup // 2))
        self.barrier()
        t0 = time.perf_counter()
        run(steps)
        self.barrier()
        return (time.perf_counter() - t0) * 1e3, mode

    # ---- per-kernel durations (CUDA events around every C-ABI call of one more step) -> roofline of the dominant kernel --
    def profile(self):
        ops = self.ops
        ops.PROFILE = []
        self.A.mamba_pass(self.model, self.X, want_eig=True)          # eager (not the graph): one event pair per C-ABI call
        torch.cuda.synchronize()
        per = {}
        for name, s0, s1 in ops.PROFILE:
            per.setdefault(name, []).append(s0.elapsed_time(s1))
        ops.PROFILE = None
        return per


def c2_alg_bytes(tokens, D_, N_, H):
    """STRICT algorithmic bytes per launch of every C-ABI call of the C2 step: operands read once + results written once, fp32.
    The extractor partials that the GLU epilogue leaves for eigb200_mamba2_eig_partials are this design's intermediate, not algorithmic traffic:
    they are counted for neither kernel (the combine kernel's algorithmic output is lambda + the next block's LayerNorm statistics)."""
    d_in = D_ + 2 * N_ + H
    alg = {"eigb200_mamba2_eig": tokens * (D_ * 4 + 4 * H + 8), "eigb200_mamba_conv_ssd": tokens * (d_in + D_) * 4,
           "eigb200_layernorm": tokens * 2 * D_ * 4, "eigb200_embedding": tokens * (8 + D_ * 4), "eigb200_embedding_stats": tokens * (8 + D_ * 4 + 8),
           "eigb200_linear_ln[N%d K%d none]" % (d_in, D_): tokens * (D_ + d_in) * 4 + tokens * 8,
           "eigb200_linear[N%d K%d none]" % (d_in, D_): tokens * (D_ + d_in) * 4,
           "eigb200_linear[N%d K%d gelu]" % (D_, D_): tokens * 2 * D_ * 4,
           "eigb200_linear[N%d K%d glu_residual]" % (2 * D_, D_): tokens * 3 * D_ * 4,
           "eigb200_linear_glu_extract[N%d K%d glu_residual+extract]" % (2 * D_, D_): tokens * 3 * D_ * 4,
           "eigb200_out_glu_fused[D%d K%d out_proj+gelu -> glu_residual+extract]" % (D_, D_): tokens * 3 * D_ * 4,   # read y, read the skip, write x_out
           "eigb200_mamba_front_fused[D%d P%d N%d ln+in_proj+conv+ssd]" % (D_, D_, N_): tokens * (D_ * 4 + 8 + D_ * 4),   # read x + its statistics, write y
           "eigb200_mamba2_eig_partials": tokens * (4 * H + 8)}
    flops = {"eigb200_linear_ln[N%d K%d none]" % (d_in, D_): 2.0 * tokens * D_ * d_in, "eigb200_linear[N%d K%d none]" % (d_in, D_): 2.0 * tokens * D_ * d_in,
             "eigb200_linear[N%d K%d gelu]" % (D_, D_): 2.0 * tokens * D_ * D_, "eigb200_linear[N%d K%d glu_residual]" % (2 * D_, D_): 2.0 * tokens * D_ * 2 * D_,
             "eigb200_linear_glu_extract[N%d K%d glu_residual+extract]" % (2 * D_, D_): 2.0 * tokens * D_ * 2 * D_,
             "eigb200_out_glu_fused[D%d K%d out_proj+gelu -> glu_residual+extract]" % (D_, D_): 2.0 * tokens * D_ * 3 * D_,
             # in_proj on the tensor cores + the recurrence on the FMA pipe (conv 4 + state update N + output N + D x, 2 flop each, per channel and token)
             "eigb200_mamba_front_fused[D%d P%d N%d ln+in_proj+conv+ssd]" % (D_, D_, N_): 2.0 * tokens * D_ * d_in + 2.0 * tokens * D_ * (2 * N_ + 5)}
    global SCAN_FMA
    SCAN_FMA = {"eigb200_mamba_front_fused[D%d P%d N%d ln+in_proj+conv+ssd]" % (D_, D_, N_): tokens * D_ * (2 * N_ + 5),
                "eigb200_mamba_conv_ssd": tokens * D_ * (2 * N_ + 5)}
    return alg, flops


def roofline_from_profile(per, alg, flops, pk, traffic_file=True):
    totals = {k: sum(v) for k, v in per.items()}
    step_total = sum(totals.values())
    dom = max(totals, key=totals.get)
    dom_avg_ms = totals[dom] / len(per[dom])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per launch from the committed ncu --set full captures
    if traffic_file and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    achieved = alg.get(dom, 0) / (dom_avg_ms * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
            "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS, "traffic": traffic,
            "peak_kind": pk["kind"] + " copy (MEASURED_PEAKS.json hbm_gbs)", "share_of_step": totals[dom] / step_total,
            "algorithmic_bytes_per_launch": alg.get(dom), "avg_launch_ms": dom_avg_ms}
    if dom in flops:
        roof["fp32_equiv_TFLOPs"] = flops[dom] / (dom_avg_ms * 1e-3) / 1e12
    if dom in SCAN_FMA:
        # this kernel is paced by the fp32 FMA pipe, not by HBM (profiles/: issue slots 67 % busy, DRAM 20 %): the selective-scan recurrence needs 2 N + 5 FMAs
        # per channel and token whatever the memory system does; FMA-pipe peak = SMs x 128 lanes x clock
        props = torch.cuda.get_device_properties(0)
        fma_peak = props.multi_processor_count * 128 * 1.965e9
        roof["paced_by"] = "fp32 FMA issue of the selective-scan recurrence (2 N + 5 FMAs per channel and token), not HBM"
        roof["fma_per_s"] = SCAN_FMA[dom] / (dom_avg_ms * 1e-3)
        roof["frac_of_fma_peak_at_1965MHz"] = roof["fma_per_s"] / fma_peak
    kernels = {k: {"launches_per_step": len(v), "ms_per_step": sum(v)} for k, v in per.items()}
    for k in kernels:
        if k in alg:
            kernels[k]["GBps_alg"] = alg[k] * len(per[k]) / (kernels[k]["ms_per_step"] * 1e-3) / 1e9
            kernels[k]["frac_of_hbm_peak"] = kernels[k]["GBps_alg"] / pk["hbm"]
            kernels[k]["frac_of_nominal_8TBs"] = kernels[k]["GBps_alg"] / NOMINAL_HBM_GBS
    return roof, kernels


def max_over_ranks(vals, dev, world):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def run_c2(args, rank, local, world):
    torch.cuda.set_device(local)
    per_gpu, glob = batch_plan(args, world)
    run = C2Run(args, rank, local, world, per_gpu)
    clk_path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % rank)
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, f = clocks_sampler_start(clk_path) if rank == 0 else (None, None)
    ms, launches = run.time_resident(args.steps, args.warmup)
    e2e_ms, e2e_mode = run.time_e2e(args.steps, args.warmup)
    clocks = clocks_summary(proc, f, clk_path, local) if rank == 0 else None
    per = run.profile()
    ms, e2e_ms = max_over_ranks([ms, e2e_ms], run.dev, world)
    n_eig_job = run.n_eig * world                                 # every rank runs the largest shard (batch_plan)
    if args.scaling == "strong":
        n_eig_job = glob * SEQ_LEN * run.H * run.nl                # the job is the 4096-sequence batch, whatever the padding of the last shard
    h2d, d2h = run.h2d, run.d2h

    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_curve:
        # the weak-scaling curve of round 1 (4096 sequences on EVERY GPU) as an extra key: same code, larger shard
        del run
        torch.cuda.empty_cache()
        wargs = argparse.Namespace(**vars(args)); wargs.scaling = "weak"
        wrun = C2Run(wargs, rank, local, world, args.batch)
        wms, _ = wrun.time_resident(args.steps, args.warmup)
        (wms,) = max_over_ranks([wms], wrun.dev, world)
        weak = {"value": wrun.n_eig * world * args.steps / (wms * 1e-3), "unit": UNIT, "ms_per_step": wms / args.steps,
                "batch_per_gpu": args.batch, "global_batch": args.batch * world, "scaling": "weak"}
        del wrun
        torch.cuda.empty_cache()
    if rank != 0:
        return

    pk = peaks()
    D_, N_, H = C2["hidden_dim"], C2["state_dim"], C2["num_heads"]
    alg, flops = c2_alg_bytes(per_gpu * SEQ_LEN, D_, N_, H)
    roof, kernels = roofline_from_profile(per, alg, flops, pk, traffic_file=(per_gpu == 4096))
    value = n_eig_job * args.steps / (ms * 1e-3)
    e2e_val = n_eig_job * args.steps / (e2e_ms * 1e-3)
    path_gbs = 2696.0 * value / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "runtime": {"launch": "cuda-graph replay (one launch per pass)" if args.graph else "eager", "gemm": args.gemm,
                        "exchange": ("one NCCL all-reduce of the (2, L, H, 8) int64 moments (%d bytes) per pass, on a side stream: it overlaps the next pass" % (2 * run_nl_h(C2) * 8 * 8)) if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "mode": e2e_mode},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels,
            "alg_bytes_per_eig": 2696,
            "e2e_path_frac_of_hbm": path_gbs / world / pk["hbm"], "e2e_path_frac_of_nominal_8TBs": path_gbs / world / NOMINAL_HBM_GBS}
    if weak is not None:
        line["weak"] = weak
    threads = len(os.sched_getaffinity(0))
    if not args.no_cpu_baseline and world == 1:
        eps, sec = cpu_pass_eigs_per_s(dict(C2), args.cpu_sample, 2, 1, threads)
        line["cpu_baseline"] = {"value": eps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d sequences x T=%d x %d layers, mean of 2 passes after 1 warm-up" % (args.cpu_sample, SEQ_LEN, C2["num_layers"])}
    else:
        line["cpu_baseline"] = None
    if world == 1 and not args.no_gpu_baseline:
        line["gpu_baseline"] = gpu_baseline_c2(local, min(per_gpu, args.gpu_sample))
    if world == 1 and not args.no_other_configs:
        line["other_configs"] = other_configs_summary(local, pk)
    emit(line)


def run_nl_h(cfg):
    return cfg["num_layers"] * cfg["num_heads"]


def gpu_baseline_c2(local, sample):
    """Eager PyTorch (+ fla's Triton simple-GLA for the SSD scan) on the same GPU, on a bounded sample of the same workload."""
    try:
        import eigb200.layers as Ly
        import gpu_baseline as GB
        sd = Ly.init_mamba_state_dict(dict(C2), SEED)
        X = torch.randint(0, C2["vocab_size"], (sample, SEQ_LEN), generator=torch.Generator().manual_seed(42)).to("cuda:%d" % local)
        out = GB.time_mamba_pass_eager(dict(C2), sd, X)
        torch.cuda.empty_cache()
        return out
    except Exception as e:                                         # a comparator must never take the headline down
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm, the other BASELINE configs (tools/workloads.py)
# ----------------------------------------------------------------------------------------------------------------------
def time_workload(wl, steps, warmup, flush=None, world=1):
    import eigb200.ops as ops
    for _ in range(warmup):
        wl.step()
    torch.cuda.synchronize()
    ops.LAUNCHES["n"] = 0
    ts = []
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); wl.step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    launches = ops.LAUNCHES["n"]
    wl.step_e2e()
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.step_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    ops.PROFILE = []
    wl.step()
    torch.cuda.synchronize()
    per = {}
    for name, s0, s1 in ops.PROFILE:
        per.setdefault(name, []).append(s0.elapsed_time(s1))
    ops.PROFILE = None
    return sum(ts), e2e_ms, launches, per


def other_configs_summary(local, pk):
    """Compact N=1 numbers of the other BASELINE configs at their per-GPU shapes (2 timed passes each after 1 warm-up; full lines: --config X)."""
    import workloads as WL
    out = {}
    dev = torch.device("cuda", local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for kind in OTHER_CONFIGS:
        try:
            wl = WL.make(kind, WL.DEFAULT_BATCH[kind], dev)
            ms, e2e_ms, launches, per = time_workload(wl, 2, 1, flush)
            tot = {k: sum(v) for k, v in per.items()}
            dom = max(tot, key=tot.get)
            out[kind] = {"workload": wl.workload, "batch_per_gpu": WL.DEFAULT_BATCH[kind], "batch_note": WL.BATCH_NOTE[kind],
                         "value": wl.units * 2 / (ms * 1e-3), "unit": wl.unit, "ms_per_step": ms / 2,
                         "e2e_value": wl.units * 2 / (e2e_ms * 1e-3), "gpu_launches_per_step": launches // 2,
                         "dominant_call": dom, "dominant_share": tot[dom] / sum(tot.values())}
            del wl
        except Exception as e:
            out[kind] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        torch.cuda.empty_cache()
    return out


def run_other(args, rank, local, world):
    import workloads as WL
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    kind = args.config
    b = args.batch_other if args.batch_other else WL.DEFAULT_BATCH[kind]
    wl = WL.make(kind, b, dev, seed_offset=rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clk_path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % rank)
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, f = clocks_sampler_start(clk_path) if rank == 0 else (None, None)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ms, e2e_ms, launches, per = time_workload(wl, args.steps, args.warmup, flush, world)
    clocks = clocks_summary(proc, f, clk_path, local) if rank == 0 else None
    ms, e2e_ms = max_over_ranks([ms, e2e_ms], dev, world)
    if rank != 0:
        return
    pk = peaks()
    tot = {k: sum(v) for k, v in per.items()}
    dom = max(tot, key=tot.get)
    kernels = {k: {"launches_per_step": len(v), "ms_per_step": sum(v)} for k, v in per.items()}
    roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, "traffic": None,
            "share_of_step": tot[dom] / sum(tot.values()), "avg_launch_ms": tot[dom] / len(per[dom])}
    if kind.startswith("c3") and "eigb200_diag_scan" in per:       # the scan is the path's kernel at C3: 16 B per state update
        sc = per["eigb200_diag_scan"]
        ach = b * wl.T * wl.P * 16 / (sum(sc) / len(sc) * 1e-3) / 1e9
        roof.update({"kernel": "eigb200_diag_scan", "achieved": ach, "frac": ach / pk["hbm"], "frac_of_nominal_8TBs": ach / NOMINAL_HBM_GBS,
                     "algorithmic_bytes_per_launch": b * wl.T * wl.P * 16, "avg_launch_ms": sum(sc) / len(sc),
                     "share_of_step": sum(sc) / sum(tot.values())})
    import re
    mt = re.match(r"eigb200_linear(?:_ln)?\[N(\d+) K(\d+) (\w+)\]", dom)
    if roof["achieved"] is None and mt and hasattr(wl, "T") and int(mt.group(2)) > 256:
        # streamed-operand GEMM (K > 256): memory-bound by intensity no more -- the fp16 split issues 3 tensor-core products per useful one; report the ISSUED
        # tensor rate (3 x 2 M N K) against the measured sustained cuBLAS bf16 peak (a kernel timed inside a long step), and the useful fp32-equivalent rate
        N_, K_ = int(mt.group(1)), int(mt.group(2))
        M_ = b * wl.T
        t = roof["avg_launch_ms"] * 1e-3
        useful = 2.0 * M_ * N_ * K_ / t / 1e12
        roof.update({"bound": "tensor", "achieved": 3 * useful, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": 3 * useful / pk["tf_sust"],
                     "frac_of_burst_peak": 3 * useful / pk["tf_burst"], "fp32_equiv_TFLOPs": useful, "peak_kind": "measured cuBLAS bf16, sustained (MEASURED_PEAKS.json)",
                     "note": "fp16-split operands: 3 kind::f16 MMAs per product term, counted as issued"})
    if roof["achieved"] is None and dom.startswith("eigb200_linattn_forward") and hasattr(wl, "T"):
        # chunked attention kernel: reads q, k, v and writes the context once: 4 d_model floats per token (SURVEY 8d counts every activation moved once)
        Dm = wl.cfg["hidden_dim"]
        nbytes = b * wl.T * 4 * Dm * 4
        ach = nbytes / (roof["avg_launch_ms"] * 1e-3) / 1e9
        roof.update({"achieved": ach, "frac": ach / pk["hbm"], "frac_of_nominal_8TBs": ach / NOMINAL_HBM_GBS, "algorithmic_bytes_per_launch": nbytes,
                     "paced_by": "legacy tensor path (mma.sync tf32 x3: 2.4 ms per launch at this shape) plus its non-MMA instructions, not HBM"})
    line = {"metric": "%s in eval_eig [%s]" % (wl.unit.replace("/s", "/sec"), kind), "value": wl.units * world * args.steps / (ms * 1e-3), "unit": wl.unit,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": workload_config(args, world), "workload_detail": wl.workload,
            "e2e": {"value": wl.units * world * args.steps / (e2e_ms * 1e-3), "unit": wl.unit, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": wl.h2d * world, "d2h_bytes_per_step": wl.d2h * world, "mode": "serial copy-pass-copy"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels}
    line["runtime"] = {"launch": getattr(wl, "launch", "eager: one C-ABI call per kernel from Python")}
    if not args.no_cpu_baseline and world == 1:
        threads = len(os.sched_getaffinity(0))
        u, s = wl.cpu(WL.CPU_SAMPLE[kind], threads)
        line["cpu_baseline"] = {"value": u / s, "unit": wl.unit, "cores": threads, "kind": "port",
                                "sample": "%d of the workload's sequences / matrices, one pass" % WL.CPU_SAMPLE[kind]}
    else:
        line["cpu_baseline"] = None
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract goes to the process' original stdout; everything libraries print (NCCL's INFO / VERSION chatter goes
    to stdout) has been redirected to stderr by main(), so NCCL_DEBUG stays whatever the caller set."""
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
    ap.add_argument("--config", default="c2", choices=["c2"] + OTHER_CONFIGS)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="c2: strong = --batch is the global analysis batch, sharded over the GPUs (BASELINE configs[1]); weak = --batch per GPU")
    ap.add_argument("--batch", type=int, default=4096, help="c2: sequences of the analysis batch (global with --scaling strong, per GPU with weak)")
    ap.add_argument("--batch-other", type=int, default=0, help="other configs: sequences (c4: features) per GPU; 0 = the config's default")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tc3", "tc1"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="sequences in the CPU baseline sample")
    ap.add_argument("--gpu-sample", type=int, default=512, help="sequences in the eager-PyTorch GPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-weak-curve", action="store_true")
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
        import eigb200.dist as D                                   # NCCL_DEBUG is left as the caller set it: its output lands on stderr (fd 1 is dup'ed)
        D.init_from_env("nccl")
    try:
        if args.config == "c2":
            run_c2(args, rank, local, world)
        else:
            run_other(args, rank, local, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
