"""`eval_eig` -- drop-in for analysis/eval_eig.py:462-857 on the eigb200 kernels.

Same signature, same return tuple, same files on disk (10 .npy + used_config.yaml under save_path+name, and
./percentage_file.txt), same quirks: the extractor of layer i is applied to the OUTPUT of block i (eval_eig.py:512-517);
`model_config.pop("layer")` mutates the caller's dict (:479); only the first batch of `loader` is analysed (:502); bins are
closed on both ends (:350, :359).  What differs is where the work happens: activations, eigenvalues and bin counts stay
in HBM; the eigenvalue array is copied to the host once, at the end, because the reference returns it.

Optional analysis-YAML keys beyond the reference's {batch_size, save_path} (defaults preserve reference behaviour):
  materialize_eig: bool = True   copy eig / eig_init to the host and save them (False: statistics only, eig = None)
  quantiles: [q, ...]            also save radius_quantiles.npy / radius_quantiles_init.npy (len(q), H, L) and radius_loghist(.._init).npy (H*L, 515): on-device
                                 log-spaced histogram (512 bins over [1e-8, 1e2)) of |eig| per (head, layer), summed over the GPUs, and its quantiles
  compare: "float64" | "float32" NumPy promotion reproduced by the threshold compare
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import yaml

from . import _lib as L
from . import dist as D
from . import extractors as E
from . import layers as Ly
from . import ops
from . import ssm

thresholds_radius = E.THRESHOLDS_RADIUS
thresholds_phase = E.THRESHOLDS_PHASE


# ----------------------------------------------------------------------------------------------------------------------
# device passes
# ----------------------------------------------------------------------------------------------------------------------

class PassResult:
    """eig: device tensor (B, T', H, L) -- already the reference's layout (np.concatenate along the last axis,
    eval_eig.py:524-526), each layer's kernel writes its own l-column -- or None; counts: (L, B, H, 8) int32 device tensor;
    n_per_seq = T'."""

    def __init__(self, eig, counts, n_per_seq, x_last):
        self.eig, self.counts, self.n_per_seq, self.x_last = eig, counts, n_per_seq, x_last

    def eig_host(self, pinned_out=None):
        """One device->host copy; no host-side permutation."""
        if self.eig is None:
            return None
        if pinned_out is not None:
            pinned_out.copy_(self.eig, non_blocking=True)
            return pinned_out
        # through page-locked memory: 0.6 ms instead of 15 ms for the 33 MB array of a C2 pass once torch's pinned allocator has a block of that size
        # (the first call pays the cudaHostAlloc); the returned ndarray keeps the block alive
        try:
            host = torch.empty(self.eig.shape, dtype=self.eig.dtype, pin_memory=True)
        except RuntimeError:
            return self.eig.cpu().numpy()
        host.copy_(self.eig, non_blocking=True)
        torch.cuda.current_stream(self.eig.device).synchronize()
        return host.numpy()


def mamba_pass(model: "Ly.MambaDev", X, pseudoLTI=False, want_eig=True, compare="float64") -> PassResult:
    """eval_eig.py:501-526 / :575-600 for one batch on the current device."""
    # LayerNorm fusion: the kernel that PRODUCES a block's input also emits that row's (mean, rstd) -- the embedding for block 0,
    # the extractor of block i for block i+1 -- and the in_proj GEMM normalises its A operand on the fly (eigb200_linear_ln).
    fuse = (not pseudoLTI) and isinstance(model.encoder, Ly.TokenEmbeddingsDev) and all(b.fuses_layernorm() for b in model.blocks)
    Bn, Tn = X.shape[0], X.shape[1]
    stats = torch.empty(2, Bn, Tn, 2, dtype=torch.float32, device=X.device) if fuse else None
    x = model.encoder(X, rowstats_out=stats[0]) if fuse else model.encoder(X)
    B, T, _ = x.shape
    H = model.blocks[0].mamba.nheads
    nl = len(model.blocks)
    eig = torch.empty(B, T, H, nl, dtype=torch.float32, device=x.device) if want_eig else None
    counts = torch.zeros(nl, B, H, ops.NSLOT, dtype=torch.int32, device=x.device)
    # extractor fusion: the GLU + residual GEMM that produces a block's output also leaves the per-row partials of x_out . W_dt and of the LayerNorm
    # moments; a small kernel finishes lambda, counts and statistics from them instead of K1 re-reading x_out (one head, M >= 1024 rows)
    fuse_x = (fuse and B * T >= 1024 and all(b.fuses_extractor() for b in model.blocks) and os.environ.get("EIGB200_EXTRACT_FUSION", "1") != "0")
    part = torch.empty(model.blocks[0].mamba.d_model // 16, 3, B * T, dtype=torch.float32, device=x.device) if fuse_x else None
    for i, blk in enumerate(model.blocks):
        if fuse_x:
            x = blk(x, stats[i & 1], extract_partials=part)
            ops.mamba2_eig_partials(part, B, T, blk.mamba.dt_bias, blk.mamba.A_log, want_lam=want_eig, counts=counts[i], compare=compare,
                                    lam_out=eig[..., i] if want_eig else None, rowstats_out=stats[(i + 1) & 1] if i + 1 < nl else None)
            continue
        x = blk(x, stats[i & 1]) if fuse else blk(x)
        out_i = eig[..., i] if want_eig else None
        if pseudoLTI:
            E.get_eig_mamba2_LTI_device(x, blk, want_eig=want_eig, counts=counts[i], compare=compare, lam_out=out_i)
        else:
            E.get_eig_mamba2_device(x, blk, want_eig=want_eig, counts=counts[i], compare=compare, lam_out=out_i,
                                    rowstats_out=stats[(i + 1) & 1] if (fuse and i + 1 < nl) else None)
    return PassResult(eig, counts, T, x)


def with_range_fallback(run_pass, model=None):
    """Run one analysis pass; if an fp16-split GEMM met an activation beyond its range (sticky device flag, see eigb200_gemm_overflow) the pass is
    repeated with the 3xTF32 operands, which have fp32's exponent range.  The flag is read once per pass, at a point where the host waits anyway."""
    ops.gemm_overflow(reset=True)
    res = run_pass()
    if ops.gemm_precision() == "f16x3" and ops.gemm_overflow(reset=True):
        import warnings
        warnings.warn("eigb200: an activation exceeded the range of the fp16-split GEMM operands; repeating the pass with 3xTF32", RuntimeWarning)
        ops.set_gemm_precision("tf32x3")
        try:
            if model is not None and hasattr(model, "invalidate_prepared"):
                model.invalidate_prepared()
            res = run_pass()
            torch.cuda.synchronize()
        finally:
            ops.set_gemm_precision(None)
            if model is not None and hasattr(model, "invalidate_prepared"):
                model.invalidate_prepared()
    return res


class MambaPassGraph:
    """One analysis pass captured as a CUDA graph: every kernel of mamba_pass is stream-ordered and allocation-free on the device side
    (outputs come from the graph's private pool), so a batch of fixed shape replays with ONE launch -- no per-kernel launch gaps and no host
    work between the ~40 kernels of a 4-layer pass.  run(X) copies the token ids into the static input and replays; the returned PassResult
    (eig, counts) is overwritten by the next run()."""

    def __init__(self, model: "Ly.MambaDev", X_example, pseudoLTI=False, want_eig=True, compare="float64", warmup=2, moments=False):
        """moments=True: the pass also leaves the batch moments (2,L,H,8) int64 of its bin counts in self.moments (inside the graph) -- the
        buffer a multi-GPU run all-reduces (eval_eig.py:620-623), so that no host work sits between the pass and the exchange."""
        if not X_example.is_cuda:
            raise L.Eigb200Error("MambaPassGraph: the example batch must live on the CUDA device")
        self.X = X_example.clone()
        self.moments = None

        def one_pass():
            res = mamba_pass(model, self.X, pseudoLTI, want_eig=want_eig, compare=compare)
            mom = ops.count_moments_layers(res.counts) if moments else None
            return res, mom
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                       # first calls may set function attributes / fill caches: outside the capture
                one_pass()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES["n"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result, self.moments = one_pass()
        self.launches_per_run = ops.LAUNCHES["n"] - n0

    def run(self, X=None) -> PassResult:
        if X is not None:
            self.X.copy_(X, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES["n"] += self.launches_per_run
        return self.result


class PassGraph:
    """Any stream-ordered pass `fn(X) -> result` over a batch of fixed shape (the LRU / S5 layer stack with its eigenvalue counts, a user's own composition of
    the layer calls) captured as a CUDA graph: run(X) copies X into the static input and replays; the returned result is overwritten by the next run()."""

    def __init__(self, fn, X_example, warmup=2):
        if not X_example.is_cuda:
            raise L.Eigb200Error("PassGraph: the example batch must live on the CUDA device")
        self.X = X_example.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                       # first calls set function attributes / fill caches: outside the capture
                fn(self.X)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES["n"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn(self.X)
        self.launches_per_run = ops.LAUNCHES["n"] - n0

    def run(self, X=None):
        if X is not None:
            self.X.copy_(X, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES["n"] += self.launches_per_run
        return self.result


class TransformerPassGraph:
    """transformer_pass captured as a CUDA graph, as MambaPassGraph: a small analysis batch (the reference's C1 run: 8 sequences x 64 tokens, 17 kernels of a few
    microseconds each) is bound by launch latency and host work between the kernels, not by the device; one replay removes both.  run(X) copies the token ids
    into the static input and replays; the returned PassResult is overwritten by the next run()."""

    def __init__(self, model: "Ly.TransformerDev", X_example, cfg, want_eig=True, compare="float64", warmup=2):
        if not X_example.is_cuda:
            raise L.Eigb200Error("TransformerPassGraph: the example batch must live on the CUDA device")
        self.X = X_example.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                       # first calls set function attributes / fill caches: outside the capture
                transformer_pass(model, self.X, cfg, want_eig=want_eig, compare=compare)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES["n"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = transformer_pass(model, self.X, cfg, want_eig=want_eig, compare=compare)
        self.launches_per_run = ops.LAUNCHES["n"] - n0

    def run(self, X=None) -> PassResult:
        if X is not None:
            self.X.copy_(X, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES["n"] += self.launches_per_run
        return self.result


def transformer_pass(model: "Ly.TransformerDev", X, cfg, want_eig=True, compare="float64") -> PassResult:
    """eval_eig.py:528-564 / :627-663."""
    x = model.encoder(X)
    B, T, _ = x.shape
    H, dqk, dm = cfg["num_heads"], cfg["state_dim"], cfg["hidden_dim"]
    nl = len(model.layers)
    eig = torch.empty(B, T - 1, H, nl, dtype=torch.float64, device=x.device) if want_eig else None
    counts = torch.zeros(nl, B, H, ops.NSLOT, dtype=torch.int32, device=x.device)
    fn = cfg["attention_fn"]
    for i, layer in enumerate(model.layers):
        x = layer(x)
        if fn == "lin-attention":
            att = layer.attention
            qk = ops.linear(x, att.W_qk, att.b_qk)
            nu = ops.linattn_nu(qk, 2 * dqk, B, T, H, att.head_dim, dqk)
            ops.ratio_hist(nu, L.RATIO_CUR_OVER_NEXT, want_out=want_eig, counts=counts[i], compare=compare, out=eig[..., i] if want_eig else None)
        elif fn == "norm-attention":
            att = layer.attention
            n = ops.normattn_gate(x, att.W_n, att.b_n, att.inner_attn.offset if cfg["offset"] else None, cfg["norm_fn"])
            ops.ratio_hist(n, L.RATIO_NEXT_OVER_CUR, want_out=want_eig, counts=counts[i], compare=compare, out=eig[..., i] if want_eig else None)
        elif fn == "sm-attention":
            att = layer.attention
            qk = ops.linear(x, att.W_qk, att.b_qk)
            nu, m = ops.softmax_nu(qk, 2 * dqk, B, T, H, att.head_dim, dqk)
            eta, _ = ops.softmax_eta(nu, m, want_out=want_eig, counts=counts[i])
            if want_eig:
                eig[..., i].copy_(eta)
        else:
            raise RuntimeError("{0} is not a valid model option".format(fn))
    return PassResult(eig, counts, T - 1, x)


# ----------------------------------------------------------------------------------------------------------------------
# report files (byte-compatible with create_file_percentage(_ssm), eval_eig.py:393-459)
# ----------------------------------------------------------------------------------------------------------------------

def create_file_percentage(thr_radius, percentage, percentage_init, percentage_mean, percentage_init_mean, percentage_std,
                           percentage_init_std, path="percentage_file.txt"):
    nheads, nlayers = np.shape(percentage)[2], np.shape(percentage)[3]
    batch_selection = np.array([0, 2, 4, 6])
    lines = []

    def emit(*parts):
        lines.append(" ".join(str(p) for p in parts))

    emit("threshold radius:", thr_radius, "\n")
    emit("batch selection:", batch_selection, "\n")
    for bi, b in enumerate(batch_selection):
        for h in range(nheads):
            for tag, arr in (("radius init: ", percentage_init), ("radius: ", percentage)):
                for l in range(nlayers):
                    emit("percentage batch dimension", b, "head", h, "layer", l, tag, np.round(arr[:, b, h, l], 1))
            if bi == 0:
                for stat, a_init, a in (("mean", percentage_init_mean, percentage_mean), ("std", percentage_init_std, percentage_std)):
                    for tag, arr in (("radius init: ", a_init), ("radius: ", a)):
                        for l in range(nlayers):
                            emit("percentage batch %s head" % stat, h, "layer", l, tag, np.round(arr[:, h, l], 1))
            emit("\n")
        emit("\n")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def create_file_percentage_ssm(thr_radius, thr_phase, percentage, percentage_init, percentage_phase, percentage_phase_init,
                               path="percentage_file.txt"):
    nlayers = np.shape(percentage)[1]
    lines = []

    def emit(*parts):
        lines.append(" ".join(str(p) for p in parts))

    emit("threshold radius:", thr_radius, "\n")
    emit("threshold phase:", thr_phase, "\n")
    blocks = (("radius init: ", percentage_init), ("radius: ", percentage), ("phase init: ", percentage_phase_init), ("phase: ", percentage_phase))
    for k, (tag, arr) in enumerate(blocks):
        for l in range(nlayers):
            emit("percentage layer", l, tag, np.round(arr[:, l], 1))
        if k < len(blocks) - 1:
            emit("\n")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


RESULT_FILES = ["eig", "eig_init", "percentage", "percentage_init", "percentage_phase", "percentage_phase_init",
                "percentage_mean", "percentage_init_mean", "percentage_std", "percentage_init_std"]


def _save_results(args, conf_args, wandb_config, data_config, model_config, train_config, perf, arrays: Dict[str, object]):
    """eval_eig.py:750-851: W&B artifact or 10 .npy + used_config.yaml under save_path + name_model."""
    path_to_percentage_file = os.path.abspath(os.getcwd()) + "/percentage_file.txt"
    dim_conv = model_config["dim_conv"] if "dim_conv" in model_config else 0
    if wandb_config is not None:
        import tempfile
        import wandb
        print("Saving artifact on W&B....")
        base = data_config["name"] + "{0}-dmodel{1}-seed{4}-num_layers{5}-dqk{2}-conv_dim{6}-lr{3}".format(
            wandb_config["name"], model_config["hidden_dim"], model_config["state_dim"], train_config["lr"], args["seed"],
            model_config["num_layers"], dim_conv)
        name_model = base + "-perf{0:0.3f}".format(perf)
        wandb.init(group="artifact_upload", entity=wandb_config["entity"], project=wandb_config["project"],
                   name="upload" + name_model, job_type="add-dataset")
        artifact = wandb.Artifact(name="eigen_values_" + base, type="dataset")
        art_names = ["eigen_values_", "eigen_values_init_", "percentage_", "percentage_init_", "percentage_phase_",
                     "percentage_pase_init_", "percentage_mean_", "percentage_init_mean_", "percentage_std_", "percentage_init_std_"]
        with tempfile.TemporaryDirectory() as tmpdir:
            for fname, aname in zip(RESULT_FILES, art_names):
                pth = os.path.join(tmpdir, fname + ".npy")
                np.save(pth, arrays[fname])
                artifact.add_file(local_path=pth, name=aname + name_model)
            cfg_path = os.path.join(tmpdir, "used_config.yaml")
            with open(cfg_path, "w") as file:
                yaml.dump(args, file, default_flow_style=False, sort_keys=False)
            artifact.add_file(local_path=path_to_percentage_file, name="percentage_file_" + name_model)
            artifact.add_file(local_path=cfg_path, name="used_config-" + name_model)
            artifact.save()
        try:
            wandb.finish()
        except Exception:
            pass
        return None
    print("Saving artifact locally....")
    save_path = conf_args["save_path"] if "save_path" in conf_args else ""
    base = data_config["name"] + "dmodel{0}-seed{3}-num_layers{4}-dqk{1}-conv_dim{5}-lr{2}".format(
        model_config["hidden_dim"], model_config["state_dim"], train_config["lr"], args["seed"], model_config["num_layers"], dim_conv)
    out_dir = save_path + base + "-perf{0:0.3f}".format(perf)
    try:
        os.mkdir(out_dir)
        print(f"Directory '{out_dir}' created successfully.")
    except FileExistsError:
        print(f"Directory '{out_dir}' already exists.")
    except PermissionError:
        print(f"Permission denied: Unable to create '{out_dir}'.")
    except Exception as e:
        print(f"An error occurred: {e}")
    # the two eigenvalue arrays are tens of MB each: write the ten files concurrently (np.save releases the GIL while it copies and writes)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(lambda fname: np.save(os.path.join(out_dir, fname + ".npy"), arrays[fname]), RESULT_FILES))
    with open(os.path.join(out_dir, "used_config.yaml"), "w") as file:
        yaml.dump(args, file, default_flow_style=False, sort_keys=False)
    return out_dir


# ----------------------------------------------------------------------------------------------------------------------
# the entry point
# ----------------------------------------------------------------------------------------------------------------------

def _first_batch(loader, device, lo=None, hi=None):
    """eval_eig.py:502 -- first batch only.  With several ranks the batch is drawn ONCE, on rank 0, and broadcast: a shuffling loader (or
    per-rank RNG state) would otherwise hand every rank a different batch and the all-reduced statistics would silently mix them."""
    if not torch.cuda.is_available():
        raise L.Eigb200Error("eval_eig needs a CUDA device (the reference refuses to run without one too, launch.py:63-64)")
    rank, world = D.rank_world()
    if world > 1:
        import torch.distributed as tdist
        meta = [None]
        X = None
        if rank == 0:
            X, _y, _ = next(iter(loader))
            X = torch.as_tensor(X)
            meta = [(tuple(X.shape), X.dtype)]
        tdist.broadcast_object_list(meta, src=0)
        shape, dtype = meta[0]
        Xd = X.to(device) if rank == 0 else torch.empty(shape, dtype=dtype, device=device)
        tdist.broadcast(Xd, src=0)
        return Xd[lo:hi].contiguous() if lo is not None else Xd
    X, y, _ = next(iter(loader))
    if lo is not None:
        X = X[lo:hi]
    return X.to(device, non_blocking=True)


def _load_torch_checkpoint(path):
    return torch.load(path, weights_only=True, map_location="cpu")          # eval_eig.py:569


def eval_eig(args, conf_args, wandb_config, data_config, loader, path_file, perf):
    model_config = args["model"]
    train_config = args["train"]
    data_config = args["dataset"]
    seed = args["seed"]
    batch_size = conf_args["batch_size"]
    num_layers = model_config["num_layers"]
    pseudoLTI = model_config["pseudoLTI"] if "pseudoLTI" in model_config else False
    materialize = conf_args.get("materialize_eig", True)
    quantile_list = [float(q) for q in conf_args.get("quantiles", [])] if hasattr(conf_args, "get") else []
    compare = conf_args.get("compare", "float64")

    path = path_file if os.path.isabs(path_file) else os.path.abspath(os.getcwd()) + "/" + path_file
    layer_type = model_config.pop("layer")                                   # mutates the caller's dict, like :479

    rank, world = D.rank_world()
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None

    if layer_type in ["mamba", "transformer"]:
        num_heads = model_config["num_heads"]
        lo, hi = D.shard_bounds(batch_size, rank, world) if world > 1 else (0, batch_size)

        def run(state_dict):
            X = _first_batch(loader, device, lo if world > 1 else None, hi if world > 1 else None)
            if layer_type == "mamba":
                model = Ly.MambaDev(model_config, state_dict, device)
                res = with_range_fallback(lambda: mamba_pass(model, X, pseudoLTI, want_eig=materialize, compare=compare), model)
            else:
                model = Ly.TransformerDev(model_config, state_dict, device)
                res = with_range_fallback(lambda: transformer_pass(model, X, model_config, want_eig=materialize, compare=compare), None)
            counts = D.allreduce_counts(res.counts, batch_size, lo, batch_axis=1)          # the one exchange step
            res.loghist = res.quant = None
            if quantile_list and res.eig is not None:                                          # finer on-device view of the same radii (north star; not in the reference)
                mag = res.eig if res.eig.dtype != torch.complex64 else torch.abs(res.eig)
                hist = ops.log_hist(torch.abs(mag).reshape(mag.shape[0], mag.shape[1], -1))
                if world > 1:
                    import torch.distributed as tdist
                    tdist.all_reduce(hist, op=tdist.ReduceOp.SUM)
                res.loghist = hist.cpu().numpy()
                res.quant = ops.hist_quantiles(hist, quantile_list).cpu().numpy().T.reshape((len(quantile_list),) + tuple(mag.shape[2:]))
            eig = res.eig
            if eig is not None and world > 1:
                eig = D.gather_batch(eig, batch_size, lo, batch_axis=0)
            res.eig, res.counts = eig, counts
            return res

        # init pass: the reference constructs the model under torch.manual_seed(seed) (:484-497) and never calls model.eval() on it, so
        # with dropout > 0 its eig_init / percentage_init carry one random dropout mask; this path is deterministic (dropout = identity
        # in both passes).  All reference analysis configs use dropout 0; say so loudly otherwise.
        if float(model_config.get("dropout", 0.0) or 0.0) > 0.0 or float(model_config.get("att_dropout", 0.0) or 0.0) > 0.0:
            import warnings
            warnings.warn("eigb200.eval_eig: dropout > 0 -- the reference's init pass runs in train mode (random dropout mask in eig_init); "
                          "eigb200 evaluates both passes without dropout, so eig_init / percentage_init differ from the reference's sample",
                          RuntimeWarning)
        if layer_type == "mamba":
            init_sd = Ly.init_mamba_state_dict(model_config, seed)
        else:
            init_sd = Ly.init_transformer_state_dict(model_config, seed)
        res_init = run(init_sd)
        res = run(_load_torch_checkpoint(path))                                             # :569-600 / :627-663

        nb_r, nb_p = len(thresholds_radius) + 1, len(thresholds_phase) + 1
        c_init = res_init.counts.cpu().numpy(); c = res.counts.cpu().numpy()
        n_per = res.n_per_seq
        percentage_init = E.percentages_from_counts(c_init, n_per, nb_r)
        percentage = E.percentages_from_counts(c, n_per, nb_r)
        percentage_phase_init = E.phase_percentages_from_counts(c_init, n_per, nb_p)
        percentage_phase = E.phase_percentages_from_counts(c, n_per, nb_p)
        percentage_init_mean = np.mean(percentage_init, axis=1); percentage_init_std = np.std(percentage_init, axis=1)
        percentage_mean = np.mean(percentage, axis=1); percentage_std = np.std(percentage, axis=1)
        eig_init = res_init.eig_host(); eig = res.eig_host()
        if rank == 0:
            create_file_percentage(thresholds_radius, percentage, percentage_init, percentage_mean, percentage_init_mean,
                                   percentage_std, percentage_init_std)
        extra_arrays = {}
        if quantile_list and res.quant is not None:
            extra_arrays = dict(radius_quantiles=res.quant, radius_quantiles_init=res_init.quant, radius_loghist=res.loghist, radius_loghist_init=res_init.loghist)

    elif layer_type in ["lru", "s4", "s5"]:
        SEQ_LEN = model_config["seq_len"]
        dim_idx = 1                                                                          # :689
        init_layers = ssm.get_init_layers_ssm(seed, data_config, train_config, model_config, SEQ_LEN, layer_type, batch_size)
        trained_layers = ssm.get_trained_layers_ssm(path)
        eig_init = np.concatenate([ssm.get_eigvals_ssm(layer_type, init_layers, i, dim_idx, SEQ_LEN) for i in range(num_layers)], axis=-1)
        eig = np.concatenate([ssm.get_eigvals_ssm(layer_type, trained_layers, i, dim_idx, SEQ_LEN) for i in range(num_layers)], axis=-1)
        rad_init, ph_init = ssm.radius_phase(eig_init)
        rad, ph = ssm.radius_phase(eig)
        percentage_init = E.threshold_analysis_ssm(rad_init, thresholds_radius, num_layers, compare)
        percentage = E.threshold_analysis_ssm(rad, thresholds_radius, num_layers, compare)
        percentage_phase_init = E.threshold_analysis_ssm(ph_init, thresholds_phase, num_layers, compare)
        percentage_phase = E.threshold_analysis_ssm(ph, thresholds_phase, num_layers, compare)
        percentage_init_mean = percentage_init_std = percentage_mean = percentage_std = 0   # :740-743
        if rank == 0:
            create_file_percentage_ssm(thresholds_radius, thresholds_phase, percentage, percentage_init, percentage_phase, percentage_phase_init)
    else:
        raise RuntimeError("{0} is not a valid model option".format(layer_type))             # :748

    if rank == 0:
        out_dir = _save_results(args, conf_args, wandb_config, data_config, model_config, train_config, perf,
                      dict(eig=eig, eig_init=eig_init, percentage=percentage, percentage_init=percentage_init,
                           percentage_phase=percentage_phase, percentage_phase_init=percentage_phase_init,
                           percentage_mean=percentage_mean, percentage_init_mean=percentage_init_mean,
                           percentage_std=percentage_std, percentage_init_std=percentage_init_std))
        if out_dir is not None and layer_type in ["mamba", "transformer"]:
            for name, arr in extra_arrays.items():                                           # beyond the reference's 10 files, only when `quantiles` was asked for
                np.save(os.path.join(out_dir, name + ".npy"), arr)
    return eig, eig_init, percentage, percentage_init, percentage_phase, percentage_phase_init
