"""CPU tests of the host-side logic: report files byte-identical to the reference's, statistics from integer counts, the
N>1 path (batch sharding + the one all-reduce of the statistics buffer) on the gloo backend with world_size 2, checkpoint-tree
readers, and the torch-CPU baseline port against the reference's golden output."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import oracle as O
from conftest import load_golden, golden_model


def test_report_files_match_reference_bytes(tmp_path):
    import eigb200.analysis as A
    g = load_golden("report_files")
    thr = np.array([0.1, 0.5, 0.9, 1.0, 10, 100]); thp = np.array([1, 10, 45, 90, 180])
    p1 = str(tmp_path / "a.txt"); p2 = str(tmp_path / "b.txt")
    pct, pct_i = g["pct"], g["pct_i"]
    A.create_file_percentage(thr, pct, pct_i, pct.mean(1), pct_i.mean(1), pct.std(1), pct_i.std(1), path=p1)
    A.create_file_percentage_ssm(thr, thp, g["ps"], g["psi"], g["pp"], g["ppi"], path=p2)
    assert open(p1).read() == bytes(g["txt_torch"]).decode()
    assert open(p2).read() == bytes(g["txt_ssm"]).decode()


def test_percentages_from_counts_and_moments():
    import eigb200.extractors as E
    import eigb200.dist as D
    rng = np.random.default_rng(0)
    L, B, H, T = 3, 10, 2, 64
    eig = np.exp(rng.normal(0, 2, (B, T, H, L))).astype(np.float32)
    c = O.threshold_counts(eig, O.THRESHOLDS_RADIUS, axis=1)                   # (7,B,H,L)
    counts = np.zeros((L, B, H, 8), np.int32)
    counts[..., :7] = np.transpose(c, (3, 1, 2, 0))
    counts[..., 7] = T
    pct = E.percentages_from_counts(counts, T, 7)
    np.testing.assert_array_equal(pct, O.threshold_analysis(eig, O.THRESHOLDS_RADIUS))
    ph = E.phase_percentages_from_counts(counts, T, 6)
    np.testing.assert_array_equal(ph, O.threshold_analysis(np.zeros_like(eig), O.THRESHOLDS_PHASE))
    mean, std = D.mean_std_from_moments(counts.astype(np.int64).sum(1), (counts.astype(np.int64) ** 2).sum(1), T, B)
    np.testing.assert_allclose(np.transpose(mean[..., :7], (2, 1, 0)), pct.mean(1), rtol=1e-12)
    np.testing.assert_allclose(np.transpose(std[..., :7], (2, 1, 0)), pct.std(1), rtol=1e-9, atol=1e-9)


def test_shard_bounds_cover_batch():
    import eigb200.dist as D
    for B in (1, 7, 8, 64, 4097):
        for W in (1, 2, 3, 8):
            spans = [D.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, B, out_dir):
    import torch.distributed as dist
    import eigb200.dist as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    full = rng.integers(0, 500, (3, B, 2, 8)).astype(np.int32)                # (L,B,H,8) global per-sample counts
    lo, hi = D.shard_bounds(B, rank, world)
    local = torch.from_numpy(full[:, lo:hi].copy())
    glob = D.allreduce_counts(local, B, lo, batch_axis=1)                     # the one exchange step
    s1, s2 = D.allreduce_moments(local.long().sum(1), (local.long() ** 2).sum(1))
    eig_local = torch.from_numpy(np.arange(B * 5, dtype=np.float64).reshape(B, 5)[lo:hi].copy())
    gathered = D.gather_batch(eig_local, B, lo, batch_axis=0)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), glob=glob.numpy(), s1=s1.numpy(), s2=s2.numpy(), full=full, gathered=gathered.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_statistics_allreduce_gloo(B):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_gloo_worker, args=(world, _free_port(), B, d), nprocs=world, join=True)
        for r in range(world):
            z = np.load(os.path.join(d, "r%d.npz" % r))
            np.testing.assert_array_equal(z["glob"], z["full"])              # every rank holds the global per-sample counts, bit exact
            np.testing.assert_array_equal(z["s1"], z["full"].astype(np.int64).sum(1))
            np.testing.assert_array_equal(z["s2"], (z["full"].astype(np.int64) ** 2).sum(1))
            np.testing.assert_array_equal(z["gathered"], np.arange(B * 5, dtype=np.float64).reshape(B, 5))


def test_trained_layer_readers(tmp_path):
    import eigb200.ssm as S
    rng = np.random.default_rng(0)
    flat = {}
    for i in (0, 1, 10):
        flat["model/params/encoder/layers_%d/seq/nu_log" % i] = rng.normal(size=4).astype(np.float32) + i
        flat["model/params/encoder/layers_%d/seq/theta_log" % i] = rng.normal(size=4).astype(np.float32)
        flat["model/params/encoder/layers_%d/out1/kernel" % i] = rng.normal(size=(2, 2)).astype(np.float32)
    flat["model/params/encoder/encoder/kernel"] = rng.normal(size=(2, 2)).astype(np.float32)
    p = str(tmp_path / "ckpt.npz")
    np.savez(p, **flat)
    layers = S.get_trained_layers_ssm(p)
    assert len(layers) == 3 and set(layers[0]) == {"nu_log", "theta_log"}
    np.testing.assert_array_equal(layers[2]["nu_log"], flat["model/params/encoder/layers_10/seq/nu_log"])   # numeric, not string, order
    tree = {"model": {"params": {"encoder": {"layers_0": {"seq": {"nu_log": torch.zeros(3)}}, "encoder": {"kernel": torch.zeros(2)}}}}}
    p2 = str(tmp_path / "ckpt.pt")
    torch.save(tree, p2)
    assert len(S.get_trained_layers_ssm(p2)) == 1


def test_trained_layer_readers_refuse_pickle(tmp_path, monkeypatch):
    """A checkpoint path is user input: pickle files and torch files holding arbitrary objects are refused unless the caller opts in."""
    import pickle
    import eigb200._lib as L
    import eigb200.ssm as S
    monkeypatch.delenv("EIGB200_ALLOW_PICKLE", raising=False)
    tree = {"model": {"params": {"encoder": {"layers_0": {"seq": {"nu_log": np.zeros(3)}}}}}}
    p = str(tmp_path / "ckpt.pkl")
    with open(p, "wb") as f:
        pickle.dump(tree, f)
    with pytest.raises(L.Eigb200Error, match="pickle"):
        S.get_trained_layers_ssm(p)

    class Evil:
        def __reduce__(self):
            return (print, ("executed on load",))
    p2 = str(tmp_path / "evil.pt")
    torch.save({"model": {"params": Evil()}}, p2)
    with pytest.raises(Exception):
        S.get_trained_layers_ssm(p2)                              # weights_only=True rejects the non-tensor payload
    monkeypatch.setenv("EIGB200_ALLOW_PICKLE", "1")
    assert len(S.get_trained_layers_ssm(p)) == 1                  # explicit opt-in for a trusted legacy file


def test_init_layers_distributions():
    import eigb200.ssm as S
    cfg = dict(state_dim=64, hidden_dim=16, num_layers=2, r_min=0.9, r_max=0.99, max_phase=6.28)
    layers = S.get_init_layers_ssm(1919, {}, {}, cfg, 128, "lru", 4)
    lam = O.lru_lambda(layers[0]["nu_log"], layers[0]["theta_log"])
    assert (np.abs(lam) >= 0.9 - 1e-6).all() and (np.abs(lam) <= 0.99 + 1e-6).all()
    np.testing.assert_allclose(np.exp(layers[0]["gamma_log"]) ** 2, 1 - np.abs(lam) ** 2, rtol=1e-4)
    s4 = S.get_init_layers_ssm(1919, {}, {}, dict(state_dim=16, hidden_dim=4, num_layers=1), 128, "s4", 4)
    Lam, P, _, _, _ = O.make_dplr_hippo(16)
    np.testing.assert_allclose(s4[0]["Lambda_re"][:, 0], Lam.real, rtol=1e-6)
    np.testing.assert_allclose(np.sort(s4[0]["Lambda_im"][:, 2]), np.sort(Lam.imag), atol=1e-5)
    s5 = S.get_init_layers_ssm(1919, {}, {}, dict(state_dim=32, hidden_dim=4, num_layers=1, num_blocks=4), 128, "s5", 4)
    assert s5[0]["Lambda_re"].shape == (16,) and s5[0]["log_step"].shape == (16, 1)


def test_cpu_baseline_port_matches_reference_golden():
    sd, cfg, g = golden_model("model_mamba2")
    D = cfg["hidden_dim"]; hd = D // cfg["num_heads"]
    c = dict(num_layers=cfg["num_layers"], d_inner=D, ngroups=1, d_state=cfg["state_dim"], nheads=D // hd, headdim=hd, prenorm=True)
    eig, pct, ph = O.mamba_eval_pass_torch_cpu(g["X"], sd, c, chunk=8)
    np.testing.assert_allclose(eig, g["eig"], rtol=2e-5)
    assert np.abs(pct - g["percentage"]).max() <= 100.0 / eig.shape[1] + 1e-9


def test_gpu_baseline_ssd_matches_oracle():
    """tools/gpu_baseline.py is the eager-PyTorch comparator of bench.py: its chunked SSD must compute the same recurrence as the oracle
    (a comparator that computed something else would make the speed-up meaningless)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gpu_baseline as GB
    rng = np.random.default_rng(0)
    B, T, H, P, N = 2, 150, 2, 8, 16
    x = rng.normal(size=(B, T, H, P)).astype(np.float32); dt = rng.uniform(0.01, 0.2, (B, T, H)).astype(np.float32)
    A = -rng.uniform(1, 8, H).astype(np.float32); Bm = rng.normal(size=(B, T, 1, N)).astype(np.float32); Cm = rng.normal(size=(B, T, 1, N)).astype(np.float32)
    D = rng.normal(size=H).astype(np.float32)
    y = GB._ssd_torch_chunked(*[torch.from_numpy(a).double() for a in (x, dt, A, Bm, Cm, D)], chunk=64).numpy()
    ref = O.ssd_scan_sequential(x, dt, A, Bm, Cm, D)
    np.testing.assert_allclose(y, ref, rtol=1e-9, atol=1e-10)


def test_gpu_baseline_pass_matches_oracle_port():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gpu_baseline as GB
    sd, cfg, g = golden_model("model_mamba2")
    sdt = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
    eig, pct = GB.mamba_pass_eager(torch.from_numpy(g["X"]), sdt, cfg, ssd_impl="torch")
    np.testing.assert_allclose(eig, g["eig"], rtol=2e-5)
    assert np.abs(pct - g["percentage"]).max() <= 100.0 / eig.shape[1] + 1e-9


def test_bench_arms_share_one_config():
    """The driver compares the `config` object of the eigb200 arm with the --impl reference arm's: both come from one function, and the strong-scaling
    plan shards BASELINE configs[1]'s 4096 sequences over the GPUs."""
    import argparse
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    a = argparse.Namespace(config="c2", scaling="strong", batch=4096, batch_other=0)
    assert bench.batch_plan(a, 8) == (512, 4096) and bench.batch_plan(a, 1) == (4096, 4096) and bench.batch_plan(a, 3) == (1366, 4096)
    a.scaling = "weak"
    assert bench.batch_plan(a, 8) == (4096, 32768)
    a.scaling = "strong"
    c8 = bench.workload_config(a, 8)
    assert c8["global_batch"] == 4096 and c8["batch_per_gpu"] == 512 and c8["scaling"] == "strong"
    assert bench.workload_config(a, 8) == c8                       # deterministic: no arm-specific keys in it
    assert "launch" not in c8 and "gemm" not in c8
    alg, _ = bench.c2_alg_bytes(4096 * 512, 128, 16, 1)
    assert alg["eigb200_linear_glu_extract[N256 K128 glu_residual+extract]"] == 4096 * 512 * 3 * 128 * 4   # strict: no extractor partials
