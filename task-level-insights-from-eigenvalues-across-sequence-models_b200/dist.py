"""Multi-GPU plumbing: one process per GPU (torchrun), the analysis batch sharded by contiguous slices, and the ONE
exchange step of the path -- an integer all-reduce of the statistics buffer (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  Integer sums are order independent, so the combined statistics are bit-reproducible."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank`; the first batch % world ranks get one extra sequence."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_counts(local_counts: torch.Tensor, batch: int, lo: int, batch_axis: int = 1) -> torch.Tensor:
    """local_counts: integer tensor whose `batch_axis` spans this rank's slice [lo, lo+n).  Every rank writes its slice into
    a zeroed buffer of the global batch size and ONE all-reduce(sum) assembles the global per-sample bin counts on all ranks."""
    shape = list(local_counts.shape)
    n = shape[batch_axis]
    shape[batch_axis] = batch
    buf = torch.zeros(shape, dtype=local_counts.dtype, device=local_counts.device)
    idx = [slice(None)] * len(shape)
    idx[batch_axis] = slice(lo, lo + n)
    buf[tuple(idx)] = local_counts
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def allreduce_moments(sum_c: torch.Tensor, sum_c2: torch.Tensor):
    """Alternative for statistics-only runs: all-reduce sum_b c and sum_b c^2 (int64) and form mean/std from them."""
    if is_distributed():
        both = torch.stack([sum_c, sum_c2])
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
        return both[0], both[1]
    return sum_c, sum_c2


def mean_std_from_moments(sum_c, sum_c2, n_per_seq: int, batch: int):
    s1 = np.asarray(sum_c, np.float64); s2 = np.asarray(sum_c2, np.float64)
    mean = s1 * 100.0 / (n_per_seq * batch)
    var = np.maximum(s2 / batch - (s1 / batch) ** 2, 0.0)
    return mean, 100.0 / n_per_seq * np.sqrt(var)


def gather_batch(t: torch.Tensor, batch: int, lo: int, batch_axis: int = 0) -> torch.Tensor:
    """Assemble a batch-sharded floating tensor on every rank (used only when the caller wants the eigenvalue array itself)."""
    if not is_distributed():
        return t
    world = dist.get_world_size()
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    parts = []
    for r, (a, b) in enumerate(sizes):
        shape = list(t.shape); shape[batch_axis] = b - a
        parts.append(torch.empty(shape, dtype=t.dtype, device=t.device))
    dist.all_gather(parts, t.contiguous()) if len({b - a for a, b in sizes}) == 1 else _uneven_all_gather(parts, t, sizes)
    return torch.cat(parts, dim=batch_axis)


def _uneven_all_gather(parts, t, sizes):
    for r, p in enumerate(parts):
        if r == dist.get_rank():
            p.copy_(t)
        dist.broadcast(p, src=r)


class StatsComm:
    """Single-process, several-GPU form of the exchange step through the C ABI (include/eigb200.h: eigb200_stats_comm_init_all / eigb200_stats_allreduce):
    one communicator per listed device, allreduce(moments) sums the per-device (2, L, inner, 8) int64 moment buffers in place on every device."""

    def __init__(self, devices):
        import ctypes as C
        from . import _lib as L
        self._L, self._C = L, C
        self.devices = [int(d) for d in devices]
        lib = L.load()
        if not lib.eigb200_stats_available():
            raise L.Eigb200Error("NCCL is not available to libeigb200.so (set EIGB200_NCCL_LIB)")
        devs = (C.c_int * len(self.devices))(*self.devices)
        comms = (C.c_void_p * len(self.devices))()
        L.check(lib.eigb200_stats_comm_init_all(len(self.devices), devs, comms), "eigb200_stats_comm_init_all")
        self.comms = [C.c_void_p(c) for c in comms]

    def allreduce(self, moments):
        """moments: one contiguous int64 CUDA tensor per device of the communicator (same shape); summed in place."""
        L, C = self._L, self._C
        lib = L.load()
        assert len(moments) == len(self.devices)
        L.check(lib.eigb200_stats_group_start(), "eigb200_stats_group_start")
        try:
            for dev, comm, m in zip(self.devices, self.comms, moments):
                assert m.is_cuda and m.dtype == torch.int64 and m.is_contiguous() and m.device.index == dev
                L.check(lib.eigb200_set_device(dev), "eigb200_set_device")
                st = torch.cuda.current_stream(m.device).cuda_stream
                L.check(lib.eigb200_stats_allreduce(C.c_void_p(st), comm, C.c_void_p(m.data_ptr()), m.numel()), "eigb200_stats_allreduce")
        finally:
            L.check(lib.eigb200_stats_group_end(), "eigb200_stats_group_end")
        return moments

    def close(self):
        lib = self._L.load()
        for c in self.comms:
            lib.eigb200_stats_comm_destroy(c)
        self.comms = []
