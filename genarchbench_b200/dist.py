"""Multi-GPU plumbing for the bsw path: one process per GPU, pairs sharded with no data-path
collective (SURVEY.md 8e -- pairs are independent; the reference's only parallelism is an OpenMP loop
over pair batches, main_banded.cpp:338-350). torch.distributed is used only for the barrier and for
reducing the per-rank timing / counters that bench.py reports."""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import numpy as np


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend: str | None = None):
    """Initialises torch.distributed when WORLD_SIZE > 1; returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                        device_id=torch.device("cuda", local_rank))
            else:
                dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of n pairs owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def split_by_cells(len1: np.ndarray, len2: np.ndarray, parts: int) -> np.ndarray:
    """Cut points of a contiguous split balanced by sum(len1*len2) (strong-scaling host split)."""
    w = np.cumsum(len1.astype(np.int64) * len2.astype(np.int64))
    total = int(w[-1]) if len(w) else 0
    cuts = [0]
    for p in range(1, parts):
        cuts.append(int(np.searchsorted(w, total * p / parts)))
    cuts.append(len(w))
    return np.maximum.accumulate(np.array(cuts, dtype=np.int64))


def shard_seed(base_seed: int, rank: int) -> int:
    """Weak scaling: every rank generates its own batch of the same shape from a distinct seed."""
    return base_seed + 7919 * rank


def barrier() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shutdown() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def reduce_stats(sums: Sequence[float], maxes: Sequence[float], device=None) -> Tuple[list, list]:
    """All-reduces `sums` with SUM and `maxes` with MAX over ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(sums), list(maxes)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    s = torch.tensor(list(sums), dtype=torch.float64, device=device)
    m = torch.tensor(list(maxes), dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s.tolist(), m.tolist()
