"""Multi-GPU plumbing for independent signal pairs: one process per GPU, pairs dealt
round-robin to ranks, NO data-path collective — the only communication is the barrier and the
max-over-ranks of the device time that bench.py reports. torch.distributed (NCCL on the GPU box, gloo in
the CPU tests) is used for exactly that.
"""
from __future__ import annotations

import os


def shard_pairs(num_pairs: int, rank: int, world_size: int) -> list[int]:
    """Pair indices (= generator seeds) owned by `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_pairs, world_size))


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str):
    """Joins the job torchrun described (MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE); no-op for a single process."""
    import torch.distributed as dist
    rank, _, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def shutdown():
    """Leaves the process group (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
