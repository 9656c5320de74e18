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


def broadcast_bytes(payload: bytes | None, size: int, src: int = 0, device=None) -> bytes:
    """`payload` (given on rank `src`, `size` bytes) to every rank; identity for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return bytes(payload)
    rank = dist.get_rank()
    t = torch.zeros(size, dtype=torch.uint8)
    if rank == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def row_blocks(num_slices: int, num_rows: int, world_size: int) -> list[tuple[int, int]]:
    """Rows [r0, r1) of each rank when `num_slices` 32-row slices are dealt in contiguous blocks (dist.cu uses the same
    rule: block k starts at slice floor(num_slices * k / world_size))."""
    starts = [min(num_rows, 32 * (num_slices * k // world_size)) for k in range(world_size)] + [num_rows]
    return [(starts[k], starts[k + 1]) for k in range(world_size)]


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
