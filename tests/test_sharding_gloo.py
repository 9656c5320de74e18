"""CPU tier: the N>1 plumbing (pair sharding, barrier, max-over-ranks) on world_size-2 gloo."""
import os

import torch.multiprocessing as mp

from meshopticalflow_b200 import sharding


def test_shards_partition_the_pairs():
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            seen = sorted(p for r in range(world) for p in sharding.shard_pairs(n, r, world))
            assert seen == list(range(n))
            sizes = [len(sharding.shard_pairs(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w = sharding.init_process_group("gloo")
    assert (r, w) == (rank, world)
    mine = sharding.shard_pairs(7, r, w)
    sharding.barrier()
    # each rank "processes" its pairs; the job time is the slowest rank's, the work is the sum
    t = sharding.max_over_ranks(10.0 * (rank + 1))
    total = sharding.sum_over_ranks(float(len(mine)))
    out.put((rank, mine, t, total))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_gloo_job():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == [0, 2, 4, 6] and results[1][1] == [1, 3, 5]
    assert all(r[2] == 20.0 and r[3] == 7.0 for r in results)


def test_row_blocks_cover_the_rows():
    """The row blocks of the partitioned-mesh path (dist.cu deals 32-row slices the same way)."""
    for rows in (1, 31, 32, 33, 1000, 3145728):
        slices = (rows + 31) // 32
        for world in (1, 2, 3, 8):
            blocks = sharding.row_blocks(slices, rows, world)
            assert blocks[0][0] == 0 and blocks[-1][1] == rows
            assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
            assert all(b[0] % 32 == 0 for b in blocks)
            if rows >= 32 * world:
                sizes = [b[1] - b[0] for b in blocks]
                assert max(sizes) - min(sizes) <= 64


def _id_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sharding.init_process_group("gloo")
    payload = bytes(range(128)) if rank == 0 else None  # what rank 0 gets from mof_dist_unique_id
    got = sharding.broadcast_bytes(payload, 128, 0)
    out.put((rank, got))
    sharding.shutdown()


def test_communicator_id_reaches_every_rank():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29811 + os.getpid() % 200
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == bytes(range(128)) and results[1][1] == bytes(range(128))
