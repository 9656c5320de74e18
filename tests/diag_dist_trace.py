"""Diagnostic (several GPUs, under torchrun): where a partitioned-mesh iteration spends its time. Runs `iterations` UpdateFlow
iterations of one synthetic mesh with the solves partitioned over the ranks, first as in production (graphs), then eagerly with
MOF_DIST_TRACE=1 (every exchange bracketed by CUDA events; rank 0 prints the totals when its context closes).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/diag_dist_trace.py [level] [iterations]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, sharding, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rank, local_rank, world = sharding.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init_process_group("nccl")
    v, t = synthetic.octahedron_sphere(level)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    for mode in ("graphs", "traced"):
        if mode == "traced":
            os.environ["MOF_DIST_TRACE"], os.environ["MOF_DIST_GRAPH"] = "1", "0"
        uid = sharding.broadcast_bytes(api.dist_unique_id() if rank == 0 else None, 128, 0, dev)
        al = api.Aligner(local_rank)
        al.dist_init(world, rank, uid)
        for rep in range(2):
            al.set_mesh(v, t)
            al.set_signals(a, b)
            al.reset_stats()
            sharding.barrier()
            t0 = time.perf_counter()
            al.iterate(iterations)
            al.synchronize()
            wall = time.perf_counter() - t0
        s = al.stats()
        if rank == 0:
            print(json.dumps({"mode": mode, "world": world, "vertices": int(v.shape[0]), "iterations": iterations, "wall_s": wall, "flow_solve_ms": s["flowSolveMs"],
                              "smooth_solve_ms": s["smoothSolveMs"], "flow_iterations": s["flowCgIterations"], "smooth_iterations": s["smoothCgIterations"]}), flush=True)
        al.close()
        sharding.barrier()
    sharding.shutdown()


if __name__ == "__main__":
    main()
