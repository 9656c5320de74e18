"""One mesh over several GPUs: every rank aligns the SAME pair with the flow solves row-partitioned across the ranks,
then again on its own GPU alone, and compares. Run under torchrun (one rank per GPU) or plainly (world of 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_worker.py [level] [iterations]

Rank 0 prints one JSON line. No oracle, no reference: the single-GPU path is the comparison (it has its own parity tests)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, sharding, synthetic  # noqa: E402


def run(al, v, t, a, b, iterations):
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.reset_stats()
    t0 = time.perf_counter()
    al.iterate(iterations)
    al.synchronize()
    wall = time.perf_counter() - t0
    s = al.stats()
    return al.flow(), s, wall


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rank, local_rank, world = sharding.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init_process_group("nccl")
    v, t = synthetic.octahedron_sphere(level)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))

    uid = api.dist_unique_id() if rank == 0 else None
    uid = sharding.broadcast_bytes(uid, 128, 0, dev)
    al = api.Aligner(local_rank)
    al.dist_init(world, rank, uid)
    extra = {}
    if len(sys.argv) > 3 and sys.argv[3] == "ab":  # also with every coarse level replicated (the round-1 scheme), same process, same mesh
        os.environ["MOF_DIST_LEVEL_CELLS"] = "2000000000"
        run(al, v, t, a, b, iterations)
        _, s_r, wall_r = run(al, v, t, a, b, iterations)
        extra = {"replicated_coarse_levels": {"flow_solve_ms": sharding.max_over_ranks(s_r["flowSolveMs"], dev), "smooth_solve_ms": sharding.max_over_ranks(s_r["smoothSolveMs"], dev),
                                              "wall_s": sharding.max_over_ranks(wall_r, dev), "flow_iterations": s_r["flowCgIterations"]}}
        os.environ.pop("MOF_DIST_LEVEL_CELLS")
    flow_d, s_d, wall_d = run(al, v, t, a, b, iterations)
    flow_d2, s_d2, wall_d2 = run(al, v, t, a, b, iterations)  # steady state (pool warm)
    col_a, col_b = al.advect_vertices(0.5)
    al.close()

    os.environ["MOF_SMOOTH_AHEAD"] = "0"  # one stream on both sides of the comparison (the partitioned path has one)
    single = api.Aligner(local_rank)
    run(single, v, t, a, b, iterations)
    flow_s, s_s, wall_s = run(single, v, t, a, b, iterations)
    ref_a, ref_b = single.advect_vertices(0.5)
    single.close()

    rel = float(np.linalg.norm(flow_d2 - flow_s) / max(np.linalg.norm(flow_s), 1e-300))
    same_run = bool(np.array_equal(flow_d, flow_d2))
    worst = sharding.max_over_ranks(rel, dev)
    # every rank must hold the same result
    digest = float(np.abs(flow_d2).sum())
    spread = sharding.max_over_ranks(digest, dev) + sharding.max_over_ranks(-digest, dev)
    ms_d = sharding.max_over_ranks(s_d2["flowSolveMs"], dev)
    line = {"world": world, "level": level, "vertices": int(v.shape[0]), "iterations": iterations, "flow_rel_diff_vs_single_gpu": worst,
            "ranks_agree": spread == 0.0, "repeatable": same_run, "colour_max_diff": float(max(np.abs(col_a - ref_a).max(), np.abs(col_b - ref_b).max())),
            "flow_solve_ms_partitioned": ms_d, "flow_solve_ms_single_gpu": s_s["flowSolveMs"], "flow_iterations_partitioned": s_d2["flowCgIterations"],
            "flow_iterations_single_gpu": s_s["flowCgIterations"], "halo_entries_rank0": s_d2["haloEntries"], "rows": s_d2["flowRows"],
            "smooth_solve_ms_partitioned": s_d2["smoothSolveMs"], "smooth_solve_ms_single_gpu": s_s["smoothSolveMs"],
            "smooth_iterations_partitioned": s_d2["smoothCgIterations"], "smooth_iterations_single_gpu": s_s["smoothCgIterations"],
            "last_smooth_residual": s_d2["lastSmoothResidual"],
            "last_flow_residual": s_d2["lastFlowResidual"], "wall_s_partitioned": sharding.max_over_ranks(wall_d2, dev), "wall_s_single_gpu": wall_s}
    line.update(extra)
    sharding.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    sharding.shutdown()
    ok = worst < 1e-5 and spread == 0.0 and s_d2["lastFlowResidual"] <= 1.01e-8
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
