#!/usr/bin/env python
"""bench.py — 1M-vertex pair alignments/sec on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--level L] [--partitioned]

A step is one whole alignment of one signal pair on the synthetic 1 048 578-vertex sphere
(BASELINE.json configs[2]): mesh-operator assembly, DoG normalisation, 10 UpdateFlow iterations and the
final halfway advection — what the reference does between loading its inputs and writing its output.

  value   inputs (positions, triangles, both colour signals) already resident in HBM, results left in HBM;
          timed with CUDA events on the stream the library launches on.
  e2e     the same step through the host-pointer C ABI (mof_set_mesh / mof_set_signals / mof_iterate /
          mof_advect_vertices) from pinned host buffers: the host->device copies of the mesh and signals and
          the device->host read of the advected colours are inside the timed region.
  N > 1   independent pairs sharded over the ranks (configs[3]), one process per GPU under torchrun, no
          data-path collective; weak scaling; time = max over ranks.
          --partitioned (configs[4]): instead, ONE pair per step for the whole job, its linear solves row-partitioned
          over the ranks (NCCL halo exchange + all-reduce); strong scaling; use with --level 10 / 11.
  roofline      the PCG's SpMV+dot kernel on the workload's own flow matrix, timed live with CUDA events; roofline.kernels = the
                same for every large kernel of a PCG iteration and the walk (mof_time_kernel: each alone, the solver's own grid),
                with launches per step and share of the step; roofline.flow_iteration_frac = the streaming kernels' algorithmic
                bytes of one flow PCG iteration / peak bandwidth / the measured time of an iteration.
  cpu_baseline  the reference's own binary (oracle/_ref, built from the unmodified sources) timed on this box's
                host cores on a bounded sample, rank 0, N=1 only; cpu_baseline.like_for_like = the SAME input (65 538-vertex
                sphere pair, 3 iterations) timed through the reference binary and through this library's host-pointer API in
                this run; cpu_baseline.growth = the reference's seconds per alignment at the sizes it finishes in seconds.
  N > 1 also appends `partitioned`: one 4.2M-vertex mesh (3 iterations) with its solves row-partitioned over all ranks, against
                the same on rank 0 alone (BASELINE.json configs[4]).
  --impl reference   only the reference arm, same metric/unit/config, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1M-vertex pair alignments/sec"
UNIT = "alignments/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "OpticalFlow_ref")
SAMPLE_LEVEL = 6  # 16 386 vertices: about 10 s of reference CPU work per alignment
LIKE_LEVEL, LIKE_ITERATIONS = 7, 3  # the like-for-like pair: 65 538 vertices, 3 iterations (~45 s of reference CPU work)
PARTITIONED_LEVEL, PARTITIONED_ITERATIONS = 10, 3  # the `partitioned` sub-record of N > 1: 4 194 306 vertices


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- reference arm

def reference_sample_inputs(tmp: str, level: int = SAMPLE_LEVEL):
    from meshopticalflow_b200 import synthetic
    v, t = synthetic.octahedron_sphere(level)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    synthetic.write_ply_colored(os.path.join(tmp, "A.ply"), v, a, t)
    synthetic.write_ply_colored(os.path.join(tmp, "B.ply"), v, b, t)
    return v.shape[0]


def time_reference_once(tmp: str, iterations: int = 10) -> float:
    """One alignment of the bounded sample by the reference's CPU implementation, all host threads."""
    t0 = time.perf_counter()
    if os.path.exists(REF_BIN):
        subprocess.check_call([REF_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", str(iterations)], cwd=tmp, stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    else:  # the oracle port (only when the reference binary could not be built)
        from meshopticalflow_b200 import synthetic
        from oracle import mof_oracle as O
        p = synthetic.read_ply(os.path.join(tmp, "A.ply"))
        q = synthetic.read_ply(os.path.join(tmp, "B.ply"))
        v = np.stack([p["vertex"][k] for k in "xyz"], 1).astype(np.float64)
        ca = np.stack([p["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.float64)
        cb = np.stack([q["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.float64)
        O.align_vertices(v, p["face"]["vertex_indices"], ca, cb, O.Params(iterations=iterations))
    return time.perf_counter() - t0


def reference_like_for_like():
    """The reference on the like-for-like input (files on disk -> file on disk, like its command line), and its growth with size."""
    out = {"growth": []}
    for level in (5, SAMPLE_LEVEL):
        with tempfile.TemporaryDirectory() as tmp:
            nv = reference_sample_inputs(tmp, level)
            out["growth"].append({"vertices": nv, "iterations": 10, "seconds": time_reference_once(tmp)})
    with tempfile.TemporaryDirectory() as tmp:
        nv = reference_sample_inputs(tmp, LIKE_LEVEL)
        sec = time_reference_once(tmp, LIKE_ITERATIONS)
    out["growth"].append({"vertices": nv, "iterations": LIKE_ITERATIONS, "seconds": sec})
    out["like_for_like"] = {"vertices": nv, "iterations": LIKE_ITERATIONS, "ref_s": sec}
    return out


def gpu_like_for_like(al, api):
    """The same input through this library's host-pointer API (uploads and the read-back inside the clock), best of three."""
    from meshopticalflow_b200 import synthetic
    v, t = synthetic.octahedron_sphere(LIKE_LEVEL)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    v = v.astype(np.float32).astype(np.float64)  # what a PLY file holds
    p = api.default_params()
    p.iterations = LIKE_ITERATIONS
    al.set_params(p)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        al.set_mesh(v, t)
        al.set_signals(a, b)
        al.iterate(LIKE_ITERATIONS)
        al.advect_vertices(0.5)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    al.set_params(api.default_params())
    return best


def cpu_baseline_dict(seconds: float, sample_vertices: int, target_vertices: int):
    kind = "reference" if os.path.exists(REF_BIN) else "port"
    cores = os.cpu_count() or 1
    scale = target_vertices / sample_vertices
    return {
        "value": 1.0 / (seconds * scale), "unit": UNIT, "cores": cores if kind == "reference" else 1, "kind": kind,
        "sample_seconds": seconds, "sample_vertices": sample_vertices, "extrapolated": True,
        "sample": (f"one full alignment (defaults, 10 iterations) of the level-{SAMPLE_LEVEL} sphere pair ({sample_vertices} vertices) by "
                   f"{'oracle/_ref/OpticalFlow_ref (unmodified reference, Eigen LDLT/LLT, OpenMP, MKL off)' if kind == 'reference' else 'the oracle port (scipy SuperLU)'}"
                   f": {seconds:.2f} s; value = 1/(seconds * {scale:.0f}), a LINEAR-in-vertices extrapolation to {target_vertices} vertices — the reference's sparse"
                   " Cholesky grows superlinearly (SURVEY.md §6: 16k V 1.0 s/it, 108k V 24 s/it, 262k V 83-141 s/it), so this overstates its real 1M-vertex rate"),
    }


def run_reference_arm(args, config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        nv = reference_sample_inputs(tmp)
        for _ in range(args.warmup):
            time_reference_once(tmp)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            time_reference_once(tmp)
        sec = (time.perf_counter() - t0) / args.steps
    target = config["vertices"]
    base = cpu_baseline_dict(sec, nv, target)
    if not args.quick:
        base.update(reference_like_for_like())  # once, outside the step loop
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------- the GPU arm

def partitioned_record(api, sharding, synthetic, torch, dev, stream, rank, local_rank, world):
    """One PARTITIONED_LEVEL mesh, PARTITIONED_ITERATIONS iterations: all ranks together (row-partitioned solves, NCCL halo
    exchange + all-reduce) and rank 0 alone; device time of mof_iterate (CUDA events, max over ranks)."""
    verts, tris = synthetic.octahedron_sphere(PARTITIONED_LEVEL)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(verts, 0))
    p = api.default_params()
    p.iterations = PARTITIONED_ITERATIONS
    out = {"vertices": int(verts.shape[0]), "iterations": PARTITIONED_ITERATIONS, "n_gpus": world}

    def run(al):
        al.set_params(p)
        ms = []
        for _ in range(2):  # the second run is the steady state (memory pool, NCCL channels warm)
            al.set_mesh(verts, tris)
            al.set_signals(a, b)
            torch.cuda.synchronize(dev)
            sharding.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            al.iterate(PARTITIONED_ITERATIONS)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        return ms[-1], al.stats(), al.flow()

    al = api.Aligner(local_rank, stream.cuda_stream)
    try:
        uid = sharding.broadcast_bytes(api.dist_unique_id() if rank == 0 else None, 128, 0, dev)
        al.dist_init(world, rank, uid)
        ms_local, st, flow = run(al)
    finally:
        al.close()
    ms_all = sharding.max_over_ranks(ms_local, dev)
    out.update({"ms_all_ranks": ms_all, "flow_iterations": st["flowCgIterations"] / 2, "smooth_iterations": st["smoothCgIterations"] / 2,
                "flow_solve_ms": st["flowSolveMs"] / 2, "smooth_solve_ms": st["smoothSolveMs"] / 2, "halo_entries_rank0": st["haloEntries"]})
    ms_one, flow_one = 0.0, None
    if rank == 0:
        al = api.Aligner(local_rank, stream.cuda_stream)
        try:
            os.environ["MOF_SMOOTH_AHEAD"] = "0"  # one stream on both sides of the comparison
            ms_one, st1, flow_one = run_single(al, p, verts, tris, a, b, torch, dev, stream)
        finally:
            os.environ.pop("MOF_SMOOTH_AHEAD", None)
            al.close()
        out.update({"ms_rank0_alone": ms_one, "speedup": ms_one / ms_all, "flow_rel_difference": float(np.linalg.norm(flow - flow_one) / np.linalg.norm(flow_one)),
                    "flow_iterations_rank0_alone": st1["flowCgIterations"] / 2})
    sharding.barrier()
    return out


def concurrent_record(api, torch, dev, verts, tris, pairs_h, contexts=2, pairs_each=2):
    """`contexts` solver contexts on this ONE GPU, one host thread each, aligning independent pairs at the same time from host buffers
    (the batched configuration, configs[3], per GPU): the latency-bound part of one pair's solves is filled by another's. Wall clock
    between device-wide synchronisations (several streams: no single stream's events bracket it)."""
    als = [api.Aligner(dev.index) for _ in range(contexts)]
    outs = [(np.empty((verts.shape[0], 3)), np.empty((verts.shape[0], 3))) for _ in range(contexts)]
    p = api.default_params()

    def work(k, n):
        al = als[k]
        a, b = pairs_h[k % len(pairs_h)]
        for _ in range(n):
            al.set_mesh(verts, tris)
            al.set_signals(a, b)
            al.iterate(p.iterations)
            al.advect_vertices(0.5, *outs[k])

    try:
        for k in range(contexts):
            work(k, 1)  # warm-up: pools, graphs
        torch.cuda.synchronize(dev)
        threads = [threading.Thread(target=work, args=(k, pairs_each)) for k in range(contexts)]
        t0 = time.perf_counter()
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    finally:
        for al in als:
            al.close()
    return {"contexts_per_gpu": contexts, "alignments": contexts * pairs_each, "seconds": dt, "value": contexts * pairs_each / dt, "unit": UNIT,
            "what": "independent pairs aligned at the same time by several contexts on one GPU, host buffers in and out, wall clock; `value` above is one pair at a time"}


def run_single(al, p, verts, tris, a, b, torch, dev, stream):
    al.set_params(p)
    ms = []
    for _ in range(2):
        al.set_mesh(verts, tris)
        al.set_signals(a, b)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        al.iterate(p.iterations)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    return ms[-1], al.stats(), al.flow()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--level", type=int, default=9, help="octahedron subdivision level of the workload (9 = 1 048 578 vertices)")
    ap.add_argument("--partitioned", action="store_true",
                    help="configs[4]: ONE pair per step for the whole job, its flow solves row-partitioned over the ranks (NCCL halo exchange + all-reduce); strong scaling")
    ap.add_argument("--quick", action="store_true", help="skip the like-for-like reference timing (about a minute of CPU work) — for the CPU-tier test of the contract line")
    args = ap.parse_args()

    V = 4 * 4 ** args.level + 2
    config = {"workload": f"synthetic subdivided-octahedron sphere, {V} vertices / {2 * V - 4} triangles / {3 * V - 6} Whitney unknowns, vertices and triangles "
                          "numbered along a Morton curve by the generator (outside the timed region; the library's own renumbering of badly ordered meshes, mof_set_reorder, measures it as local and leaves it), smooth random RGB "
                          "per-vertex signals (B = A rotated 4 deg), reference defaults (10 iterations), one pair per GPU per step",
              "numbering": "Morton-sorted (synthetic.octahedron_sphere spatial_sort=True)",
              "vertices": V, "pairs_per_step": 1 if args.partitioned else args.gpus,
              "parallelism": (f"one mesh, flow solves row-partitioned x{args.gpus} (NCCL halo exchange + all-reduce), everything else replicated" if args.partitioned
                              else f"independent pairs x{args.gpus} (no communication)"),
              "l2": "inputs larger than L2: each step streams > 1 GB of operators per PCG iteration; no explicit flush", "pcg_tol": 1e-8,
              "streams": ("one" if os.environ.get("MOF_SMOOTH_AHEAD", "1") == "0" or args.partitioned
                          else "two: the next iteration's smoothing solve runs under the flow solve (solve times below overlap)")}
    if args.impl == "reference":
        return run_reference_arm(args, config)

    import torch

    from meshopticalflow_b200 import api, sharding, synthetic

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device — the GPU arm has no CPU fallback", file=sys.stderr)
        return 2
    rank, local_rank, world = sharding.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init_process_group("nccl")

    verts, tris = synthetic.octahedron_sphere(args.level)
    T = tris.shape[0]

    def pair(seed):
        a, b = synthetic.smooth_rgb_pair(verts, seed)
        return a.astype(np.float64), b.astype(np.float64)

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        al = api.Aligner(local_rank, stream.cuda_stream)
        if args.partitioned:
            uid = sharding.broadcast_bytes(api.dist_unique_id() if rank == 0 else None, 128, 0, dev)
            al.dist_init(world, rank, uid)
        params = api.default_params()
        al.set_params(params)
        d_v, d_t = torch.from_numpy(verts).to(dev), torch.from_numpy(tris).to(dev)
        d_oa, d_ob = torch.empty((V, 3), dtype=torch.float64, device=dev), torch.empty((V, 3), dtype=torch.float64, device=dev)
        # a few resident pairs, cycled (a new seed every step and every rank)
        n_sets = 2
        pairs_h = [pair(k if args.partitioned else rank + world * k) for k in range(n_sets)]  # partitioned: every rank holds the same pair
        pairs_d = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in pairs_h]

        def step_resident(k):
            da, db = pairs_d[k % n_sets]
            al.set_mesh_device(d_v.data_ptr(), V, d_t.data_ptr(), T)
            al.set_signals_device(da.data_ptr(), db.data_ptr(), 3)
            al.iterate(params.iterations)
            al.advect_vertices_device(0.5, d_oa.data_ptr(), d_ob.data_ptr())

        # pinned host buffers for the end-to-end arm
        h_v, h_t = torch.from_numpy(verts).pin_memory(), torch.from_numpy(tris).pin_memory()
        h_pairs = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs_h]
        h_oa, h_ob = torch.empty((V, 3), dtype=torch.float64).pin_memory(), torch.empty((V, 3), dtype=torch.float64).pin_memory()

        def step_e2e(k):
            ha, hb = h_pairs[k % n_sets]
            al.set_mesh(h_v.numpy(), h_t.numpy())
            al.set_signals(ha.numpy(), hb.numpy())
            al.iterate(params.iterations)
            al.advect_vertices(0.5, h_oa.numpy(), h_ob.numpy())
            return float(h_oa[0, 0])

        for k in range(args.warmup):
            step_resident(k)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        al.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(args.steps):
            step_resident(args.warmup + k)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        ms_local = e0.elapsed_time(e1)
        stats = al.stats()
        clocks = sampler.stop() if rank == 0 else None
        ms = sharding.max_over_ranks(ms_local, dev)

        # roofline of the dominant kernel on this workload's own flow matrix
        spmv_ms = al.time_flow_spmv(50)
        spmv_bytes = stats["flowSpmvBytes"]

        # end to end
        step_e2e(0)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        al.reset_stats()
        e0.record(stream)
        for k in range(args.steps):
            step_e2e(1 + k)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        sharding.barrier()
        e2e_ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
        e2e_stats = al.stats()

        # every large kernel of a PCG iteration and the walk, each alone (rank 0; not on a partitioned mesh)
        kernel_rows = []
        if rank == 0 and not args.partitioned:
            for name, which in api.KERNELS.items():
                try:
                    us, nbytes = al.time_kernel(which, 30)
                    kernel_rows.append((name, us, nbytes))
                except api.MofError as e:
                    kernel_rows.append((name, None, str(e)))
        like_gpu_s = gpu_like_for_like(al, api) if (rank == 0 and world == 1 and not args.partitioned and not args.quick) else None
        al.close()
        concurrent = concurrent_record(api, torch, dev, verts, tris, pairs_h) if (world == 1 and not args.partitioned and not args.quick) else None

        # N > 1: BASELINE.json configs[4] next to the sharded pairs — one mesh, its solves row-partitioned over all ranks, against rank 0 alone
        partitioned = None
        if world > 1 and not args.partitioned:
            partitioned = partitioned_record(api, sharding, synthetic, torch, dev, stream, rank, local_rank, world)
    sharding.shutdown()

    if rank != 0:
        return 0
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_spmv_dot_dram_bytes_per_launch")
        except Exception:
            traffic = None
    peak, peak_src = measured_peak_gbs()
    iters = max(stats["flowCgIterations"], 1)
    jobs = 1 if args.partitioned else world  # alignments completed per step by the whole job
    # per-kernel table: launches per step from the iteration counts (per PCG iteration: one matrix product, two fine sweeps — the
    # cycle's residual and its post-smoothing —, one update, one restriction, one prolongation, one new direction, three level-1
    # stencil applications — two residuals of the W visit and the post-smoothing; per UpdateFlow iteration: one walk launch)
    per_iter = {"flow_spmv": 1, "flow_fine_sweep": 2, "flow_update": 1, "flow_restrict": 1, "flow_prolong": 1, "flow_direction": 1, "flow_level1": 3,
                "scalar_spmv": 1, "scalar_fine_sweep": 2, "scalar_update": 1, "scalar_level1": 3}
    flow_it, smooth_it = stats["flowCgIterations"] / args.steps, stats["smoothCgIterations"] / args.steps
    step_us = ms / args.steps * 1e3
    kernels, flow_stream_bytes, flow_stream_us = [], 0.0, 0.0
    for name, us, nbytes in kernel_rows:
        if us is None:
            kernels.append({"name": name, "unavailable": nbytes})
            continue
        count = (per_iter.get(name, 0) * (flow_it if name.startswith("flow") else smooth_it)) if name != "walk" else float(params.iterations)
        row = {"name": name, "us": us, "launches_per_step": count, "share_of_step": count * us / step_us}
        if nbytes > 0:
            gbs = nbytes / (us * 1e-6) / 1e9
            row.update({"algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / peak})
        else:
            row.update({"algorithmic_bytes": None, "note": "latency / divergence bound: trip counts are data dependent"})
        kernels.append(row)
        if name.startswith("flow") and nbytes > 0:
            flow_stream_bytes += per_iter[name] * nbytes
            flow_stream_us += per_iter[name] * us
    us_per_flow_it = stats["flowSolveMs"] * 1e3 / iters
    line = {
        "metric": METRIC, "value": jobs * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.partitioned else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "clocks": clocks,
        "e2e": {"value": jobs * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(V * 3 * 8 + T * 3 * 4 + 2 * V * 3 * 8),
                "d2h_bytes_per_step": int(2 * V * 3 * 8), "ms_per_step": e2e_ms / args.steps,
                "flow_solve_ms_per_alignment": e2e_stats["flowSolveMs"] / args.steps, "smooth_solve_ms_per_alignment": e2e_stats["smoothSolveMs"] / args.steps,
                "setup_ms_per_alignment": e2e_stats["setupMs"] / args.steps, "advect_ms_per_alignment": e2e_stats["advectMs"] / args.steps,
                "flow_iterations_per_alignment": e2e_stats["flowCgIterations"] / args.steps},
        "gpu_launches": int(stats["kernelLaunches"]),
        "roofline": {"bound": "hbm", "kernel": "k_spmv_dot (the fp64 SpMV of the flow PCG, y = A d fused with d.y; flow system in the sliced SELL-32 layout, fp64 values / int32 columns)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "traffic": traffic, "bytes_per_launch": spmv_bytes, "us_per_launch": spmv_ms * 1e3, "rows": stats["flowRows"], "nnz": stats["flowNnz"],
                     "kernels": kernels,
                     "flow_iteration_bytes": flow_stream_bytes, "flow_iteration_us_at_peak": flow_stream_bytes / (peak * 1e9) * 1e6,
                     "flow_iteration_us_streaming_kernels": flow_stream_us, "flow_iteration_us_measured": us_per_flow_it,
                     "flow_iteration_frac": (flow_stream_bytes / (peak * 1e9) * 1e6) / us_per_flow_it if flow_stream_bytes else None,
                     "flow_iteration_note": ("bytes = the algorithmic bytes of the kernels listed with a flow_ prefix (fine level and the largest coarse level); the rest of an "
                                             "iteration is ~40 dependent launches of 4-9 us on the smaller multigrid levels (latency bound), see DESIGN.md §4.2; "
                                             "with two streams the measured time also contains whatever of the smoothing solve overlapped it")},
        "pcg": {"flow_iterations_per_alignment": stats["flowCgIterations"] / args.steps, "smooth_iterations_per_alignment": stats["smoothCgIterations"] / args.steps,
                "flow_solve_ms_per_alignment": stats["flowSolveMs"] / args.steps, "smooth_solve_ms_per_alignment": stats["smoothSolveMs"] / args.steps,
                "us_per_flow_iteration": stats["flowSolveMs"] * 1e3 / iters, "setup_ms_per_alignment": stats["setupMs"] / args.steps,
                "advect_ms_per_alignment": stats["advectMs"] / args.steps, "halo_entries_rank0": stats["haloEntries"]},
    }
    if partitioned is not None:
        line["partitioned"] = partitioned
    if concurrent is not None:
        line["concurrent_contexts"] = concurrent
    if world == 1 and not args.partitioned:
        with tempfile.TemporaryDirectory() as tmp:
            nv = reference_sample_inputs(tmp)
            sec = time_reference_once(tmp)
        base = cpu_baseline_dict(sec, nv, V)
        if not args.quick:
            base.update(reference_like_for_like())
        if like_gpu_s and "like_for_like" in base:
            lfl = base["like_for_like"]
            lfl.update({"gpu_s": like_gpu_s, "ratio": lfl["ref_s"] / like_gpu_s,
                        "what": "the same 65 538-vertex pair and iteration count through both: the reference binary from files to file (wall), this library from host "
                                "buffers to host buffers (wall, uploads and read-back inside); not extrapolated"})
        line["cpu_baseline"] = base
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
