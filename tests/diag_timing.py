"""Timing/iteration-count probe of the CUDA path on a synthetic sphere (GPU needed; no oracle).
`python tests/diag_timing.py [level] [iterations] [vfMode] [cMode] [dogWeight]`"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    t0 = time.time()
    numbering = os.environ.get("MOF_SYNTH_NUMBERING", "morton")  # morton | subdivision (the generator's own order) | random
    v, t = synthetic.octahedron_sphere(level, spatial_sort=numbering == "morton")
    if numbering == "random":
        rng = np.random.default_rng(1)
        order = rng.permutation(v.shape[0])
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        v, t = np.ascontiguousarray(v[order]), np.ascontiguousarray(rank[t][rng.permutation(t.shape[0])].astype(np.int32))
    ca, cb = synthetic.smooth_rgb_pair(v, 0)
    ca, cb = ca.astype(np.float64), cb.astype(np.float64)
    print(f"level {level}: V={v.shape[0]} T={t.shape[0]} numbering {numbering} (generated in {time.time() - t0:.1f}s)", flush=True)
    al = api.Aligner(0)
    vf_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p = api.default_params()
    p.vfMode, p.cMode = vf_mode, int(sys.argv[4]) if len(sys.argv) > 4 else 0
    p.dogWeight = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    p.vfSmooth = 0.0  # the mode's default
    al.set_params(p)
    print(f"vfMode {p.vfMode} cMode {p.cMode} dogWeight {p.dogWeight}", flush=True)
    t0 = time.time()
    al.set_mesh(v, t)
    print("set_mesh %.3fs E=%d" % (time.time() - t0, al.num_edges), flush=True)
    t0 = time.time()
    al.set_mesh(v, t)
    print("set_mesh again (steady state) %.3fs, setupMs so far %.1f" % (time.time() - t0, al.stats()["setupMs"]), flush=True)
    t0 = time.time()
    al.set_signals(ca, cb)
    s = al.stats()
    print("set_signals %.3fs smoothIters=%d" % (time.time() - t0, s["smoothCgIterations"]), flush=True)
    prev = s
    for i in range(iters):
        t0 = time.time()
        al.iterate(1)
        s = al.stats()
        print(f"it{i} {time.time() - t0:.3f}s flowIters {s['flowCgIterations'] - prev['flowCgIterations']} ({s['flowSolveMs'] - prev['flowSolveMs']:.1f} ms)"
              f" smoothIters {s['smoothCgIterations'] - prev['smoothCgIterations']} ({s['smoothSolveMs'] - prev['smoothSolveMs']:.1f} ms)"
              f" advect {s['advectMs'] - prev['advectMs']:.2f} ms relres {s['lastFlowResidual']:.2e} |flow| {np.abs(al.flow()).max():.4f}", flush=True)
        prev = s
    t0 = time.time()
    a, b = al.advect_vertices(0.5)
    print("final advect %.3fs; mean|A-B| before %.3f after %.3f" % (time.time() - t0, np.abs(ca - cb).mean(), np.abs(a - b).mean()))
    s = al.stats()
    if vf_mode != 0:
        print("unknowns", al.num_coeffs, "stats", s)
        al.close()
        return
    ms = al.time_flow_spmv(50)
    print(f"flow SpMV {ms * 1e3:.1f} us, {s['flowSpmvBytes'] / ms / 1e6:.1f} GB/s algorithmic ({s['flowRows']} rows, {s['flowNnz']} nnz)")
    print("stats", s)
    al.close()


if __name__ == "__main__":
    main()
