"""Diagnostic (CPU, no GPU): a whole alignment of a synthetic sphere pair on the EMULATED build of the library
(tests/host_emulation: every .cu file compiled for the host), at any size, optionally against the checker's direct solves.

    python tests/diag_emulated_alignment.py LEVEL ITERATIONS [--oracle] [--vfMode M] [--cxxflags "-O1 -g -fsanitize=address"]

LEVEL 8 (262 146 vertices; five-level hierarchies, the 9-warp stencil kernel of the large levels) takes about 3 minutes per
iteration here; with --cxxflags "-O1 -g -fsanitize=address" run it under LD_PRELOAD=$(g++ -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0. Prints iteration counts, residuals and the flow's distance to the checker."""
import argparse
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshopticalflow_b200 import api, synthetic  # noqa: E402

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")


def build(out, extra):
    base = ["g++"] + extra + ["-std=c++17", "-fPIC", "-c", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-I.", "-w"]
    jobs = [base + ["-DEMUL_UNIT=%d" % u, "-o", os.path.join(out, "unit%d.o" % u), "library_emul.cpp"] for u in range(7)]
    jobs += [base + ["-o", os.path.join(out, "dist_stub.o"), "dist_stub.cpp"], base + ["-o", os.path.join(out, "runtime.o"), "emul_runtime.cpp"]]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda cmd: subprocess.check_call(cmd, cwd=EMU_DIR), jobs))
    lib = os.path.join(out, "libmof_emul.so")
    subprocess.check_call(["g++", "-shared"] + [f for f in extra if f.startswith("-fsanitize")] + ["-o", lib] + [j[j.index("-o") + 1] for j in jobs] + ["-lpthread"])
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("level", type=int)
    ap.add_argument("iterations", type=int)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--vfMode", type=int, default=0)
    ap.add_argument("--cxxflags", default="-O2")
    args = ap.parse_args()
    os.environ["MOF_SMOOTH_AHEAD"] = "0"  # the emulator's thread/block registers are per OS thread only in the MOF_EMUL_THREADS build
    with tempfile.TemporaryDirectory() as d:
        api.LIB_PATH, api._lib = build(d, args.cxxflags.split()), None
        v, t = synthetic.octahedron_sphere(args.level)
        a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
        p = api.default_params()
        p.iterations, p.vfMode, p.vfSmooth = args.iterations, args.vfMode, (3e-6, 5e-7, 1e4)[args.vfMode]
        t0 = time.time()
        out, flow, stats = api.align_vertices(v, t, a, b, p)
        print(f"{v.shape[0]} vertices, {args.iterations} iterations on the emulated build: {time.time() - t0:.1f} s;",
              {k: stats[k] for k in ("kernelLaunches", "flowCgIterations", "smoothCgIterations", "lastFlowResidual", "lastSmoothResidual")}, flush=True)
        if args.oracle:
            from oracle import mof_oracle as O
            t0 = time.time()
            st, ref = O.align_vertices(v, t, a, b, O.Params(iterations=args.iterations, vfMode=args.vfMode))
            print(f"checker: {time.time() - t0:.1f} s; flow rel-L2 {np.linalg.norm(flow - st.tfield) / np.linalg.norm(st.tfield):.3e}, "
                  f"max colour difference {np.abs(out - ref).max():.3e}")


if __name__ == "__main__":
    main()
