"""Diagnostic (GPU): the per-kernel table of bench.py's roofline.kernels on its own — every large kernel of a PCG iteration and the
walk, each launched alone on the operators of a synthetic sphere pair (mof_time_kernel).  python tests/diag_kernels.py [level] [reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    peak = 6549.4
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["hbm_gbs"])
    v, t = synthetic.octahedron_sphere(level)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    al = api.Aligner(0)
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.iterate(2)
    print(f"{v.shape[0]} vertices; peak {peak:.0f} GB/s")
    for name, which in api.KERNELS.items():
        us, nbytes = al.time_kernel(which, reps)
        if nbytes > 0:
            gbs = nbytes / us * 1e-3
            print(f"{name:20s} {us:9.2f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.0f} GB/s  {gbs / peak:5.2f} of peak")
        else:
            print(f"{name:20s} {us:9.2f} us")
    al.close()


if __name__ == "__main__":
    main()
