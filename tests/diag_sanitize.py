"""Diagnostic (GPU, under compute-sanitizer): a small alignment with a forced renumbering, the Spectrum eigen-solve for the Whitney and
Connection bases, on meshes small enough for memcheck / racecheck to finish in a minute.
    compute-sanitizer --tool memcheck python tests/diag_sanitize.py     (where the tool is available: it is closed on the pool this round ran on;
    the CPU tier runs the same sources under AddressSanitizer / UBSan instead, tests/host_emulation/run_sanitizers.sh)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    os.environ["MOF_SMOOTH_AHEAD"] = "0"
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    v, t = synthetic.octahedron_sphere(level)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    al = api.Aligner(0)
    al.set_reorder(1)
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.iterate(2)
    ca, cb = al.advect_vertices(0.5)
    print("alignment (renumbered):", float(np.abs(al.flow()).max()), float(ca.mean()))
    ev, _, its, res = al.spectrum(6, 1e-8, 2000)
    print("spectrum Whitney:", its, res, ev[:3])
    p = api.default_params()
    p.vfMode = 2
    al.set_params(p)
    al.set_mesh(v, t)
    ev, _, its, res = al.spectrum(6, 1e-6, 300)
    print("spectrum Connection:", its, res, ev[:3])
    al.close()


if __name__ == "__main__":
    main()
