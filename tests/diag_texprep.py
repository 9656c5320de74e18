"""Diagnostic (GPU box): the texture configuration's one-time preparation, host (csrc/host/texture_prep.cpp) against device
(csrc/texprep_kernels.cu), on an input of the reference Example's size — a 20 000-vertex uv torus subdivided at
--eLength 0.006 (≈ 10^5 vertices) under two 388 x 388 textures.

    python tests/diag_texprep.py [nu nv [size]]

Prints the stage times of the device preparation through the C ABI, and the wall time of the whole command line with
MOF_GPU_TEXPREP=0 and =1 (same output picture expected; the difference is the preparation)."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    nu, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 100)
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 388
    from PIL import Image
    v, t, uv = synthetic.uv_torus(nu, nv)
    ta, tb = synthetic.smooth_texture_pair(size, size, 1)
    lo, hi = v.astype(np.float64).min(0), v.astype(np.float64).max(0)
    e_len = float(np.float32(np.float32(0.006) * float(np.sqrt(((hi - lo) ** 2).sum()))))
    al = api.Aligner(0)
    for attempt in ("cold", "warm"):
        t0 = time.perf_counter()
        sv, st, suv = al.subdivide(v, t, uv, e_len)
        t1 = time.perf_counter()
        al.set_mesh(sv.astype(np.float64), st)
        t2 = time.perf_counter()
        srcT, _ = al.build_texture_map(size, size, 2, suv, ta, tb)
        t3 = time.perf_counter()
        al.sample_textures_to_vertices()
        t4 = time.perf_counter()
        print(f"device preparation ({attempt}): {sv.shape[0]} vertices / {st.shape[0]} triangles, {int((srcT >= 0).sum())} of {size * size} texels covered; "
              f"subdivide {1e3 * (t1 - t0):.1f} ms, set_mesh {1e3 * (t2 - t1):.1f} ms, texel map {1e3 * (t3 - t2):.1f} ms, vertex colours {1e3 * (t4 - t3):.1f} ms")
    al.close()
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_textured(os.path.join(d, "m.ply"), v, t, uv)
        Image.fromarray(ta).save(os.path.join(d, "A.png"))
        Image.fromarray(tb).save(os.path.join(d, "B.png"))
        cli = os.path.join(ROOT, "meshopticalflow_b200", "OpticalFlow")
        # the host restatement is test infrastructure: a checker binary with it compiled in (MOF_GPU_TEXPREP=0 selects it there)
        host, libdir = os.path.join(ROOT, "meshopticalflow_b200", "csrc", "host"), os.path.join(ROOT, "meshopticalflow_b200")
        checker = os.path.join(d, "OpticalFlow_hostprep")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMOF_WITH_HOST_TEXPREP", "-I" + os.path.join(ROOT, "include"), "-o", checker] +
                              [os.path.join(host, f) for f in ("optical_flow_main.cpp", "ply_io.cpp", "png_codec.cpp", "texture_prep.cpp", "cmdline.cpp")] +
                              ["-L" + libdir, "-lmof_b200", "-lz", "-Wl,-rpath," + libdir])
        pictures = []
        for mode in ("0", "1", "0", "1"):
            t0 = time.perf_counter()
            subprocess.check_call([checker if mode == "0" else cli, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "r%s.png" % mode], cwd=d, stdout=subprocess.DEVNULL,
                                  env=dict(os.environ, MOF_GPU_TEXPREP=mode))
            print(f"command line, MOF_GPU_TEXPREP={mode}: {time.perf_counter() - t0:.2f} s wall")
            pictures.append(np.asarray(Image.open(os.path.join(d, "r%s.png" % mode))).astype(int))
        print("pictures differ by at most", int(np.abs(pictures[0] - pictures[1]).max()), "of 255")


if __name__ == "__main__":
    main()
