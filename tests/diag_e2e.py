"""Per-call wall-clock of the host-pointer path against the device-pointer path (GPU needed; no oracle).
`python tests/diag_e2e.py [level] [steps]`"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    v, t = synthetic.octahedron_sphere(level)
    print("dtypes", v.dtype, t.dtype)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    V, T = v.shape[0], t.shape[0]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        al = api.Aligner(0, stream.cuda_stream)
        d_v, d_t, d_a, d_b = (torch.from_numpy(x).to(dev) for x in (v, t, a, b))
        d_oa, d_ob = torch.empty((V, 3), dtype=torch.float64, device=dev), torch.empty((V, 3), dtype=torch.float64, device=dev)
        h_v, h_t, h_a, h_b = (torch.from_numpy(x).pin_memory() for x in (v, t, a, b))
        h_oa, h_ob = torch.empty((V, 3), dtype=torch.float64).pin_memory(), torch.empty((V, 3), dtype=torch.float64).pin_memory()

        def timed(label, fn):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            return f"{label} {1e3 * (time.perf_counter() - t0):7.1f} ms"

        for mode in ("device", "host", "device", "host"):
            for k in range(steps):
                al.reset_stats()
                if mode == "device":
                    parts = [timed("mesh", lambda: al.set_mesh_device(d_v.data_ptr(), V, d_t.data_ptr(), T)),
                             timed("signals", lambda: al.set_signals_device(d_a.data_ptr(), d_b.data_ptr(), 3)),
                             timed("iterate", lambda: al.iterate(10)),
                             timed("advect", lambda: al.advect_vertices_device(0.5, d_oa.data_ptr(), d_ob.data_ptr()))]
                else:
                    parts = [timed("mesh", lambda: al.set_mesh(h_v.numpy(), h_t.numpy())),
                             timed("signals", lambda: al.set_signals(h_a.numpy(), h_b.numpy())),
                             timed("iterate", lambda: al.iterate(10)),
                             timed("advect", lambda: al.advect_vertices(0.5, h_oa.numpy(), h_ob.numpy()))]
                s = al.stats()
                print(f"{mode:6s} step {k}: " + " | ".join(parts) + f" | setupMs {s['setupMs']:.1f} smooth {s['smoothSolveMs']:.1f} flow {s['flowSolveMs']:.1f}", flush=True)
        al.close()


if __name__ == "__main__":
    main()
