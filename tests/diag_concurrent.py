"""Diagnostic (GPU): K contexts on ONE GPU aligning K independent pairs at the same time (one host thread each) against one at a time —
how much of the latency-bound part of a solve another pair's work can fill.  python tests/diag_concurrent.py [level] [contexts] [pairs]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    v, t = synthetic.octahedron_sphere(level)
    sig = [tuple(x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, s)) for s in range(K)]
    als = [api.Aligner(0) for _ in range(K)]

    def work(k, n):
        al = als[k]
        for _ in range(n):
            al.set_mesh(v, t)
            al.set_signals(*sig[k])
            al.iterate(10)
            al.advect_vertices(0.5)

    work(0, 1)  # warm-up
    for k in range(1, K):
        work(k, 1)
    t0 = time.perf_counter()
    work(0, pairs)
    one = (time.perf_counter() - t0) / pairs
    print(f"one context: {one * 1e3:.1f} ms per alignment, {1 / one:.3f} alignments/s (host buffers in and out)", flush=True)
    threads = [threading.Thread(target=work, args=(k, pairs)) for k in range(K)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    print(f"{K} contexts at once: {K * pairs} alignments in {dt:.3f} s = {K * pairs / dt:.3f} alignments/s ({K * pairs / dt * one:.2f}x)")
    for al in als:
        al.close()


if __name__ == "__main__":
    main()
