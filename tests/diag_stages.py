"""Stage-by-stage diff of the CUDA path against the CPU oracle on a synthetic sphere (GPU needed).

Not a pytest file: a diagnostic that prints one line per stage so that one GPU run shows where a
divergence starts.  `python tests/diag_stages.py [level] [iterations]`
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshopticalflow_b200 import api, synthetic  # noqa: E402
from oracle import mof_oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    v, t = synthetic.octahedron_sphere(level)
    ca, cb = synthetic.smooth_rgb_pair(v, 0)
    ca, cb = ca.astype(np.float64), cb.astype(np.float64)
    print(f"level {level}: V={v.shape[0]} T={t.shape[0]}")

    params = O.Params()
    st = O.init(v, t, ca, cb, params)

    al = api.Aligner(0)
    t0 = time.time()
    al.set_mesh(v, t)
    print("set_mesh ok %.3fs, E=%d" % (time.time() - t0, al.num_edges))
    print("metric      ", rel(al.array(api.ARR_METRIC), st.g))
    print("area        ", rel(al.array(api.ARR_AREA), st.area))
    print("opposite    ", np.array_equal(al.array(api.ARR_OPPOSITE), st.opp))
    print("xform lin   ", rel(al.array(api.ARR_XFORM_LINEAR), st.lin))
    print("xform cst   ", rel(al.array(api.ARR_XFORM_CONSTANT), st.cst))
    for name, which, ref in (("sMass", api.CSR_SCALAR_MASS, st.M), ("sStiffness", api.CSR_SCALAR_STIFFNESS, st.S), ("whitney S", api.CSR_WHITNEY_SMOOTH, st.whitney.S)):
        m = al.csr(which)
        same = np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
        print(f"{name:12s} pattern {same} nnz {m.nnz}/{ref.nnz} values {rel(m.data, ref.data) if same else float('nan')}")
    print("reduced     ", np.array_equal(al.array(api.ARR_REDUCED_EDGE), st.whitney.reduced))
    print("expanded    ", np.array_equal(al.array(api.ARR_EXPANDED_EDGE), st.whitney.expanded))
    print("positive    ", np.array_equal(al.array(api.ARR_POSITIVE_EDGE), st.whitney.positive.astype(np.int32)))
    Pd = st.whitney.P.toarray() if st.whitney.P.shape[0] < 5000 else None
    P = al.array(api.ARR_PROLONGATION).reshape(-1, 3, 2)
    Pref = np.zeros_like(P)
    red = st.whitney.reduced.reshape(-1, 3)
    Pcsr = st.whitney.P.tocsr()
    for k in range(3):
        for r in range(2):
            rows = 2 * np.arange(P.shape[0]) + r
            Pref[:, k, r] = np.asarray(Pcsr[rows, red[:, k]]).reshape(-1)
    print("prolongation", rel(P, Pref))

    t0 = time.time()
    al.set_signals(ca, cb)
    print("set_signals ok %.3fs" % (time.time() - t0))
    sig = al.array(api.ARR_SIGNALS)
    print("signals(DoG)", rel(sig[:, :3], st.signals[0]), rel(sig[:, 3:], st.signals[1]))

    sw, vw = params.sSmooth, params.vfSmooth
    for i in range(iters):
        t0 = time.time()
        al.iterate(1)
        dt = time.time() - t0
        O.update_flow(st, sw, vw, "it.")
        sw *= params.sMultiply
        sm, rs = al.array(api.ARR_SMOOTHED), al.array(api.ARR_RESAMPLED)
        A = al.csr(api.CSR_FLOW_SYSTEM)
        print(f"it{i} ({dt:.3f}s) smoothed {rel(sm[:, :3], st.taps['it.smoothed0']):.2e} {rel(sm[:, 3:], st.taps['it.smoothed1']):.2e}"
              f" resampled {rel(rs[:, :3], st.taps['it.resampled0']):.2e} {rel(rs[:, 3:], st.taps['it.resampled1']):.2e}"
              f" D {rel(al.array(api.ARR_DATA_TERM), st.taps['it.dataTerm']):.2e} rhs {rel(al.array(api.ARR_DATA_RHS), st.taps['it.rhs']):.2e}"
              f" x {rel(al.array(api.ARR_FLOW_SOLUTION), st.taps['it.x']):.2e} coeffs {rel(al.coeffs(), st.taps['it.coeffs']):.2e}"
              f" flow {rel(al.flow(), st.taps['it.tFlowField']):.2e} |A| {abs(A).sum():.6e}")
    a, b = al.advect_vertices(0.5)
    oa, ob = O.advect_vertices(st, ca, cb)
    print("advected max|diff|", np.abs(a - oa).max(), np.abs(b - ob).max())
    print("stats", al.stats())
    al.close()


if __name__ == "__main__":
    main()
