"""CPU tier: meshopticalflow_b200/csrc/vector_fields.cu (Conformal / Connection bases: kernels and the matrix-free PCG
driver) compiled for the HOST by tests/host_emulation — the real source, CUDA runtime calls and launch syntax
replaced by stand-ins, thread blocks run on fibers — and checked against the numpy checker and the reference's
golden flows. This is how the new CUDA file is exercised where there is no GPU; the GPU tier (tests/test_gpu_modes.py)
runs the same checks on the device build."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, VF_MODES, rel
from oracle import mof_oracle as O

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
_D, _I = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("vf_emul") / "libvf_emul.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-I.", "-w", "-o", out, "vf_emul.cpp",
                           "emul_runtime.cpp"], cwd=EMU_DIR)
    return ctypes.CDLL(out)


def _half_edge_rows(S, tri):
    """The product's sHe: for CSR entry (a, b) the half-edge a -> b (h = 3t + j, a = corner j+1, b = corner j+2), -1 on the diagonal."""
    nv = S.shape[0]
    T = tri.shape[0]
    j = np.arange(3)
    a, b = tri[:, (j + 1) % 3].reshape(-1).astype(np.int64), tri[:, (j + 2) % 3].reshape(-1).astype(np.int64)
    h = (3 * np.arange(T)[:, None] + j[None]).reshape(-1)
    lookup = dict(zip((a * nv + b).tolist(), h.tolist()))
    rows = np.repeat(np.arange(nv), np.diff(S.indptr))
    return np.array([lookup.get(int(r) * nv + int(c), -1) for r, c in zip(rows, S.indices)], dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.mark.parametrize("name,hierarchy", [(n, 0) for n in sorted(VF_MODES)] + [("conformal", 1)])
def test_vector_field_source_on_the_host(emul, golden_modes, name, hierarchy, monkeypatch):
    """hierarchy = 1: the Conformal PCG preconditioned by two "cycles" of the scalar hierarchy around the lumped mass (here a
    dense solve of M + eps K rounded to fp32 stands in for a cycle) instead of block Jacobi."""
    # this harness' stand-in for a cycle is a host function, not launches: it cannot be captured into the PCG's batch graph
    # (the whole-library emulation, tests/test_library_host_emulation.py, runs the same solves WITH the graph)
    monkeypatch.setenv("MOF_VF_GRAPH", "0")
    g = golden_modes
    vf_mode, c_mode = VF_MODES[name]
    v = g["input_vertices_f32"].astype(np.float64)
    t = g["triangles"]
    a, b = g["input_a"].astype(np.float64), g["input_b"].astype(np.float64)
    steps = 3
    params = O.Params(iterations=steps, vfMode=vf_mode, cMode=c_mode)
    st = O.init(v, t, a, b, params)
    O.iterate(st, params, taps=True)
    V, T = v.shape[0], t.shape[0]
    N = 2 * V if vf_mode == 1 else 2 * T
    S = st.S
    he = _half_edge_rows(S, t)
    assert (he < 0).sum() == V
    m0 = np.zeros(V)
    np.add.at(m0, t.reshape(-1), np.repeat(st.area / 3.0, 3))
    D = np.ascontiguousarray(np.stack([st.taps["it%02d.dataTerm" % i] for i in range(steps)]))
    rhs = np.ascontiguousarray(np.stack([st.taps["it%02d.rhs" % i] for i in range(steps)]))
    outB, outX, outC = np.zeros((steps, N)), np.zeros((steps, N)), np.zeros((steps, N))
    outF, outS = np.zeros((steps, T, 2)), np.zeros(steps)
    iters, relres, cycles = ctypes.c_longlong(), ctypes.c_double(), ctypes.c_int()
    Mc = st.M
    assert np.array_equal(Mc.indptr, S.indptr) and np.array_equal(Mc.indices, S.indices)
    arrs = dict(g=np.ascontiguousarray(st.g), area=np.ascontiguousarray(st.area), opp=np.ascontiguousarray(st.opp, dtype=np.int32), lin=np.ascontiguousarray(st.lin),
                cst=np.ascontiguousarray(st.cst), tri=np.ascontiguousarray(t, dtype=np.int32), rp=S.indptr.astype(np.int32), col=S.indices.astype(np.int32),
                val=np.ascontiguousarray(S.data))
    rc = emul.emul_vf_run(V, T, _p(arrs["g"], _D), _p(arrs["area"], _D), _p(arrs["opp"], _I), _p(arrs["lin"], _D), _p(arrs["cst"], _D), _p(arrs["tri"], _I),
                          _p(arrs["rp"], _I), _p(arrs["col"], _I), _p(he, _I), _p(arrs["val"], _D), _p(np.ascontiguousarray(Mc.data), _D), hierarchy, _p(m0, _D), vf_mode, c_mode, ctypes.c_double(params.vfSmooth),
                          ctypes.c_double(1e-8), steps, _p(D, _D), _p(rhs, _D), _p(outB, _D), _p(outX, _D), _p(outF, _D), _p(outS, _D), _p(outC, _D),
                          ctypes.byref(iters), ctypes.byref(relres), ctypes.byref(cycles))
    assert rc == 0
    assert relres.value <= 1e-8 and iters.value > 0
    if hierarchy:  # two cycles per iteration (+ the start of each solve), and far fewer iterations than block Jacobi needs (~470 per solve here)
        assert cycles.value >= 2 * iters.value and iters.value <= 3 * 175, (cycles.value, iters.value)
    else:
        assert cycles.value == 0
    P = st.whitney.P
    for i in range(steps):
        A, bvec, Dt, scale = O.flow_system(st.whitney, D[i], rhs[i], params.vfSmooth)
        assert abs(outS[i] - scale) <= 1e-12 * scale                       # 1 / ||R D P||_F
        assert rel(outB[i], bvec) < 1e-12                                  # s R rhs
        assert np.linalg.norm(A @ outX[i] - bvec) <= 1.01e-8 * np.linalg.norm(bvec)  # the matrix-free operator IS the assembled matrix
        assert rel(P @ outX[i], P @ st.taps["it%02d.x" % i]) < 1e-5
        assert rel(outF[i], st.taps["it%02d.tFlowField" % i]) < 1e-5      # step length, coefficient update, P coeffs
        assert rel(outF[i], g["%s.it%02d.tFlowField" % (name, i)]) < 1e-3  # the reference itself (north_star gate)
        if vf_mode == 2:
            assert rel(outC[i], g["%s.it%02d.coeffs" % (name, i)]) < 1e-5
