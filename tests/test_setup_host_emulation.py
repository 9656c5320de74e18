"""CPU tier: meshopticalflow_b200/csrc/setup_kernels.cu (metric, half-edge adjacency, edge
transforms, scalar mass / stiffness CSR, Whitney numbering, prolongation, smoothness operator in the sliced layout) — the
real CUDA source compiled for the HOST by tests/host_emulation (fibers; counted barriers, warp shuffles and atomics
emulated) — against the numpy checker and the reference's golden fixture: adjacency, numbering and sparsity patterns
bit-exact, values to round-off, and the reference's error paths."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, csr_from_golden, rel
from meshopticalflow_b200 import synthetic
from oracle import mof_oracle as O

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("setup_emul") / "libsetup_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-I.", "-w", "-o", out, "setup_emul.cpp",
                           "emul_runtime.cpp"], cwd=EMU_DIR)
    return ctypes.CDLL(out)


def _build(emul, v, t):
    v = np.ascontiguousarray(v, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.int32)
    sizes = (ctypes.c_longlong * 5)()
    msg = ctypes.create_string_buffer(256)
    rc = emul.emul_mesh_build(v.shape[0], t.shape[0], v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), t.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), sizes, msg, 256)
    return rc, msg.value.decode(), [int(x) for x in sizes]


def _get(emul, name, dtype, count):
    out = np.zeros(count, dtype=dtype)
    assert emul.emul_mesh_get(name.encode(), out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def _unsliced(rowptr, base, col, val):
    """The product's sliced layout back to CSR (entry j of row r at base[r / 32] + 32 j + r % 32)."""
    n = rowptr.size - 1
    cols, vals = [], []
    for r in range(n):
        k = base[r // 32] + 32 * np.arange(rowptr[r + 1] - rowptr[r]) + r % 32
        cols.append(col[k]), vals.append(val[k])
    return sp.csr_matrix((np.concatenate(vals), np.concatenate(cols), rowptr), shape=(n, n))


@pytest.mark.parametrize("mesh", ["golden_sphere", "jittered_sphere"])
def test_operator_assembly_from_the_cuda_source(emul, golden_sphere, mesh):
    if mesh == "golden_sphere":
        g = golden_sphere
        v, t = g["input_vertices_f32"].astype(np.float64), g["triangles"]
    else:  # irregular triangle areas and valences; a second scan tile (more than 1024 half-edges per tile boundary)
        v, t = synthetic.octahedron_sphere(4)
        rng = np.random.default_rng(5)
        v = v + 0.02 * rng.standard_normal(v.shape)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
    V, T = v.shape[0], t.shape[0]
    rc, msg, (E, nnzS, nnzW, slices, padded) = _build(emul, v, t)
    assert rc == 0, msg
    st = O.init(v, t, np.zeros((V, 3)), np.zeros((V, 3)), O.Params(dogWeight=0.0))
    w = st.whitney
    # a1-a3
    assert rel(_get(emul, "g", np.float64, 3 * T).reshape(T, 3), st.g) < 1e-13 and rel(_get(emul, "area", np.float64, T), st.area) < 1e-13
    assert np.array_equal(_get(emul, "opp", np.int32, 3 * T), st.opp)
    assert rel(_get(emul, "xlin", np.float64, 12 * T).reshape(-1, 4), st.lin) < 1e-12 and rel(_get(emul, "xcst", np.float64, 6 * T).reshape(-1, 2), st.cst) < 1e-12
    # a4: one pattern, columns ascending, for mass and stiffness
    assert nnzS == st.M.nnz
    rowptr, col = _get(emul, "sRowptr", np.int32, V + 1), _get(emul, "sCol", np.int32, nnzS)
    assert np.array_equal(rowptr, st.M.indptr) and np.array_equal(col, st.M.indices)
    assert rel(_get(emul, "sMass", np.float64, nnzS), st.M.data) < 1e-12 and rel(_get(emul, "sStiff", np.float64, nnzS), st.S.data) < 1e-12
    he = _get(emul, "sHe", np.int32, nnzS)
    from test_vf_host_emulation import _half_edge_rows
    assert np.array_equal(he, _half_edge_rows(st.S, t))
    # a6, a7: numbering bit-exact, prolongation to round-off
    assert E == w.expanded.size
    assert np.array_equal(_get(emul, "reduced", np.int32, 3 * T), w.reduced) and np.array_equal(_get(emul, "expanded", np.int32, E), w.expanded)
    assert np.array_equal(_get(emul, "positive", np.int32, 3 * T), w.positive.astype(np.int32))
    P = _get(emul, "P", np.float64, 6 * T).reshape(T, 3, 2)
    Pd = w.P.tocsr()
    for k in range(3):
        for r in range(2):
            assert rel(P[:, k, r], np.asarray(Pd[2 * np.arange(T) + r, w.reduced[3 * np.arange(T) + k]]).ravel()) < 1e-12
    # a8: the smoothness operator, pattern bit-exact
    assert nnzW == w.S.nnz
    wrow, base = _get(emul, "wRowptr", np.int32, E + 1), _get(emul, "wSliceBase", np.int32, slices + 1)
    assert base[-1] == padded and slices == (E + 31) // 32
    S = _unsliced(wrow, base, _get(emul, "wCol", np.int32, padded), _get(emul, "wS", np.float64, padded))
    assert np.array_equal(S.indptr, w.S.indptr) and np.array_equal(S.indices, w.S.indices)
    assert rel(S.data, w.S.data) < 1e-11
    assert rel(_get(emul, "m0", np.float64, V), np.bincount(t.reshape(-1), np.repeat(st.area / 3.0, 3), V)) < 1e-13
    if mesh == "golden_sphere":  # and the reference itself
        assert np.array_equal(_get(emul, "opp", np.int32, 3 * T), g["oppositeEdge"]) and np.array_equal(_get(emul, "reduced", np.int32, 3 * T), g["reducedEdgeIndex"])
        ref = csr_from_golden(g, "smoothOperator", (E, E))
        assert np.array_equal(S.indptr, ref.indptr) and np.array_equal(S.indices, ref.indices) and rel(S.data, ref.data) < 1e-11


def test_bad_meshes_are_rejected_like_the_reference(emul):
    rc, msg, _ = _build(emul, np.eye(4, 3), np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32))
    assert rc == -3 and "Edge is occupied" in msg       # FEM.inl:599
    rc, msg, _ = _build(emul, np.eye(3), np.array([[0, 1, 2]], dtype=np.int32))
    assert rc == -3 and "Boundary edge" in msg          # FEM.inl:554
    rc, msg, _ = _build(emul, np.eye(3), np.array([[0, 1, 7]], dtype=np.int32))
    assert rc == -1
