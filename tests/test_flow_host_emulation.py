"""CPU tier: meshopticalflow_b200/csrc/flow_kernels.cu (DoG, smoothing right-hand sides, triangle walks, vertex gather,
data term, Whitney system assembly on the sliced layout, step length and coefficient update, final advection) together
with vector_fields.cu — the real CUDA sources, compiled for the HOST by tests/host_emulation (CUDA runtime calls and the
launch macro replaced by stand-ins, thread blocks on fibers) — run a whole alignment and are checked against the numpy
checker and the reference's golden fixtures. The linear solvers (pcg_kernels.cu, multigrid.cu) are NOT part of this:
host conjugate-gradient loops stand in for them here; tests/test_library_host_emulation.py runs them too (the whole
library, through the C ABI), and the GPU tier runs everything on a B200."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel
from oracle import mof_oracle as O
from test_vf_host_emulation import _half_edge_rows

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
_D, _I = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("flow_emul") / "libflow_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-DEMUL_WITH_FLOW", "-I.", "-w", "-o", out,
                           "flow_emul.cpp", "vf_emul.cpp", "emul_runtime.cpp", "-lpthread"], cwd=EMU_DIR)
    return ctypes.CDLL(out)


def _sliced(S):
    """CSR -> the product's sliced layout (mof_internal.cuh, sell_pos): 32-row slices padded to their longest row, entry j
    of row r at base[r / 32] + 32 j + r % 32, padding = value 0 on the row's own column."""
    n = S.shape[0]
    lens = np.diff(S.indptr)
    slices = (n + 31) // 32
    base = np.zeros(slices + 1, dtype=np.int32)
    for s in range(slices):
        base[s + 1] = base[s] + 32 * int(lens[32 * s:32 * s + 32].max())
    col = np.zeros(base[-1], dtype=np.int32)
    val = np.zeros(base[-1])
    for r in range(n):
        width = (base[r // 32 + 1] - base[r // 32]) // 32
        pos = base[r // 32] + 32 * np.arange(width) + r % 32
        col[pos] = r
        k = np.arange(S.indptr[r], S.indptr[r + 1])
        col[pos[:k.size]], val[pos[:k.size]] = S.indices[k], S.data[k]
    return base, col, val, slices


def _run(emul, v, t, a, b, params: O.Params):
    st = O.init(v, t, a, b, params)
    V, T = v.shape[0], t.shape[0]
    w = st.whitney
    E = w.expanded.size
    he = _half_edge_rows(st.S, t)
    m0 = np.zeros(V)
    np.add.at(m0, t.reshape(-1), np.repeat(st.area / 3.0, 3))
    # P as the product stores it: [t][k][r] = entry of row 2t+r on the k-th edge of t (only meaningful for the Whitney basis)
    P = np.zeros((T, 3, 2))
    if params.vfMode == 0:
        Pd = w.P.tocsr()
        for k in range(3):
            for r in range(2):
                P[:, k, r] = np.asarray(Pd[2 * np.arange(T) + r, w.reduced[3 * np.arange(T) + k]]).ravel()
        Sw = w.S
    else:  # the Whitney operators are not used by the other bases; any valid sliced matrix will do
        import scipy.sparse as sp
        Sw = sp.identity(E, format="csr")
    base, wcol, wval, slices = _sliced(Sw)
    raw6 = np.ascontiguousarray(np.hstack([a, b]))
    it = params.iterations
    out = dict(sig6=np.zeros((V, 6)), lo6=np.zeros((V, 6)), field=np.zeros((it, T, 2)), data=np.zeros((it, T, 3)), adv=np.zeros((V, 6)), rhs=np.zeros((it, E)),
               x=np.zeros((it, E)))
    launches = ctypes.c_longlong()
    A = lambda x, ty=np.float64: np.ascontiguousarray(x, dtype=ty)
    keep = [A(t, np.int32), A(st.opp, np.int32), A(st.g), A(st.area), A(st.lin), A(st.cst), A(st.S.indptr, np.int32), A(st.S.indices, np.int32), A(he, np.int32),
            A(st.M.data), A(st.S.data), A(m0), A(w.reduced, np.int32), A(w.expanded, np.int32), A(P), A(Sw.indptr, np.int32), A(base, np.int32), A(wcol, np.int32),
            A(wval), raw6]
    ptr = lambda x: x.ctypes.data_as(_I if x.dtype == np.int32 else _D)
    d = ctypes.c_double
    rc = emul.emul_flow_run(V, T, E, *[ptr(x) for x in keep[:16]], ptr(keep[16]), slices, ctypes.c_longlong(int(base[-1])), ptr(keep[17]), ptr(keep[18]), ptr(keep[19]),
                            it, d(params.sSmooth), d(params.sMultiply), d(params.vfSmooth), d(params.dogWeight), d(params.dogSmooth), params.vfMode, params.cMode,
                            d(1e-10), d(1e-12), ptr(out["sig6"]), ptr(out["lo6"]), ptr(out["field"]), ptr(out["data"]), ptr(out["adv"]), ptr(out["rhs"]), ptr(out["x"]), ctypes.byref(launches))
    assert rc == 0
    assert launches.value > 10 * it
    return st, out


def _inputs(g):
    return g["input_vertices_f32"].astype(np.float64), g["triangles"], g["input_a"].astype(np.float64), g["input_b"].astype(np.float64)


def test_whitney_alignment_from_the_cuda_source(emul, golden_sphere):
    g = golden_sphere
    v, t, a, b = _inputs(g)
    params = O.Params(iterations=4)
    st, out = _run(emul, v, t, a, b, params)
    assert rel(out["sig6"][:, :3], g["signals0"]) < 1e-9 and rel(out["sig6"][:, 3:], g["signals1"]) < 1e-9        # DoG
    O.iterate(st, params, taps=True)
    for i in range(4):
        assert rel(out["data"][i], st.taps["it%02d.dataTerm" % i]) < 1e-7
        assert rel(out["field"][i], st.taps["it%02d.tFlowField" % i]) < 1e-7
        assert rel(out["field"][i], g["it%02d.tFlowField" % i]) < 1e-7          # the reference itself
    ca, cb = O.advect_vertices(st, a, b)
    assert np.abs(out["adv"][:, :3] - ca).max() < 1e-5 and np.abs(out["adv"][:, 3:] - cb).max() < 1e-5


def test_six_channel_blend_from_the_cuda_source(emul, golden_modes):
    g = golden_modes
    v, t, a, b = _inputs(g)
    params = O.Params(iterations=4, dogWeight=0.5)
    st, out = _run(emul, v, t, a, b, params)
    for s in range(2):  # the reference's Point<Real,6>: channels 0-2 = (1-w) raw, 3-5 = w DoG
        assert rel(out["lo6"][:, 3 * s:3 * s + 3], st.signals[s][:, :3]) < 1e-12 and rel(out["sig6"][:, 3 * s:3 * s + 3], st.signals[s][:, 3:]) < 1e-9
    O.iterate(st, params, taps=True)
    for i in range(4):
        assert rel(out["field"][i], st.taps["it%02d.tFlowField" % i]) < 1e-7
    ca, cb = O.advect_vertices(st, a, b)
    blended = (out["adv"][:, :3] + out["adv"][:, 3:]) / 2.0
    assert np.abs(O.to_uchar_ply(blended).astype(int) - g["blend.output_rgb"].astype(int)).max() <= 1


def test_connection_basis_through_the_whole_loop(emul, golden_modes):
    g = golden_modes
    v, t, a, b = _inputs(g)
    params = O.Params(iterations=4, vfMode=2, cMode=1)
    st, out = _run(emul, v, t, a, b, params)
    for i in range(4):
        assert rel(out["field"][i], g["connection1.it%02d.tFlowField" % i]) < 1e-6
    assert np.abs(out["adv"][:, :3] - g["connection1.advected0"]).max() < 1e-4 and np.abs(out["adv"][:, 3:] - g["connection1.advected1"]).max() < 1e-4


def test_texel_advection_from_the_cuda_source(emul, golden_torus):
    """k_advect_texels on the reference's final flow field and texel map (uv torus, 48 x 48 texels): the advected textures."""
    from conftest import colour_outliers
    g = golden_torus
    v, t, uv = g["vertices"], g["triangles"], g["triangleTextures"].reshape(-1, 6)
    gm = np.ascontiguousarray(O.make_unit_area(O.metric_from_embedding(v, t)))
    opp = O.opposite_half_edges(t)
    lin, cst = O.edge_xforms(gm, opp)
    W = H = 48
    out = np.zeros((2, W * H, 3))
    A = lambda x, ty=np.float64: np.ascontiguousarray(x, dtype=ty)
    keep = [A(opp, np.int32), A(gm), A(lin), A(cst), A(g["it09.tFlowField"]), A(g["textureSource_tIdx"], np.int32), A(g["textureSource_p"]), A(uv),
            A(g["input_tex_a"], np.uint8), A(g["input_tex_b"], np.uint8)]
    P = lambda x: x.ctypes.data_as(_I if x.dtype == np.int32 else ctypes.POINTER(ctypes.c_ubyte) if x.dtype == np.uint8 else _D)
    rc = emul.emul_advect_texels(t.shape[0], P(keep[0]), P(keep[1]), P(keep[2]), P(keep[3]), P(keep[4]), W, H, P(keep[5]), P(keep[6]), P(keep[7]), P(keep[8]), P(keep[9]),
                                 ctypes.c_double(0.5), 1, out.ctypes.data_as(_D))
    assert rc == 0
    for s in range(2):  # colours are compared, not walk paths: a point exactly on an edge can go either way (conftest.colour_outliers)
        assert colour_outliers(out[s], g["advected%d" % s], 1e-6) < 2e-3
