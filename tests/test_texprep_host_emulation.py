"""CPU tier: meshopticalflow_b200/csrc/texprep_kernels.cu (edge-length subdivision, the texel ->
(triangle, point) map with its padding and exp-map pull-back, wedge-averaged vertex colours) — the real CUDA source compiled
for the HOST by tests/host_emulation — against the C checker and the reference's golden fixture. The reference's loops are
serial and order dependent; the kernels are not, so everything integer (numbering, connectivity, triangle per texel) is
compared bit for bit, and so are the midpoint coordinates."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel
from meshopticalflow_b200 import synthetic
from oracle import mof_oracle as O

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
_F, _I, _D, _U8 = ctypes.c_float, ctypes.c_int, ctypes.c_double, ctypes.c_ubyte


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("texprep_emul") / "libtexprep_emul.so")
    extra = os.environ.get("MOF_EMUL_CXXFLAGS", "-O2").split()  # e.g. "-O1 -g -fsanitize=address", see test_library_host_emulation.py
    subprocess.check_call(["g++"] + extra + ["-std=c++17", "-fPIC", "-shared", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-I.", "-w", "-o", out, "texprep_emul.cpp",
                           "emul_runtime.cpp"], cwd=EMU_DIR)
    return ctypes.CDLL(out)


def _subdivide(emul, v, t, uv, e_len):
    v, t, uv = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(t, np.int32), np.ascontiguousarray(uv, np.float64)
    sizes = (ctypes.c_int * 2)()
    assert emul.emul_subdivide(v.shape[0], t.shape[0], _p(v, _F), _p(t, _I), _p(uv, _D), _D(e_len), sizes) == 0
    V, T = sizes[0], sizes[1]
    vo, to, uo = np.empty((V, 3), np.float32), np.empty((T, 3), np.int32), np.empty((T, 6), np.float64)
    emul.emul_get_subdivision(_p(vo, _F), _p(to, _I), _p(uo, _D))
    return vo, to, uo


def _prepare(emul, v, t, uv, ta, tb, pad=2, bilinear=True):
    v, t, uv = np.ascontiguousarray(v, np.float64), np.ascontiguousarray(t, np.int32), np.ascontiguousarray(uv, np.float64)
    ta, tb = np.ascontiguousarray(ta, np.uint8), np.ascontiguousarray(tb, np.uint8)
    H, W = ta.shape[:2]
    srcT, srcP, col = np.empty(W * H, np.int32), np.empty((W * H, 2)), np.empty((v.shape[0], 6))
    misses, launches = ctypes.c_int(0), ctypes.c_longlong(0)
    rc = emul.emul_texture_prepare(v.shape[0], t.shape[0], _p(v, _D), _p(t, _I), _p(uv, _D), W, H, pad, _p(ta, _U8), _p(tb, _U8), 1 if bilinear else 0, _p(srcT, _I),
                                   _p(srcP, _D), _p(col, _D), ctypes.byref(misses), ctypes.byref(launches))
    assert rc == 0
    return srcT, srcP, col, misses.value, launches.value


def _edge_length(v, param):
    lo, hi = v.astype(np.float64).min(0), v.astype(np.float64).max(0)
    return float(np.float32(np.float32(param) * float(np.sqrt(((hi - lo) ** 2).sum()))))  # OpticalFlow.cpp:713


def test_subdivision_matches_the_reference_golden(emul, golden_torus):
    g = golden_torus
    v, t, uv = g["input_vertices_f32"], g["input_triangles"], g["input_uv"].astype(np.float64)
    vo, to, uo = _subdivide(emul, v, t, uv, _edge_length(v, 0.08))
    assert np.array_equal(vo.astype(np.float64), g["vertices"])  # same vertices, same order, same bits
    assert np.array_equal(to, g["triangles"])
    assert rel(uo.reshape(-1, 2), g["triangleTextures"]) < 1e-15
    rv, rt, ruv = O.subdivide(v, t, uv, _edge_length(v, 0.08))
    assert np.array_equal(vo, rv) and np.array_equal(to, rt) and np.array_equal(uo, ruv)


@pytest.mark.parametrize("param", [0.2, 0.15, 0.11, 0.05])
def test_subdivision_all_split_cases_and_several_sweeps(emul, param):
    """An anisotropic torus: sweeps with one, two and three long sides per triangle, several sweeps deep."""
    v, t, uv = synthetic.uv_torus(9, 5)
    e_len = _edge_length(v, param)
    rv, rt, ruv = O.subdivide(v, t, uv.astype(np.float64), e_len)
    vo, to, uo = _subdivide(emul, v, t, uv.astype(np.float64), e_len)
    assert rv.shape[0] > v.shape[0]
    assert np.array_equal(vo, rv) and np.array_equal(to, rt) and np.array_equal(uo, ruv)


def test_subdivision_without_long_edges_is_the_identity(emul):
    v, t, uv = synthetic.uv_torus(6, 4)
    vo, to, uo = _subdivide(emul, v, t, uv.astype(np.float64), 100.0)
    assert np.array_equal(vo, v) and np.array_equal(to, t) and np.array_equal(uo, uv.astype(np.float64))
    vo, to, uo = _subdivide(emul, v, t, uv.astype(np.float64), 0.0)  # eLength <= 0: OpticalFlow.cpp:714 skips the call
    assert np.array_equal(vo, v) and np.array_equal(to, t)


def test_texel_map_and_vertex_colours_match_the_reference_golden(emul, golden_torus):
    g = golden_torus
    ta, tb = g["input_tex_a"], g["input_tex_b"]
    v, t, uv = g["vertices"], g["triangles"], g["triangleTextures"].reshape(-1, 6)
    srcT, srcP, col, misses, launches = _prepare(emul, v, t, uv, ta, tb)
    assert misses == 0 and launches == 3 + 2 * 2 + 1 + 1
    assert np.array_equal(srcT, g["textureSource_tIdx"])
    covered = srcT >= 0
    assert covered.sum() > 1000 and np.abs(srcP[covered] - g["textureSource_p"][covered]).max() < 1e-10
    # and the C checker, uncovered texels included
    st = O.init(v, t, np.zeros((v.shape[0], 3)), np.zeros((v.shape[0], 3)), O.Params(dogWeight=0.0))
    rT, rP = O.texture_source(uv, 48, 48, 2, st.opp, st.lin, st.cst, st.g)
    assert np.array_equal(srcT, rT) and np.abs(srcP - rP)[covered].max() < 1e-10
    for tex, mine in ((ta, col[:, :3]), (tb, col[:, 3:])):
        assert np.abs(mine - O.sample_texture_to_vertices(t, uv, v.shape[0], tex)).max() < 1e-11


@pytest.mark.parametrize("case", ["seams_pad3_bilinear", "large_triangles_nearest", "no_padding", "vertices_on_texels"])
def test_texel_map_on_other_charts(emul, case):
    """Charts whose seam faces span the texture backwards (many triangles per texel: the first-writer rule decides), triangles
    much larger than a texel (lanes share scan lines), nearest-neighbour sampling, padding radius 0 and 3."""
    if case == "seams_pad3_bilinear":
        (v, t, uv), W, H, pad, bil = synthetic.uv_torus(24, 12), 40, 36, 3, True
    elif case == "large_triangles_nearest":
        (v, t, uv), W, H, pad, bil = synthetic.uv_torus(6, 4), 64, 50, 2, False
    elif case == "vertices_on_texels":  # every vertex sits exactly on a texel centre: the rule of MeshFlow.inl:334 (point == corner 2) fires
        (v, t, uv), W, H, pad, bil = synthetic.uv_torus(24, 12), 49, 25, 2, True
    else:
        (v, t, uv), W, H, pad, bil = synthetic.uv_torus(16, 10), 33, 47, 0, True
    ta, tb = synthetic.smooth_texture_pair(W, H, 3)
    uv = uv.astype(np.float64)
    srcT, srcP, col, misses, _ = _prepare(emul, v, t, uv, ta, tb, pad, bil)
    vd = v.astype(np.float64)
    st = O.init(vd, t, np.zeros((v.shape[0], 3)), np.zeros((v.shape[0], 3)), O.Params(dogWeight=0.0))
    rT, rP = O.texture_source(uv, W, H, pad, st.opp, st.lin, st.cst, st.g)
    assert misses == 0 and np.array_equal(srcT, rT)
    covered = srcT >= 0
    assert covered.any() and np.abs(srcP - rP)[covered].max() < 1e-9
    assert np.all(srcP[covered] >= -1e-12) and np.all(srcP[covered].sum(1) <= 1 + 1e-12)  # every point pulled back inside its triangle
    for tex, mine in ((ta, col[:, :3]), (tb, col[:, 3:])):
        assert np.abs(mine - O.sample_texture_to_vertices(t, uv, v.shape[0], tex, bil)).max() < 1e-11
