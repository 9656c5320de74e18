"""GPU tier: the texture configuration's one-time preparation on the device (csrc/texprep_kernels.cu: mof_subdivide,
mof_build_texture_map, mof_sample_textures_to_vertices), through the C ABI, against the C
checker, the reference's golden fixture and the host preparation of the command line. Integer outputs (numbering,
connectivity, triangle per texel) and the single-precision midpoints are compared bit for bit.

tests/test_texprep_host_emulation.py runs the same CUDA source on the CPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CLI_BIN, colour_outliers
from meshopticalflow_b200 import api, synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def aligner():
    al = api.Aligner(0)
    yield al
    al.close()


def _edge_length(v, param):
    lo, hi = v.astype(np.float64).min(0), v.astype(np.float64).max(0)
    return float(np.float32(np.float32(param) * float(np.sqrt(((hi - lo) ** 2).sum()))))  # OpticalFlow.cpp:713


def test_subdivision_matches_the_reference_golden(aligner, golden_torus):
    g = golden_torus
    v, t, uv = g["input_vertices_f32"], g["input_triangles"], g["input_uv"].astype(np.float64)
    vo, to, uo = aligner.subdivide(v, t, uv, _edge_length(v, 0.08))
    assert np.array_equal(vo.astype(np.float64), g["vertices"]) and np.array_equal(to, g["triangles"])
    assert np.abs(uo.reshape(-1, 2) - g["triangleTextures"]).max() < 1e-15
    rv, rt, ruv = O.subdivide(v, t, uv, _edge_length(v, 0.08))
    assert np.array_equal(vo, rv) and np.array_equal(to, rt) and np.array_equal(uo, ruv)


@pytest.mark.parametrize("nu,nv,param", [(9, 5, 0.15), (9, 5, 0.05), (40, 24, 0.004)])
def test_subdivision_matches_the_checker(aligner, nu, nv, param):
    """One, two and three long sides per triangle, several sweeps; the last case ends at 247 200 vertices / 494 400 triangles (hundreds
    of scan tiles, a hash table of 4M slots)."""
    v, t, uv = synthetic.uv_torus(nu, nv)
    e_len = _edge_length(v, param)
    rv, rt, ruv = O.subdivide(v, t, uv.astype(np.float64), e_len)
    vo, to, uo = aligner.subdivide(v, t, uv, e_len)
    assert rv.shape[0] > v.shape[0]
    assert np.array_equal(vo, rv) and np.array_equal(to, rt) and np.array_equal(uo, ruv)
    again = aligner.subdivide(v, t, uv, e_len)  # deterministic run to run (atomics only decide minima)
    assert all(np.array_equal(x, y) for x, y in zip((vo, to, uo), again))
    same = aligner.subdivide(v, t, uv, 0.0)   # OpticalFlow.cpp:714: no subdivision without a positive edge length
    assert np.array_equal(same[0], v) and np.array_equal(same[1], t)


def test_texel_map_and_vertex_colours_match_the_reference_golden(aligner, golden_torus):
    g = golden_torus
    ta, tb = g["input_tex_a"], g["input_tex_b"]
    v, t, uv = g["vertices"], g["triangles"], g["triangleTextures"].reshape(-1, 6)
    al = aligner
    al.set_mesh(v, t)
    srcT, srcP = al.build_texture_map(48, 48, 2, uv, ta, tb)
    assert np.array_equal(srcT, g["textureSource_tIdx"])
    covered = srcT >= 0
    assert np.abs(srcP[covered] - g["textureSource_p"][covered]).max() < 1e-10
    ca, cb = al.sample_textures_to_vertices()
    for tex, mine in ((ta, ca), (tb, cb)):
        assert np.abs(mine - O.sample_texture_to_vertices(t, uv, v.shape[0], tex)).max() < 1e-11
    # the alignment on top of the device-built map and colours ends at the reference's output
    al.set_signals(ca, cb)
    al.iterate(10)
    oa, ob = al.advect_texels(0.5)
    assert colour_outliers(oa, g["advected0"], 1.0) < 2e-3 and colour_outliers(ob, g["advected1"], 1.0) < 2e-3
    # and the uploaded-map entry gives the same texels as the device-built map
    al.set_texture_map(48, 48, g["textureSource_tIdx"], g["textureSource_p"], uv, ta, tb)
    ua, ub = al.advect_texels(0.5)
    assert colour_outliers(oa, ua, 1e-6) < 2e-3 and colour_outliers(ob, ub, 1e-6) < 2e-3


@pytest.mark.parametrize("case", ["seams_pad3_bilinear", "large_triangles_nearest", "no_padding", "vertices_on_texels", "example_sized"])
def test_texel_map_matches_the_checker(aligner, case):
    W, H, pad, bil = {"seams_pad3_bilinear": (40, 36, 3, True), "large_triangles_nearest": (64, 50, 2, False), "no_padding": (33, 47, 0, True),
                      "vertices_on_texels": (49, 25, 2, True), "example_sized": (388, 388, 2, True)}[case]
    nu, nv = {"seams_pad3_bilinear": (24, 12), "large_triangles_nearest": (6, 4), "no_padding": (16, 10), "vertices_on_texels": (24, 12), "example_sized": (160, 96)}[case]
    v, t, uv = synthetic.uv_torus(nu, nv)
    uv = uv.astype(np.float64)
    ta, tb = synthetic.smooth_texture_pair(W, H, 3)
    vd = v.astype(np.float64)
    al = aligner
    al.set_mesh(vd, t)
    srcT, srcP = al.build_texture_map(W, H, pad, uv, ta, tb)
    st = O.init(vd, t, np.zeros((v.shape[0], 3)), np.zeros((v.shape[0], 3)), O.Params(dogWeight=0.0))
    rT, rP = O.texture_source(uv, W, H, pad, st.opp, st.lin, st.cst, st.g)
    assert np.array_equal(srcT, rT)
    covered = srcT >= 0
    assert covered.any() and np.abs(srcP - rP)[covered].max() < 1e-9
    ca, cb = al.sample_textures_to_vertices(bil)
    for tex, mine in ((ta, ca), (tb, cb)):
        assert np.abs(mine - O.sample_texture_to_vertices(t, uv, v.shape[0], tex, bil)).max() < 1e-11
    again = al.build_texture_map(W, H, pad, uv, ta, tb)
    assert np.array_equal(again[0], srcT) and np.array_equal(again[1], srcP)


def test_error_paths(aligner):
    al = aligner
    v, t, uv = synthetic.uv_torus(6, 4)
    ta, tb = synthetic.smooth_texture_pair(16, 16, 0)
    with pytest.raises(api.MofError) as e:  # no mesh yet
        al.build_texture_map(16, 16, 2, uv, ta, tb)
    assert e.value.code == api.MOF_E_INVALID
    al.set_mesh(v.astype(np.float64), t)
    with pytest.raises(api.MofError):       # no textures yet
        al.sample_textures_to_vertices()
    with pytest.raises(api.MofError):       # a 1 x n texture has no (W-1) to scale by
        al.build_texture_map(1, 16, 2, uv, ta, tb)
    bad = t.copy()
    bad[0, 0] = v.shape[0]
    with pytest.raises(api.MofError) as e:
        al.subdivide(v, bad, uv, 0.1)
    assert e.value.code == api.MOF_E_INVALID


def test_command_line_with_device_preparation(tmp_path, golden_torus):
    """The command line prepares the texture configuration on the GPU. The serial host restatement of the preparation
    (csrc/host/texture_prep.cpp) is no longer part of the product: it is compiled into a test binary here (-DMOF_WITH_HOST_TEXPREP, where
    MOF_GPU_TEXPREP=0 selects it) as the cross-check — the same picture from both, and the reference's."""
    from PIL import Image
    g = golden_torus
    synthetic.write_ply_textured(str(tmp_path / "m.ply"), g["input_vertices_f32"], g["input_triangles"], g["input_uv"])
    open(tmp_path / "A.png", "wb").write(g["png_a"].tobytes())
    open(tmp_path / "B.png", "wb").write(g["png_b"].tobytes())
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host, libdir = os.path.join(root, "meshopticalflow_b200", "csrc", "host"), os.path.join(root, "meshopticalflow_b200")
    checker = str(tmp_path / "OpticalFlow_hostprep")
    sources = [os.path.join(host, f) for f in ("optical_flow_main.cpp", "ply_io.cpp", "png_codec.cpp", "texture_prep.cpp", "cmdline.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMOF_WITH_HOST_TEXPREP", "-I" + os.path.join(root, "include"), "-o", checker] + sources +
                          ["-L" + libdir, "-lmof_b200", "-lz", "-Wl,-rpath," + libdir])
    outs = []
    for exe, mode in ((checker, "0"), (CLI_BIN, "1")):
        r = subprocess.run([exe, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "r%s.png" % mode, "--eLength", "0.08"], cwd=tmp_path, capture_output=True,
                           text=True, timeout=300, env=dict(os.environ, MOF_GPU_TEXPREP=mode))
        assert r.returncode == 0, r.stderr
        assert "Num vertices %d" % g["vertices"].shape[0] in r.stdout
        outs.append(np.asarray(Image.open(tmp_path / ("r%s.png" % mode))).astype(int))
    assert outs[0].shape == (48, 48, 3) and np.abs(outs[0] - outs[1]).max() <= 1
    assert colour_outliers(outs[1], g["output_pixels"], 1.0) < 2e-3
