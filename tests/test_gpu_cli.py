"""GPU tier: the drop-in command line (meshopticalflow_b200/OpticalFlow) end to end on files, against the
oracle, the golden outputs of the reference, and — where the prebuilt reference binary travelled with the
snapshot — the reference itself run on the box's CPU on the same files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CLI_BIN, REF_BIN, colour_outliers
from meshopticalflow_b200 import synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.gpu


def _rgb(path):
    out = synthetic.read_ply(path)
    return np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(int)


@pytest.mark.parametrize("binary", [True, False])
def test_vertex_configuration(tmp_path, binary):
    v, t = synthetic.octahedron_sphere(4)
    a, b = synthetic.smooth_rgb_pair(v, 7)
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), v, a, t, binary)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), v, b, t, binary)
    r = subprocess.run([CLI_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--verbose"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Vertices / Triangles: 1026 / 2048" in r.stdout and r.stdout.count("Got flow") == 10
    text = open(tmp_path / "r.ply").read().split("\n")
    assert text[:3] == ["ply", "format ascii 1.0", "element vertex 1026"] and text[10] == "property list uchar int vertex_indices"
    mine = _rgb(str(tmp_path / "r.ply"))
    # the files store float positions: binary keeps float32, ascii prints %g (6 digits)
    src = synthetic.read_ply(str(tmp_path / "A.ply"))
    vf = np.stack([src["vertex"][k] for k in "xyz"], 1).astype(np.float32).astype(np.float64)
    _, blended = O.align_vertices(vf, t, a.astype(np.float64), b.astype(np.float64))
    assert np.abs(mine - O.to_uchar_ply(blended).astype(int)).max() <= 1
    if os.path.exists(REF_BIN):
        subprocess.check_call([REF_BIN, "--in", "A.ply", "B.ply", "--out", "ref.ply"], cwd=tmp_path, stdout=subprocess.DEVNULL)
        assert np.abs(mine - _rgb(str(tmp_path / "ref.ply"))).max() <= 1
        ours, theirs = open(tmp_path / "r.ply").read().split("\n"), open(tmp_path / "ref.ply").read().split("\n")
        assert ours[:12] == theirs[:12]                          # identical header
        assert ours[12 + 1026:] == theirs[12 + 1026:]            # identical face lines
        assert [ln.split()[:3] for ln in ours[12:12 + 1026]] == [ln.split()[:3] for ln in theirs[12:12 + 1026]]  # identical positions


def test_texture_configuration(tmp_path, golden_torus):
    from PIL import Image
    g = golden_torus
    synthetic.write_ply_textured(str(tmp_path / "m.ply"), g["input_vertices_f32"], g["input_triangles"], g["input_uv"])
    open(tmp_path / "A.png", "wb").write(g["png_a"].tobytes())   # RGB
    open(tmp_path / "B.png", "wb").write(g["png_b"].tobytes())   # RGBA: alpha is dropped (PNG.inl:65-73)
    r = subprocess.run([CLI_BIN, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "r.png", "--eLength", "0.08"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Num vertices %d" % g["vertices"].shape[0] in r.stdout  # same subdivision as the reference (OpticalFlow.cpp:716)
    pixels = np.asarray(Image.open(tmp_path / "r.png"))
    assert pixels.shape == (48, 48, 3)
    assert colour_outliers(pixels, g["output_pixels"], 1.0) < 2e-3
    if os.path.exists(REF_BIN):
        subprocess.check_call([REF_BIN, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "ref.png", "--eLength", "0.08"], cwd=tmp_path,
                              stdout=subprocess.DEVNULL)
        assert colour_outliers(pixels, np.asarray(Image.open(tmp_path / "ref.png")), 1.0) < 2e-3


def test_mismatched_inputs_are_rejected(tmp_path):
    v, t = synthetic.octahedron_sphere(2)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    v3, t3 = synthetic.octahedron_sphere(3)
    a3, _ = synthetic.smooth_rgb_pair(v3, 0)
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), v, a, t)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), v3, a3, t3)
    r = subprocess.run([CLI_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "Vertex counts differ" in r.stderr  # OpticalFlow.cpp:761
    t_bad = t.copy()
    t_bad[0] = t_bad[0][[1, 2, 0]]
    synthetic.write_ply_colored(str(tmp_path / "C.ply"), v, b, t_bad)
    r = subprocess.run([CLI_BIN, "--in", "A.ply", "C.ply", "--out", "r.ply"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "Triangle indices don't match" in r.stderr  # OpticalFlow.cpp:769
