"""CPU tier: meshopticalflow_b200/SampleTextureToVertices — the drop-in for the reference's sibling tool of the same
name (it makes the per-vertex inputs A.ply / B.ply from the texture configuration's files; pure host code) — against
the reference tool's own output files: byte for byte, from the golden fixture everywhere and live where the reference
binary was built (oracle/_ref/SampleTextureToVertices_ref, compiled unmodified by oracle/ref/build_ref.sh)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

TOOL = os.path.join(ROOT, "meshopticalflow_b200", "SampleTextureToVertices")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "SampleTextureToVertices_ref")
RUNS = {"plain": ["--in", "m.ply"], "subdivided": ["--in", "m.ply", "--eLength", "0.08"], "binary": ["--in", "mb.ply", "--eLength", "0.1"]}

pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="host tools not built (make -C meshopticalflow_b200/csrc)")


@pytest.fixture()
def workdir(tmp_path):
    g = np.load(os.path.join(GOLDEN, "sample_texture_tool.npz"))
    for f in ("m.ply", "mb.ply", "A.png"):
        open(tmp_path / f, "wb").write(g["in_" + f].tobytes())
    return tmp_path, g


@pytest.mark.parametrize("name", sorted(RUNS))
def test_output_files_are_the_reference_tools(workdir, name):
    d, g = workdir
    r = subprocess.run([TOOL, "--texture", "A.png", "--out", "out.ply", "--verbose"] + RUNS[name], cwd=d, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "Vertices / Triangles: " in r.stdout
    assert open(d / "out.ply", "rb").read() == g["out_" + name].tobytes()


@pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="reference tool not built (needs /root/reference)")
def test_live_against_the_reference_tool(tmp_path):
    """A second texture and a finer subdivision than the fixture, both tools run here."""
    from PIL import Image
    from meshopticalflow_b200 import synthetic
    v, t, uv = synthetic.uv_torus(32, 16)
    _, tb = synthetic.smooth_texture_pair(64, 40, 3)  # non-square texture
    synthetic.write_ply_textured(str(tmp_path / "m.ply"), v, t, uv)
    Image.fromarray(tb).save(tmp_path / "B.png")
    for tool, out in ((REF_TOOL, "ref.ply"), (TOOL, "mine.ply")):
        subprocess.check_call([tool, "--in", "m.ply", "--texture", "B.png", "--out", out, "--eLength", "0.03"], cwd=tmp_path, stdout=subprocess.DEVNULL)
    assert open(tmp_path / "mine.ply", "rb").read() == open(tmp_path / "ref.ply", "rb").read()


def test_usage_and_errors(workdir):
    d, _ = workdir
    r = subprocess.run([TOOL, "--in", "m.ply"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stdout and "--texture" in r.stdout  # SampleTextureToVertices.cpp:123-127
    r = subprocess.run([TOOL, "--in", "m.ply", "--texture", "A.jpg", "--out", "o.ply"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "Unrecognized image extension: jpg" in r.stderr  # :72
    r = subprocess.run([TOOL, "--IN", "m.ply", "--Texture", "A.png", "--bogus"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0 and "[WARNING] Invalid option: --bogus" in r.stderr and not os.path.exists(d / "o.ply")  # no --out: nothing written (:114)
