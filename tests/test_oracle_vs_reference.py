"""CPU tier, build container only: the oracle against the LIVE reference binary (oracle/_ref/OpticalFlow_ref)
on a 1026-vertex sphere. Skipped where the binary does not exist (it is built from /root/reference)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REF_BIN, rel
from meshopticalflow_b200 import synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not built (oracle/ref/build_ref.sh needs /root/reference)")


def test_sphere_level4(tmp_path):
    v, t = synthetic.octahedron_sphere(4)
    a, b = synthetic.smooth_rgb_pair(v, 3)
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), v, a, t)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), v, b, t)
    subprocess.check_call([REF_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", "4", "--tap", "tap"], cwd=tmp_path, stdout=subprocess.DEVNULL)
    tap = lambda n: np.load(tmp_path / "tap" / (n + ".npy"))
    vf = v.astype(np.float32).astype(np.float64)
    st = O.init(vf, t, a.astype(np.float64), b.astype(np.float64), O.Params())
    assert np.array_equal(st.opp, tap("oppositeEdge"))
    assert np.array_equal(st.whitney.reduced, tap("reducedEdgeIndex"))
    assert rel(st.signals[1], tap("signals1")) < 1e-10
    O.iterate(st, O.Params(iterations=4), taps=True)
    for i in range(4):
        assert rel(st.taps["it%02d.tFlowField" % i], tap("it%02d.tFlowField" % i)) < 1e-8
    ca, cb = O.advect_vertices(st, a.astype(np.float64), b.astype(np.float64))
    assert np.abs(ca - tap("advected0")).max() < 1e-6
    out = synthetic.read_ply(str(tmp_path / "r.ply"))
    rgb = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1)
    assert np.abs(O.to_uchar_ply((ca + cb) / 2.0).astype(int) - rgb.astype(int)).max() <= 1


@pytest.mark.parametrize("flags,kw", [(["--vfMode", "1"], dict(vfMode=1)), (["--vfMode", "2", "--cMode", "1"], dict(vfMode=2, cMode=1)),
                                      (["--dogWeight", "0.25"], dict(dogWeight=0.25))])
def test_other_bases_and_the_dog_blend(tmp_path, flags, kw):
    """Conformal / Connection fields and the 6-channel blend on the 1026-vertex sphere, whole program."""
    v, t = synthetic.octahedron_sphere(4)
    a, b = synthetic.smooth_rgb_pair(v, 5)
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), v, a, t)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), v, b, t)
    subprocess.check_call([REF_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", "3"] + flags, cwd=tmp_path, stdout=subprocess.DEVNULL)
    vf = v.astype(np.float32).astype(np.float64)
    _, blended = O.align_vertices(vf, t, a.astype(np.float64), b.astype(np.float64), O.Params(iterations=3, **kw))
    out = synthetic.read_ply(str(tmp_path / "r.ply"))
    rgb = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1)
    assert np.abs(O.to_uchar_ply(blended).astype(int) - rgb.astype(int)).max() <= 1
