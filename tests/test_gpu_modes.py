"""GPU tier: the Conformal and Connection vector-field bases (--vfMode 1|2, --cMode 0|1|2) and the 6-channel DoG
blend (0 < dogWeight < 1) — SURVEY.md §8f rows 1-2 — through the C ABI and the command line, against the golden
fixtures recorded from the reference binary (tests/golden/sphere3_modes.npz) and against the CPU oracle.

Same gates as the Whitney path: flow-field relative L2 <= 1e-3 (matrix-free block-Jacobi PCG at 1e-8 relative
residual vs the reference's LDLT), colours within 1/255."""
import subprocess

import numpy as np
import pytest

from conftest import CLI_BIN, VF_MODES, rel
from meshopticalflow_b200 import api, synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.gpu

FLOW_TOL = 1e-3
COLOUR_TOL = 1.0


@pytest.fixture()
def aligner():
    al = api.Aligner(0)
    yield al
    al.close()


def _params(iterations, vf_mode=0, c_mode=0, dog_weight=1.0):
    p = api.default_params()
    p.iterations = iterations
    p.vfMode, p.cMode = vf_mode, c_mode
    p.vfSmooth = 0.0  # <= 0: the mode's default (3e-6 / 5e-7 / 1e4, OpticalFlow.cpp:1067-1069)
    p.dogWeight = dog_weight
    return p


def _golden_inputs(g):
    v = g["input_vertices_f32"].astype(np.float64) * 0.5 + g["input_vertices_f32"].astype(np.float64) * 0.5
    return v, g["triangles"], g["input_a"].astype(np.float64), g["input_b"].astype(np.float64)


@pytest.mark.parametrize("name", sorted(VF_MODES))
def test_fields_match_the_reference_golden(aligner, golden_modes, name):
    g = golden_modes
    vf_mode, c_mode = VF_MODES[name]
    v, t, a, b = _golden_inputs(g)
    al = aligner
    al.set_params(_params(4, vf_mode, c_mode))
    al.set_mesh(v, t)
    al.set_signals(a, b)
    assert al.num_coeffs == (2 * v.shape[0] if vf_mode == 1 else 2 * t.shape[0])
    for i in range(4):
        al.iterate(1)
        assert rel(al.flow(), g["%s.it%02d.tFlowField" % (name, i)]) < FLOW_TOL, i
        if vf_mode == 2:
            assert rel(al.coeffs(), g["%s.it%02d.coeffs" % (name, i)]) < FLOW_TOL, i
    assert al.stats()["lastFlowResidual"] <= 1e-8 and al.stats()["flowCgIterations"] > 0
    ca, cb = al.advect_vertices(0.5)
    assert np.abs(ca - g[name + ".advected0"]).max() < COLOUR_TOL and np.abs(cb - g[name + ".advected1"]).max() < COLOUR_TOL
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g[name + ".output_rgb"].astype(int)).max() <= 1


@pytest.mark.parametrize("vf_mode,c_mode", [(1, 0), (2, 0), (2, 2)])
def test_fields_match_the_oracle_stage_by_stage(aligner, vf_mode, c_mode):
    """4098 vertices, 3 iterations: the solved system (right-hand side, solution through P) and the flow of every iteration."""
    v, t = synthetic.octahedron_sphere(5)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 11))
    params = O.Params(iterations=3, vfMode=vf_mode, cMode=c_mode)
    st = O.init(v, t, a, b, params)
    al = aligner
    al.set_params(_params(3, vf_mode, c_mode))
    al.set_mesh(v, t)
    al.set_signals(a, b)
    sw, vw = params.sSmooth, params.vfSmooth
    for i in range(3):
        al.iterate(1)
        O.update_flow(st, sw, vw, "it.")
        sw *= params.sMultiply
        D, rhs = st.taps["it.dataTerm"], st.taps["it.rhs"]
        assert rel(al.array(api.ARR_DATA_TERM), D) < 1e-5
        A, bvec, _, _ = O.flow_system(st.whitney, D, rhs, vw)
        assert rel(al.array(api.ARR_FLOW_RHS), bvec) < 1e-4
        x = al.array(api.ARR_FLOW_SOLUTION)
        # the product's solution satisfies the ORACLE's assembled system (the Conformal one is singular: compare residuals and P x, not x)
        assert np.linalg.norm(A @ x - bvec) <= 1e-5 * np.linalg.norm(bvec)
        assert rel(st.whitney.P @ x, st.whitney.P @ st.taps["it.x"]) < FLOW_TOL
        assert rel(al.flow(), st.taps["it.tFlowField"]) < FLOW_TOL, i
    ca, cb = al.advect_vertices(0.5)
    oa, ob = O.advect_vertices(st, a, b)
    assert np.abs(ca - oa).max() < COLOUR_TOL and np.abs(cb - ob).max() < COLOUR_TOL
    with pytest.raises(api.MofError) as e:  # matrix-free: no assembled flow matrix to hand out
        al.csr(api.CSR_FLOW_SYSTEM)
    assert e.value.code == api.MOF_E_UNSUPPORTED


def test_six_channel_blend(aligner, golden_modes):
    g = golden_modes
    v, t, a, b = _golden_inputs(g)
    al = aligner
    al.set_params(_params(4, dog_weight=0.5))
    al.set_mesh(v, t)
    al.set_signals(a, b)
    params = O.Params(iterations=4, dogWeight=0.5)
    st = O.init(v, t, a, b, params)
    lo, hi = al.array(api.ARR_SIGNALS_RAW), al.array(api.ARR_SIGNALS)
    for s in range(2):  # the reference's Point<Real,6>: channels 0-2 = (1-w) raw, 3-5 = w DoG
        assert rel(lo[:, 3 * s:3 * s + 3], st.signals[s][:, :3]) < 1e-12
        assert rel(hi[:, 3 * s:3 * s + 3], st.signals[s][:, 3:]) < 1e-7
    sw, vw = params.sSmooth, params.vfSmooth
    for i in range(4):
        al.iterate(1)
        O.update_flow(st, sw, vw, "it.")
        sw *= params.sMultiply
        assert rel(al.array(api.ARR_DATA_TERM), st.taps["it.dataTerm"]) < 1e-5
        assert rel(al.flow(), st.taps["it.tFlowField"]) < FLOW_TOL, i
    assert al.stats()["smoothSolves"] == 1 + 2 * 4  # DoG + two six-channel solves per iteration
    ca, cb = al.advect_vertices(0.5)
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g["blend.output_rgb"].astype(int)).max() <= 1


def test_log_space_comparison(aligner, golden_modes):
    """mof_params.logSpace (--log): the comparison signals are log-transformed inside the library, the colours advected at the end
    are the raw ones — the reference's golden output, and the oracle stage by stage."""
    g = golden_modes
    v, t, a, b = _golden_inputs(g)
    al = aligner
    p = _params(4)
    p.logSpace = 1
    al.set_params(p)
    al.set_mesh(v, t)
    al.set_signals(a, b)
    assert rel(al.array(api.ARR_SIGNALS)[:, :3], g["log.signals0"]) < 1e-7
    al.iterate(4)
    assert rel(al.flow(), g["log.it03.tFlowField"]) < FLOW_TOL
    ca, cb = al.advect_vertices(0.5)
    assert np.abs(ca - g["log.advected0"]).max() < COLOUR_TOL
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g["log.output_rgb"].astype(int)).max() <= 1


def test_switching_bases_on_one_context(aligner):
    """Whitney -> Connection -> Whitney on the same mesh: the second Whitney run reproduces the first bit for bit, and
    a basis change without new signals is refused."""
    v, t = synthetic.octahedron_sphere(4)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 4))
    al = aligner
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.iterate(2)
    first = al.flow()
    assert al.num_coeffs == al.num_edges
    al.set_params(_params(2, 2, 1))
    with pytest.raises(api.MofError) as e:
        al.iterate(1)
    assert e.value.code == api.MOF_E_INVALID  # mof_set_signals fixes the basis
    al.set_signals(a, b)
    al.iterate(2)
    assert al.num_coeffs == 2 * t.shape[0]
    assert rel(al.flow(), al.coeffs().reshape(-1, 2)) == 0  # Connection: the prolongation is the identity
    assert rel(al.flow(), first) > 1e-3                      # a different regulariser gives a different flow
    al.set_params(api.default_params())
    al.set_signals(a, b)
    al.iterate(2)
    assert np.array_equal(al.flow(), first)


@pytest.mark.parametrize("flags,name", [(["--vfMode", "2", "--cMode", "1"], "connection1"), (["--vfMode", "1"], "conformal"), (["--dogWeight", "0.5"], "blend"), (["--log"], "log")])
def test_command_line_flags(tmp_path, golden_modes, flags, name):
    g = golden_modes
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), g["input_vertices_f32"], g["input_a"], g["triangles"], True)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), g["input_vertices_f32"], g["input_b"], g["triangles"], True)
    r = subprocess.run([CLI_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", "4"] + flags, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = synthetic.read_ply(str(tmp_path / "r.ply"))
    rgb = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(int)
    assert np.abs(rgb - g[name + ".output_rgb"].astype(int)).max() <= 1


def test_debug_dumps(tmp_path):
    """--debug: resampled.S.<i>.ply / resampled.T.<i>.ply per iteration, binary, as UpdateFlow writes them (OpticalFlow.cpp:458-465)."""
    v, t = synthetic.octahedron_sphere(3)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    synthetic.write_ply_colored(str(tmp_path / "A.ply"), v, a, t, True)
    synthetic.write_ply_colored(str(tmp_path / "B.ply"), v, b, t, True)
    r = subprocess.run([CLI_BIN, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", "2", "--debug"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    al = api.Aligner(0)
    try:
        al.set_mesh(v.astype(np.float32).astype(np.float64), t)
        al.set_signals(a.astype(np.float64), b.astype(np.float64))
        for i in range(2):
            al.iterate(1)
            res = al.array(api.ARR_RESAMPLED)
            for s, tag in enumerate("ST"):
                path = tmp_path / ("resampled.%s.%d.ply" % (tag, i))
                raw = open(path, "rb").read()
                assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 258\nproperty float x\n")
                assert len(raw) == raw.index(b"end_header\n") + 11 + 15 * 258 + 13 * 512
                out = synthetic.read_ply(str(path))
                rgb = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(int)
                expect = np.clip(res[:, 3 * s:3 * s + 3].astype(np.float32), 0, 255).astype(np.uint8).astype(int)
                assert np.abs(rgb - expect).max() <= 1
                assert np.array_equal(out["face"]["vertex_indices"], t)
    finally:
        al.close()


def test_conformal_two_cycle_preconditioner(monkeypatch):
    """Conformal basis at 16 386 vertices: PCG preconditioned by two cycles of the scalar multigrid hierarchy around the
    lumped mass (default where the hierarchy exists) against block Jacobi (MOF_CONFORMAL_MG=0): same flow, at least five
    times fewer iterations, and the smoothing solves that share the hierarchy are undisturbed."""
    v, t = synthetic.octahedron_sphere(6)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 8))
    runs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MOF_CONFORMAL_MG", mode)
        al = api.Aligner(0)
        try:
            al.set_params(_params(2, 1))
            al.set_mesh(v, t)
            al.set_signals(a, b)
            al.iterate(2)
            s = al.stats()
            assert s["lastFlowResidual"] <= 1e-8 and s["lastSmoothResidual"] <= 1.01e-10
            runs[mode] = (al.flow(), s["flowCgIterations"], s["smoothCgIterations"], al.array(api.ARR_SMOOTHED))
        finally:
            al.close()
    assert rel(runs["1"][0], runs["0"][0]) < 1e-5
    assert runs["1"][1] * 5 < runs["0"][1], (runs["1"][1], runs["0"][1])
    assert runs["1"][2] == runs["0"][2] and rel(runs["1"][3], runs["0"][3]) < 1e-9


def test_conformal_basis_at_65k_vertices():
    """The Conformal flow solve at 65 538 vertices: with each inverse of the two-cycle preconditioner sharpened by a Chebyshev polynomial
    around the cycle (degree and interval from the cycle's measured contraction) the PCG needs at most 200 iterations per solve
    (875 with one cycle per inverse) and reaches the tolerance."""
    v, t = synthetic.octahedron_sphere(7)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 8))
    al = api.Aligner(0)
    try:
        al.set_params(_params(2, 1))
        al.set_mesh(v, t)
        al.set_signals(a, b)
        al.iterate(2)
        s = al.stats()
        assert s["lastFlowResidual"] <= 1e-8 and s["flowSolves"] == 2
        assert s["flowCgIterations"] <= 2 * 200, s["flowCgIterations"]
    finally:
        al.close()
