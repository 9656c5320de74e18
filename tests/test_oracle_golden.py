"""CPU tier: the oracle (oracle/mof_oracle.py + walk.c) against the golden fixtures recorded from the
reference binary (tests/golden/make_golden.py). This is what pins the oracle on a box without /root/reference."""
import numpy as np

import pytest

from conftest import VF_MODES, colour_outliers, csr_from_golden, rel, sorted_csr
from oracle import mof_oracle as O


def _sphere_inputs(g):
    # OpticalFlow.cpp:765: geometry = mean of the two (identical) float meshes, in double
    v = g["input_vertices_f32"].astype(np.float64) * 0.5 + g["input_vertices_f32"].astype(np.float64) * 0.5
    return v, g["triangles"], g["input_a"].astype(np.float64), g["input_b"].astype(np.float64)


def test_setup_matches_reference(golden_sphere):
    g = golden_sphere
    v, t, a, b = _sphere_inputs(g)
    st = O.init(v, t, a, b, O.Params())
    assert np.array_equal(v, g["vertices"])
    # adjacency and numbering: bit-exact
    assert np.array_equal(st.opp, g["oppositeEdge"])
    assert np.array_equal(st.whitney.reduced, g["reducedEdgeIndex"])
    assert np.array_equal(st.whitney.expanded, g["expandedEdgeIndex"])
    assert np.array_equal(st.whitney.positive.astype(np.int32), g["positiveOrientedEdge"])
    # metric and edge transforms (golden stores matrices column-major)
    assert rel(st.g, g["g"][:, [0, 1, 3]]) < 1e-13
    assert rel(st.lin, g["xform_linear"][:, [0, 2, 1, 3]]) < 1e-12
    assert rel(st.cst, g["xform_constant"]) < 1e-12
    # sparsity patterns bit-exact once columns are sorted, values to round-off
    nv = v.shape[0]
    for mine, name in ((st.M, "sMass"), (st.S, "sStiffness")):
        ref = sorted_csr(csr_from_golden({**g, name + ".rowptr": g["sMass.rowptr"], name + ".col": g["sMass.col"]}, name, (nv, nv)))
        assert np.array_equal(mine.indptr, ref.indptr) and np.array_equal(mine.indices, ref.indices)
        assert rel(mine.data, ref.data) < 1e-12
    E = st.whitney.expanded.size
    ref = csr_from_golden(g, "smoothOperator", (E, E))
    assert np.array_equal(st.whitney.S.indptr, ref.indptr) and np.array_equal(st.whitney.S.indices, ref.indices)
    assert rel(st.whitney.S.data, ref.data) < 1e-12
    # DoG-normalised comparison signals
    assert rel(st.signals[0], g["signals0"]) < 1e-10 and rel(st.signals[1], g["signals1"]) < 1e-10


def test_vertex_alignment_matches_reference(golden_sphere):
    g = golden_sphere
    v, t, a, b = _sphere_inputs(g)
    st = O.init(v, t, a, b, O.Params())
    O.iterate(st, O.Params(), taps=True)
    for i in range(10):
        pre = "it%02d." % i
        assert rel(st.taps[pre + "smoothed0"], g[pre + "smoothed0"]) < 1e-9
        assert rel(st.taps[pre + "resampled1"], g[pre + "resampled1"]) < 1e-9
        assert rel(st.taps[pre + "dataTerm"], g[pre + "dataTerm"][:, [0, 1, 3]]) < 1e-8
        assert rel(st.taps[pre + "x"], g[pre + "x"]) < 1e-7
        assert rel(st.taps[pre + "tFlowField"], g[pre + "tFlowField"]) < 1e-8, i  # north_star gate is 1e-3
    ca, cb = O.advect_vertices(st, a, b)
    assert np.abs(ca - g["advected0"]).max() < 1e-6 and np.abs(cb - g["advected1"]).max() < 1e-6
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g["output_rgb"].astype(int)).max() <= 1  # 1/255


def test_texture_alignment_matches_reference(golden_torus):
    g = golden_torus
    params = O.Params(eLength=float(np.float32(0.08)))
    # the textures as the reference decoded them (RGBA input: alpha dropped)
    ta, tb = g["texture0"].reshape(48, 48, 3).astype(np.uint8), g["texture1"].reshape(48, 48, 3).astype(np.uint8)
    assert np.array_equal(ta, g["input_tex_a"]) and np.array_equal(tb, g["input_tex_b"])
    st, res = O.align_texture(g["input_vertices_f32"], g["input_triangles"], g["input_uv"], ta, tb, params, taps=True)
    # subdivision: same vertices in the same order, same triangles, same uv
    assert np.array_equal(res["vertices"].astype(np.float64), g["vertices"])
    assert np.array_equal(res["triangles"], g["triangles"])
    assert rel(res["tri_uv"].reshape(-1, 2), g["triangleTextures"]) < 1e-15
    assert np.array_equal(st.opp, g["oppositeEdge"])
    assert rel(st.signals[0], g["signals0"]) < 1e-10
    # texel map: same triangles, same points
    assert np.array_equal(res["srcT"], g["textureSource_tIdx"])
    covered = res["srcT"] >= 0
    assert np.abs(res["srcP"][covered] - g["textureSource_p"][covered]).max() < 1e-10
    for i in range(10):
        assert rel(st.taps["it%02d.tFlowField" % i], g["it%02d.tFlowField" % i]) < 1e-8
    for s in range(2):
        assert colour_outliers(res["advected"][s], g["advected%d" % s], 1e-6) < 2e-3
    assert colour_outliers(res["pixels"], g["output_pixels"], 1.0) < 2e-3


@pytest.mark.parametrize("name", sorted(VF_MODES))
def test_conformal_and_connection_fields_match_reference(golden_modes, name):
    """--vfMode 1 | 2 (--cMode 0|1|2): smoothness operator, per-iteration coefficients and flow, output colours."""
    g = golden_modes
    vf_mode, c_mode = VF_MODES[name]
    v, t, a, b = _sphere_inputs(g)
    params = O.Params(iterations=4, vfMode=vf_mode, cMode=c_mode)
    assert params.vfSmooth == (3e-6, 5e-7, 1e4)[vf_mode]
    st = O.init(v, t, a, b, params)
    n = 2 * v.shape[0] if vf_mode == 1 else 2 * t.shape[0]
    assert st.coeffs.size == n
    ref = sorted_csr(csr_from_golden(g, name + ".smoothOperator", (n, n)))
    ref.sum_duplicates()
    S = st.whitney.S
    assert abs(S - ref).max() < 1e-12 * abs(ref).max()
    assert abs(ref - ref.T).max() < 1e-12 * abs(ref).max()  # what makes CG applicable
    O.iterate(st, params, taps=True)
    for i in range(4):
        assert rel(st.taps["it%02d.tFlowField" % i], g["%s.it%02d.tFlowField" % (name, i)]) < 1e-8, i
        if vf_mode == 2:  # the Conformal system is singular (constants): coefficients are unique only up to its null space
            assert rel(st.taps["it%02d.coeffs" % i], g["%s.it%02d.coeffs" % (name, i)]) < 1e-8, i
    ca, cb = O.advect_vertices(st, a, b)
    assert np.abs(ca - g[name + ".advected0"]).max() < 1e-6 and np.abs(cb - g[name + ".advected1"]).max() < 1e-6
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g[name + ".output_rgb"].astype(int)).max() <= 1


def test_six_channel_dog_blend_matches_reference(golden_modes):
    """0 < dogWeight < 1: _main<double,6> (OpticalFlow.cpp:1114), signals = ((1-w) raw, w DoG)."""
    g = golden_modes
    v, t, a, b = _sphere_inputs(g)
    params = O.Params(iterations=4, dogWeight=0.5)
    st, blended = O.align_vertices(v, t, a, b, params)
    assert st.signals[0].shape[1] == 6
    assert np.abs(O.to_uchar_ply(blended).astype(int) - g["blend.output_rgb"].astype(int)).max() <= 1


def test_log_space_comparison_matches_reference(golden_modes):
    """--log (OpticalFlow.cpp:821) transforms the comparison signals only: the flow follows the log signals, the output is the
    RAW colours advected along it (:482-489, :1049-1054)."""
    g = golden_modes
    v, t = g["input_vertices_f32"].astype(np.float64), g["triangles"]
    a, b = g["input_a"].astype(np.float64), g["input_b"].astype(np.float64)
    st, out = O.align_vertices(v, t, a, b, O.Params(iterations=4, logSpace=True))
    assert rel(st.signals[0], g["log.signals0"]) < 1e-9
    assert rel(st.tfield, g["log.it03.tFlowField"]) < 1e-9
    assert np.abs(O.to_uchar_ply(out).astype(int) - g["log.output_rgb"].astype(int)).max() <= 1
    plain, _ = O.align_vertices(v, t, a, b, O.Params(iterations=4))
    assert rel(st.tfield, plain.tfield) > 1e-2  # the flag matters on this input: another flow ...
    logged = np.log(np.maximum(1.0, a)) * 255.0 / np.log(255.0)
    assert np.abs(out - a).mean() < 0.25 * np.abs(out - logged).mean()  # ... and an output in the raw colours' range, not the logarithms'


def test_walk_edge_cases():
    """flow() on a tiny closed mesh: zero field, zero time, and a walk long enough to wrap around."""
    v = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=np.float64)
    t = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], dtype=np.int32)
    col = np.arange(12, dtype=np.float64).reshape(4, 3)
    st = O.init(v, t, col, col[::-1].copy(), O.Params(dogWeight=0.0))
    # zero field: every triangle samples its own centroid, each vertex gets the mean over its 3 triangles
    out = O.resample_signal(st.triangles, st.opp, st.lin, st.cst, st.g, np.zeros((4, 2)), col, -0.5)
    cent = col[t].mean(1)
    expect = np.stack([cent[(t == k).any(1)].mean(0) for k in range(4)])
    assert np.abs(out - expect).max() < 1e-12
    # a constant-coordinate field, long walk: stays finite and inside the signal's range
    tf = np.tile(np.array([0.3, 0.1]), (4, 1))
    out = O.resample_signal(st.triangles, st.opp, st.lin, st.cst, st.g, tf, col, 5.0)
    assert np.isfinite(out).all() and out.min() >= col.min() - 1e-9 and out.max() <= col.max() + 1e-9


def test_non_manifold_and_open_meshes_are_rejected():
    import pytest
    t = np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32)  # half-edge 0->1 used twice
    with pytest.raises(ValueError, match="Edge is occupied"):
        O.opposite_half_edges(t)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    with pytest.raises(ValueError, match="Boundary edge"):
        O.init(v, np.array([[0, 1, 2]], dtype=np.int32), np.zeros((3, 3)), np.zeros((3, 3)), O.Params())
