"""GPU tier: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same seeded
inputs, against the golden fixtures recorded from the reference, and — at the 1M-vertex size of
BASELINE.json — through size-independent properties.

Gates (BASELINE.json north_star): adjacency and sparsity patterns bit-exact; flow-field relative L2 <= 1e-3
(PCG at 1e-8 relative residual vs the direct solve); colours within 1/255."""
import numpy as np
import pytest

from conftest import colour_outliers, csr_from_golden, rel
from meshopticalflow_b200 import api, synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.gpu

FLOW_TOL = 1e-3       # north_star
COLOUR_TOL = 1.0      # 1/255 on the 0..255 scale


@pytest.fixture()
def aligner():
    al = api.Aligner(0)
    yield al
    al.close()


def _pattern_equal(a, b):
    return np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


def test_every_stage_matches_the_oracle(aligner):
    v, t = synthetic.octahedron_sphere(4)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    params = O.Params()
    st = O.init(v, t, a, b, params)
    al = aligner
    al.set_mesh(v, t)
    # a1-a3: metric, adjacency (bit-exact), edge transforms
    assert rel(al.array(api.ARR_METRIC), st.g) < 1e-13
    assert rel(al.array(api.ARR_AREA), st.area) < 1e-13
    assert np.array_equal(al.array(api.ARR_OPPOSITE), st.opp)
    assert rel(al.array(api.ARR_XFORM_LINEAR), st.lin) < 1e-12 and rel(al.array(api.ARR_XFORM_CONSTANT), st.cst) < 1e-12
    # a4, a8: patterns bit-exact, values to round-off
    for which, ref in ((api.CSR_SCALAR_MASS, st.M), (api.CSR_SCALAR_STIFFNESS, st.S), (api.CSR_WHITNEY_SMOOTH, st.whitney.S)):
        m = al.csr(which)
        assert _pattern_equal(m, ref)
        assert rel(m.data, ref.data) < 1e-12
    # a6: Whitney numbering bit-exact
    assert np.array_equal(al.array(api.ARR_REDUCED_EDGE), st.whitney.reduced)
    assert np.array_equal(al.array(api.ARR_EXPANDED_EDGE), st.whitney.expanded)
    assert np.array_equal(al.array(api.ARR_POSITIVE_EDGE), st.whitney.positive.astype(np.int32))
    # a5: DoG
    al.set_signals(a, b)
    sig = al.array(api.ARR_SIGNALS)
    assert rel(sig[:, :3], st.signals[0]) < 1e-7 and rel(sig[:, 3:], st.signals[1]) < 1e-7
    # a9-a13, three iterations
    sw, vw = params.sSmooth, params.vfSmooth
    for i in range(3):
        al.iterate(1)
        O.update_flow(st, sw, vw, "it.")
        sw *= params.sMultiply
        sm, rs = al.array(api.ARR_SMOOTHED), al.array(api.ARR_RESAMPLED)
        assert rel(sm[:, :3], st.taps["it.smoothed0"]) < 1e-7 and rel(sm[:, 3:], st.taps["it.smoothed1"]) < 1e-7
        assert rel(rs[:, :3], st.taps["it.resampled0"]) < 1e-6 and rel(rs[:, 3:], st.taps["it.resampled1"]) < 1e-6
        assert rel(al.array(api.ARR_DATA_TERM), st.taps["it.dataTerm"]) < 1e-6
        assert rel(al.array(api.ARR_DATA_RHS), st.taps["it.rhs"]) < 1e-4
        assert rel(al.array(api.ARR_FLOW_SOLUTION), st.taps["it.x"]) < FLOW_TOL
        assert rel(al.flow(), st.taps["it.tFlowField"]) < FLOW_TOL
    # the flow system itself: A x = b to the PCG tolerance, A symmetric
    A, x, rhs = al.csr(api.CSR_FLOW_SYSTEM), al.array(api.ARR_FLOW_SOLUTION), al.array(api.ARR_FLOW_RHS)
    assert np.linalg.norm(A @ x - rhs) <= 1.01e-8 * np.linalg.norm(rhs)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    # a14: final advection of the raw colours
    ca, cb = al.advect_vertices(0.5)
    oa, ob = O.advect_vertices(st, a, b)
    assert np.abs(ca - oa).max() < 1e-3 and np.abs(cb - ob).max() < 1e-3
    assert al.stats()["kernelLaunches"] > 0


def test_vertex_alignment_matches_the_reference_golden(aligner, golden_sphere):
    g = golden_sphere
    v = g["input_vertices_f32"].astype(np.float64) * 0.5 + g["input_vertices_f32"].astype(np.float64) * 0.5
    al = aligner
    al.set_mesh(v, g["triangles"])
    assert np.array_equal(al.array(api.ARR_OPPOSITE), g["oppositeEdge"])
    assert np.array_equal(al.array(api.ARR_REDUCED_EDGE), g["reducedEdgeIndex"])
    E = al.num_edges
    ref = csr_from_golden(g, "smoothOperator", (E, E))
    S = al.csr(api.CSR_WHITNEY_SMOOTH)
    assert _pattern_equal(S, ref) and rel(S.data, ref.data) < 1e-12
    M = al.csr(api.CSR_SCALAR_MASS)
    refM = csr_from_golden(g, "sMass", (v.shape[0], v.shape[0]))
    refM.sort_indices()
    assert _pattern_equal(M, refM) and rel(M.data, refM.data) < 1e-12
    al.set_signals(g["input_a"].astype(np.float64), g["input_b"].astype(np.float64))
    for i in range(10):
        al.iterate(1)
        assert rel(al.flow(), g["it%02d.tFlowField" % i]) < FLOW_TOL, i
    ca, cb = al.advect_vertices(0.5)
    assert np.abs(ca - g["advected0"]).max() < COLOUR_TOL and np.abs(cb - g["advected1"]).max() < COLOUR_TOL
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g["output_rgb"].astype(int)).max() <= 1


def test_texture_alignment_matches_the_reference_golden(aligner, golden_torus):
    g = golden_torus
    ta, tb = g["input_tex_a"], g["input_tex_b"]
    v, t, uv = g["vertices"], g["triangles"], g["triangleTextures"].reshape(-1, 6)  # the reference's subdivided mesh
    sig = [O.sample_texture_to_vertices(t, uv, v.shape[0], tex) for tex in (ta, tb)]
    al = aligner
    al.set_mesh(v, t)
    assert np.array_equal(al.array(api.ARR_OPPOSITE), g["oppositeEdge"])
    al.set_signals(sig[0], sig[1])
    s6 = al.array(api.ARR_SIGNALS)
    assert rel(s6[:, :3], g["signals0"]) < 1e-7 and rel(s6[:, 3:], g["signals1"]) < 1e-7
    for i in range(10):
        al.iterate(1)
        assert rel(al.flow(), g["it%02d.tFlowField" % i]) < FLOW_TOL, i
    al.set_texture_map(48, 48, g["textureSource_tIdx"], g["textureSource_p"], uv, ta, tb)
    oa, ob = al.advect_texels(0.5)
    # a texel whose sample point sits on an edge can land on either side: allow 0.2 % of values
    assert colour_outliers(oa, g["advected0"], COLOUR_TOL) < 2e-3 and colour_outliers(ob, g["advected1"], COLOUR_TOL) < 2e-3
    pixels = O.to_uchar_png((oa + ob) / 2.0).reshape(48, 48, 3)[::-1]
    assert colour_outliers(pixels, g["output_pixels"], 1.0) < 2e-3


def test_pcg_against_a_direct_solve(aligner):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(5)
    n = 5003  # not a multiple of the tile size
    rows = rng.integers(0, n, 6 * n)
    cols = rng.integers(0, n, 6 * n)
    w = rng.uniform(0.1, 1.0, 6 * n)
    W = sp.coo_matrix((w, (rows, cols)), shape=(n, n)).tocsr()
    W = W + W.T
    A = (sp.diags(np.asarray(W.sum(1)).ravel() + rng.uniform(0.01, 0.1, n)) - W).tocsr()
    b = rng.standard_normal(n)
    x, iters, relres = aligner.pcg_solve_csr(A, b, 1e-10)
    ref = spla.spsolve(A.tocsc(), b)
    assert relres <= 1.01e-10 and iters > 0
    assert rel(x, ref) < 1e-7


def test_error_paths(aligner):
    al = aligner
    with pytest.raises(api.MofError) as e:
        al.set_signals(np.zeros((0, 3)), np.zeros((0, 3)))
    assert e.value.code == api.MOF_E_INVALID
    with pytest.raises(api.MofError) as e:  # half-edge 0->1 used twice (FEM.inl:599)
        al.set_mesh(np.eye(4, 3), np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32))
    assert e.value.code == api.MOF_E_MESH and "Edge is occupied" in e.value.message
    with pytest.raises(api.MofError) as e:  # open mesh (FEM.inl:554)
        al.set_mesh(np.eye(3), np.array([[0, 1, 2]], dtype=np.int32))
    assert e.value.code == api.MOF_E_MESH and "Boundary edge" in e.value.message
    with pytest.raises(api.MofError) as e:
        al.set_mesh(np.eye(3), np.array([[0, 1, 7]], dtype=np.int32))
    assert e.value.code == api.MOF_E_INVALID
    # smallest closed mesh: a tetrahedron runs end to end
    v = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=np.float64)
    t = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], dtype=np.int32)
    al.set_mesh(v, t)
    with pytest.raises(api.MofError) as e:
        al.set_signals(np.zeros((4, 6)), np.zeros((4, 6)))
    assert e.value.code == api.MOF_E_UNSUPPORTED
    p = api.default_params()
    p.vfMode = 3  # "ERROR: Unsupported vector field!" (OpticalFlow.cpp:867)
    with pytest.raises(api.MofError) as e:
        al.set_params(p)
    assert e.value.code == api.MOF_E_INVALID and "Unsupported vector field" in e.value.message
    col = np.arange(12, dtype=np.float64).reshape(4, 3) * 10
    col_b = col + np.array([[1.0, -2.0, 0.5], [0.0, 1.0, -1.0], [2.0, 0.0, 1.0], [-1.0, 1.0, 0.0]])
    al.set_signals(col, col_b)
    al.iterate(1)  # one iteration: four triangles cannot carry a stable multi-iteration flow
    st = O.init(v, t, col, col_b, O.Params())
    O.iterate(st, O.Params(iterations=1))
    assert rel(al.flow(), st.tfield) < FLOW_TOL


def test_identical_signals_give_zero_flow_and_swapping_flips_it(aligner):
    v, t = synthetic.octahedron_sphere(5)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 2))
    al = aligner
    al.set_mesh(v, t)
    al.set_signals(a, a)
    al.iterate(2)
    assert np.abs(al.flow()).max() == 0.0
    ca, cb = al.advect_vertices(0.5)
    assert np.array_equal(ca, cb)
    al.set_signals(a, b)
    al.iterate(3)
    f_ab = al.flow()
    al.set_signals(b, a)
    al.iterate(3)
    assert rel(al.flow(), -f_ab) < 1e-5  # the halfway formulation is antisymmetric in the pair


def test_multigrid_and_jacobi_preconditioned_solves_agree():
    """The multigrid-PCG (default) and the plain Jacobi-PCG kernel (MOF_FLOW_MG=0, MOF_SCALAR_MG=0) solve the same
    systems to the same residual: same flow, far fewer iterations."""
    import os
    v, t = synthetic.octahedron_sphere(6)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 1))
    flows, iters = {}, {}
    saved = {k: os.environ.get(k) for k in ("MOF_FLOW_MG", "MOF_SCALAR_MG")}
    try:
        for mode in ("1", "0"):
            os.environ["MOF_FLOW_MG"] = os.environ["MOF_SCALAR_MG"] = mode  # read when the mesh is set
            al = api.Aligner(0)
            try:
                al.set_mesh(v, t)
                al.set_signals(a, b)
                al.iterate(2)
                s = al.stats()
                assert s["lastFlowResidual"] <= 1.01e-8 and s["lastSmoothResidual"] <= 1.01e-10
                flows[mode], iters[mode] = al.flow(), (s["flowCgIterations"], s["smoothCgIterations"])
            finally:
                al.close()
    finally:
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    assert rel(flows["1"], flows["0"]) < 1e-6
    assert iters["1"][0] * 3 < iters["0"][0] and iters["1"][1] * 3 < iters["0"][1], iters


def test_texel_frame_sequence_matches_the_oracle(aligner, golden_torus):
    """InputTextureData::flow(frames) (OpticalFlow.cpp:517-539): the sample points carried along the flow in frames - 1 steps, a
    texture fetch after each; frame 0 and uncovered texels are the flipped input; the last frame of a 2-frame sequence is the
    whole flow walked in one go with the coarser minimum step."""
    g = golden_torus
    ta, tb = g["input_tex_a"], g["input_tex_b"]
    v, t, uv = g["vertices"], g["triangles"], g["triangleTextures"].reshape(-1, 6)
    sig = [O.sample_texture_to_vertices(t, uv, v.shape[0], tex) for tex in (ta, tb)]
    al = aligner
    al.set_mesh(v, t)
    al.set_signals(sig[0], sig[1])
    al.iterate(4)
    srcT, srcP = g["textureSource_tIdx"], g["textureSource_p"]
    al.set_texture_map(48, 48, srcT, srcP, uv, ta, tb)
    flow = al.flow()
    opp, lin, cst, gm = (al.array(k) for k in (api.ARR_OPPOSITE, api.ARR_XFORM_LINEAR, api.ARR_XFORM_CONSTANT, api.ARR_METRIC))
    for frames, bilinear in ((5, True), (2, False)):
        fa, fb = al.advect_texels_frames(frames, bilinear)
        for mine, tex, sign in ((fa, ta, -1.0), (fb, tb, 1.0)):
            ref = O.advect_texels_frames(48, 48, frames, srcT, srcP, opp, lin, cst, gm, flow, uv, tex, sign, bilinear)
            assert mine.shape == ref.shape == (frames, 48 * 48, 3)
            assert np.array_equal(mine[0], tex[::-1].reshape(-1, 3).astype(np.float64))
            assert np.array_equal(mine[:, srcT == -1], ref[:, srcT == -1])
            assert colour_outliers(mine, ref, 1e-6) < 2e-3  # walk ties aside, the same numbers
    with pytest.raises(api.MofError) as e:
        al.advect_texels_frames(1)
    assert e.value.code == api.MOF_E_INVALID


def test_irregular_mesh_matches_the_oracle(aligner):
    """Vertices jittered along the surface (triangle areas vary by ~10x, no flips; aggregates become ragged)."""
    v, t = synthetic.octahedron_sphere(4)
    rng = np.random.default_rng(11)
    v = v + 0.12 * np.sqrt(4 * np.pi / v.shape[0]) * rng.standard_normal(v.shape)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v *= (1.0 + 0.3 * v[:, :1] * v[:, 1:2])  # and no longer a sphere
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v / np.linalg.norm(v, axis=1, keepdims=True), 3))
    params = O.Params(iterations=3)
    st = O.init(v, t, a, b, params)
    O.iterate(st, params)
    al = aligner
    p = api.default_params()
    p.iterations = 3
    al.set_params(p)
    al.set_mesh(v, t)
    assert np.array_equal(al.array(api.ARR_OPPOSITE), st.opp)
    al.set_signals(a, b)
    al.iterate(3)
    assert rel(al.flow(), st.tfield) < FLOW_TOL
    ca, cb = al.advect_vertices(0.5)
    oa, ob = O.advect_vertices(st, a, b)
    assert colour_outliers(ca, oa, COLOUR_TOL) < 2e-3 and colour_outliers(cb, ob, COLOUR_TOL) < 2e-3


@pytest.mark.timeout(900)
def test_million_vertex_properties(aligner):
    """BASELINE.json config 3 size (1 048 578 V). The oracle cannot run here in seconds, so: topology
    identities, determinism, PCG residuals, and the alignment actually aligning."""
    v, t = synthetic.octahedron_sphere(9)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 0))
    V, T = v.shape[0], t.shape[0]
    al = aligner
    p = api.default_params()
    p.iterations = 2
    al.set_params(p)
    al.set_mesh(v, t)
    E = al.num_edges
    assert (V, T, E) == (1048578, 2097152, 3145728) and V - E + T == 2  # Euler, genus 0
    opp = al.array(api.ARR_OPPOSITE)
    assert opp.min() >= 0 and np.array_equal(opp[opp], np.arange(3 * T, dtype=np.int32))  # an involution without fixed points
    red = al.array(api.ARR_REDUCED_EDGE)
    assert np.array_equal(red, red[opp]) and np.array_equal(np.bincount(red, minlength=E), np.full(E, 2))
    assert abs(al.array(api.ARR_AREA).sum() - 1.0) < 1e-12  # makeUnitArea
    s = al.stats()
    # row e=(a,b) of the Whitney system has deg(a)+deg(b)-1 entries: sum_v deg(v)^2 - E, with 6 valence-4 vertices
    assert s["flowRows"] == E and s["flowNnz"] == 6 * 16 + (V - 6) * 36 - E
    al.set_signals(a, b)
    al.iterate(2)
    s = al.stats()
    assert s["lastFlowResidual"] <= 1.01e-8 and s["lastSmoothResidual"] <= 1.01e-10
    assert 0 < s["flowCgIterations"] <= 2 * 400  # the multigrid preconditioner is in use (Jacobi-PCG needs ~4 500 per solve here)
    f1 = al.flow()
    ca, cb = al.advect_vertices(0.5)
    assert np.abs(ca - cb).mean() < 0.5 * np.abs(a - b).mean()  # two iterations already halve the mismatch
    assert ca.min() >= -1e-9 and ca.max() <= 255 + 1e-9       # resampling is a convex combination
    # bitwise reproducible: same inputs, same flow
    al.set_signals(a, b)
    al.iterate(2)
    assert np.array_equal(al.flow(), f1)


def test_smoothing_ahead_on_a_second_stream_changes_nothing(monkeypatch):
    """The next iteration's smoothing solve runs on a second stream (a worker thread) under the flow solve. It is the
    same computation on the same inputs: with it and without it (MOF_SMOOTH_AHEAD=0) every flow, the smoothed signals
    and the iteration counts are identical, bit for bit."""
    v, t = synthetic.octahedron_sphere(7)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 6))
    runs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MOF_SMOOTH_AHEAD", mode)
        al = api.Aligner(0)
        try:
            p = api.default_params()
            p.iterations = 5
            al.set_params(p)
            al.set_mesh(v, t)
            al.set_signals(a, b)
            flows = []
            for i in range(5):  # one call per iteration, like the command line
                al.iterate(1)
                flows.append(al.flow())
            smoothed = al.array(api.ARR_SMOOTHED)
            al.set_signals(b, a)  # a solve in flight or not, new signals start clean
            al.iterate(5)
            s = al.stats()
            runs[mode] = (flows, smoothed, al.flow(), s["smoothCgIterations"], s["flowCgIterations"], s["smoothSolves"])
        finally:
            al.close()
    for x, y in zip(runs["1"][0], runs["0"][0]):
        assert np.array_equal(x, y)
    assert np.array_equal(runs["1"][1], runs["0"][1]) and np.array_equal(runs["1"][2], runs["0"][2])
    assert runs["1"][3:] == runs["0"][3:]
    assert runs["1"][5] == 2 * (1 + 5)


def test_programmatic_dependent_launches_change_nothing(monkeypatch):
    """MOF_PDL (mof_internal.cuh: every solver kernel starts with griddepcontrol.wait and is launched with programmatic stream
    serialisation, inside the captured PCG graph too) only changes WHEN a kernel's CTAs are scheduled: with it and without it the
    flows, the advected colours and the iteration counts are identical, bit for bit — also through the replay loop (MOF_MG_WHILE=0)."""
    v, t = synthetic.octahedron_sphere(7)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 8))
    runs = {}
    for mode in ("pdl", "plain", "pdl-replay"):
        monkeypatch.setenv("MOF_PDL", "0" if mode == "plain" else "1")
        monkeypatch.setenv("MOF_MG_WHILE", "0" if mode == "pdl-replay" else "1")
        al = api.Aligner(0)
        try:
            p = api.default_params()
            p.iterations = 4
            al.set_params(p)
            al.set_mesh(v, t)
            al.set_signals(a, b)
            al.iterate(4)
            s = al.stats()
            runs[mode] = (al.flow(), al.advect_vertices(0.5), s["flowCgIterations"], s["smoothCgIterations"])
        finally:
            al.close()
    for mode in ("plain", "pdl-replay"):
        assert np.array_equal(runs["pdl"][0], runs[mode][0]), mode
        assert np.array_equal(runs["pdl"][1][0], runs[mode][1][0]) and np.array_equal(runs["pdl"][1][1], runs[mode][1][1]), mode
        assert runs["pdl"][2:] == runs[mode][2:], mode
