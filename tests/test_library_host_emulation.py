"""CPU tier: the WHOLE library behind include/mof_b200.h — every .cu file of meshopticalflow_b200/csrc except dist.cu (NCCL),
kernels, host drivers and the C ABI itself — compiled for the HOST by tests/host_emulation (thread blocks on fibers, counted
barriers, warp shuffles, atomics, stream capture and graph replay; the cooperative Jacobi-PCG kernel as one CTA) and driven
through the same ctypes binding the GPU tier uses. What the GPU tier asserts on a B200 is asserted here on the CPU, at sizes
the emulation finishes in seconds: the reference's golden fixtures, the numpy checker stage by stage, multigrid-PCG against
Jacobi-PCG, the three bases, the 6-channel blend, the texture configuration end to end.

This is a build of the PRODUCT'S sources for testing, not a CPU path of the product: nothing under meshopticalflow_b200/
can load it (api.LIB_PATH is patched by the fixture below), and mof_create of the real library still fails without a GPU."""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, VF_MODES, colour_outliers, csr_from_golden, rel
from meshopticalflow_b200 import api, synthetic
from oracle import mof_oracle as O

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
UNITS = [0, 1, 2, 3, 4, 5, 6, 8, 9]  # library_emul.cpp: one translation unit per .cu file (7 = dist.cu, stubbed here)


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    """api, bound to the emulated build for the duration of this module."""
    out = tmp_path_factory.mktemp("library_emul")
    # MOF_EMUL_CXXFLAGS="-O1 -g -fsanitize=address" (with LD_PRELOAD=libasan.so, ASAN_OPTIONS=detect_leaks=0) turns every "device"
    # buffer overrun of every kernel into a test failure: "device" memory is malloc'd
    extra = os.environ.get("MOF_EMUL_CXXFLAGS", "-O2").split()
    base = ["g++"] + extra + ["-std=c++17", "-fPIC", "-c", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-I.", "-w"]
    jobs = [base + ["-DEMUL_UNIT=%d" % u, "-o", str(out / ("unit%d.o" % u)), "library_emul.cpp"] for u in UNITS]
    jobs += [base + ["-o", str(out / "dist_stub.o"), "dist_stub.cpp"], base + ["-o", str(out / "runtime.o"), "emul_runtime.cpp"]]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda cmd: subprocess.check_call(cmd, cwd=EMU_DIR), jobs))
    lib = str(out / "libmof_emul.so")
    subprocess.check_call(["g++", "-shared"] + [f for f in extra if f.startswith("-fsanitize")] + ["-o", lib] + [j[j.index("-o") + 1] for j in jobs] + ["-lpthread"])
    saved = (api.LIB_PATH, api._lib, os.environ.get("MOF_SMOOTH_AHEAD"))
    api.LIB_PATH, api._lib = lib, None
    os.environ["MOF_SMOOTH_AHEAD"] = "0"  # the second stream's worker thread would share the emulator's thread/block registers
    try:
        yield api
    finally:
        api.LIB_PATH, api._lib = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("MOF_SMOOTH_AHEAD", None)
        else:
            os.environ["MOF_SMOOTH_AHEAD"] = saved[2]


@pytest.fixture()
def aligner(emulated):
    al = emulated.Aligner(0)
    yield al
    al.close()


def test_the_emulated_build_is_the_whole_c_abi(emulated):
    lib = emulated.load_library()
    assert lib._name.endswith("libmof_emul.so")
    for name in emulated.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name


def test_vertex_alignment_matches_the_reference_golden(aligner, golden_sphere):
    """258 vertices: too small for a hierarchy, so every solve is the persistent Jacobi-PCG kernel (pcg_kernels.cu)."""
    g = golden_sphere
    v, t = g["input_vertices_f32"].astype(np.float64), g["triangles"]
    al = aligner
    al.set_mesh(v, t)
    assert np.array_equal(al.array(api.ARR_OPPOSITE), g["oppositeEdge"]) and np.array_equal(al.array(api.ARR_REDUCED_EDGE), g["reducedEdgeIndex"])
    E = al.num_edges
    S, ref = al.csr(api.CSR_WHITNEY_SMOOTH), csr_from_golden(g, "smoothOperator", (E, E))
    assert np.array_equal(S.indptr, ref.indptr) and np.array_equal(S.indices, ref.indices) and rel(S.data, ref.data) < 1e-11
    al.set_signals(g["input_a"].astype(np.float64), g["input_b"].astype(np.float64))
    for i in range(10):
        al.iterate(1)
        assert rel(al.flow(), g["it%02d.tFlowField" % i]) < 1e-6, i  # north_star gate: 1e-3
    ca, cb = al.advect_vertices(0.5)
    assert np.abs(ca - g["advected0"]).max() < 1e-4 and np.abs(cb - g["advected1"]).max() < 1e-4
    out = O.to_uchar_ply((ca + cb) / 2.0)
    assert np.abs(out.astype(int) - g["output_rgb"].astype(int)).max() <= 1
    s = al.stats()
    assert s["flowSolves"] == 10 and s["lastFlowResidual"] <= 1.01e-8 and s["kernelLaunches"] > 0


def test_renumbered_mesh_gives_the_callers_numbering_back(aligner, golden_sphere, emulated):
    """mof_set_reorder (reorder.cu): the golden sphere with its vertices and triangles shuffled, renumbering forced — the flow and the
    advected colours come back in the caller's (shuffled) numbering and are the golden's; the texture entries refuse the mesh."""
    g = golden_sphere
    v, t = g["input_vertices_f32"].astype(np.float64), g["triangles"]
    rng = np.random.default_rng(5)
    vo, to = rng.permutation(v.shape[0]), rng.permutation(t.shape[0])  # shuffled index -> original index
    rank = np.empty_like(vo)
    rank[vo] = np.arange(vo.size)
    vs, ts = np.ascontiguousarray(v[vo]), np.ascontiguousarray(rank[t][to].astype(np.int32))
    al = aligner
    al.set_reorder(1)
    al.set_mesh(vs, ts)
    on, vorder, torder = al.permutation()
    assert on and sorted(vorder.tolist()) == list(range(v.shape[0])) and sorted(torder.tolist()) == list(range(t.shape[0]))
    al.set_signals(g["input_a"].astype(np.float64)[vo], g["input_b"].astype(np.float64)[vo])
    for i in range(10):
        al.iterate(1)
        f = np.empty_like(g["it%02d.tFlowField" % i])
        f[to] = al.flow()
        assert rel(f, g["it%02d.tFlowField" % i]) < 1e-6, i
    ca, cb = al.advect_vertices(0.5)
    a0, b0 = np.empty_like(ca), np.empty_like(cb)
    a0[vo], b0[vo] = ca, cb
    assert np.abs(a0 - g["advected0"]).max() < 1e-4 and np.abs(b0 - g["advected1"]).max() < 1e-4
    with pytest.raises(api.MofError, match="mof_set_reorder"):
        al.set_texture_map(4, 4, np.zeros(16, np.int32), np.zeros((16, 2)), np.zeros((t.shape[0], 6)), np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4, 3), np.uint8))
    al.set_reorder(-1)
    al.set_mesh(vs, ts)  # small mesh: left alone
    assert not al.permutation()[0]
    # the decision itself, on a context whose first mesh this is: shuffled -> renumbered, the file's own order -> left alone
    os.environ["MOF_REORDER_MIN_VERTICES"] = "100"
    try:
        fresh = emulated.Aligner(0)
        fresh.set_mesh(vs, ts)
        assert fresh.permutation()[0]
        v5, t5 = synthetic.octahedron_sphere(5)  # numbered along a Morton curve by the generator (mean index span of a triangle: V/30)
        fresh.set_mesh(v5, t5)
        assert not fresh.permutation()[0]
        v5, t5 = synthetic.octahedron_sphere(5, spatial_sort=False)  # ... in the order the subdivision created the vertices (V/2)
        fresh.set_mesh(v5, t5)
        assert fresh.permutation()[0]
        fresh.close()
    finally:
        del os.environ["MOF_REORDER_MIN_VERTICES"]


@pytest.mark.parametrize("mode,cmode,count,level", [(0, 0, 6, 3), (2, 0, 6, 3), (2, 2, 6, 3), (1, 0, 6, 3), (1, 0, 6, 5)])
def test_spectrum_matches_the_shift_invert_lanczos_of_the_checker(aligner, mode, cmode, count, level):
    """mof_spectrum (csrc/spectrum.cu: LOBPCG) against ComputeSpectrum as the checker restates it (ARPACK shift-invert through scipy):
    eigenvalues, and the span of the prolonged eigenvectors cluster by cluster (a sphere's eigenvalues are multiple). Level 5 (4 098
    vertices) has a scalar hierarchy: the Conformal basis then runs with its two-cycle preconditioner instead of the inverse diagonal."""
    v, t = synthetic.octahedron_sphere(level)
    al = aligner
    p = api.default_params()
    p.vfMode, p.cMode = mode, cmode
    al.set_params(p)
    al.set_mesh(v, t)
    ev, fields, its, res = al.spectrum(count, 1e-9, 3000)
    ref_ev, ref_fields, _, _, _ = O.spectrum(v, t, count, mode, cmode)  # (Conformal: on the quotient by the constants of either potential)
    assert res <= 1e-9 and its > 0
    assert np.abs(ev - ref_ev).max() <= 1e-7 * np.abs(ref_ev).max(), (ev, ref_ev)
    # <f, h> = sum_t f_t^T (g_t area_t) h_t = x^T M y: both sets are orthonormal in it; on whole clusters the cross-Gram matrix is orthogonal
    g = O.make_unit_area(O.metric_from_embedding(v, t))
    area = O.triangle_areas(g)

    def weighted(f):
        return np.stack([(g[:, 0] * f[:, 0] + g[:, 1] * f[:, 1]) * area, (g[:, 1] * f[:, 0] + g[:, 2] * f[:, 1]) * area], 1)

    G = np.array([[np.sum(fields[i] * weighted(ref_fields[j])) for j in range(count)] for i in range(count)])
    assert np.abs(G @ G.T - np.eye(count)).max() < 1e-5, G


def test_spectrum_on_a_torus_finds_the_harmonic_fields(aligner, golden_torus):
    """Genus 1: S is singular (two harmonic fields, lambda = 0 — the reference's shift of 1e-8 is what lets it factorise). The block
    iteration returns them first, then the pairs of the torus' symmetric spectrum; eigenvalues against the checker's ARPACK values
    (absolutely, on the scale of the largest: the zeros are zeros to rounding on both sides)."""
    g = golden_torus
    v, t = g["vertices"].astype(np.float64), g["triangles"].astype(np.int32)
    al = aligner
    al.set_mesh(v, t)
    ev, fields, its, res = al.spectrum(6, 1e-8, 3000)
    ref_ev = O.spectrum(v, t, 6, 0, 0)[0]
    assert res <= 1e-8 and np.abs(ev - ref_ev).max() <= 1e-7 * ref_ev.max(), (ev, ref_ev)
    assert np.abs(ev[:2]).max() < 1e-9 and ev[2] > 1.0


def test_multigrid_and_jacobi_solves_agree_with_the_checker(emulated):
    """4 098 vertices: three-level hierarchies for the flow and the smoothing systems (sliced fine sweeps, 27-point stencil level,
    dense coarsest solve, two iterations per replayed graph); then the same alignment with the Jacobi-PCG kernel."""
    v, t = synthetic.octahedron_sphere(5)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 1))
    st, _ = O.align_vertices(v, t, a, b, O.Params(iterations=2))
    flows, iters = {}, {}
    saved = {k: os.environ.get(k) for k in ("MOF_FLOW_MG", "MOF_SCALAR_MG")}
    try:
        for mode in ("1", "0"):
            os.environ["MOF_FLOW_MG"] = os.environ["MOF_SCALAR_MG"] = mode  # read when the mesh is set
            al = emulated.Aligner(0)
            try:
                al.set_mesh(v, t)
                al.set_signals(a, b)
                al.iterate(2)
                s = al.stats()
                assert s["lastFlowResidual"] <= 1.01e-8 and s["lastSmoothResidual"] <= 1.01e-10
                flows[mode], iters[mode] = al.flow(), (s["flowCgIterations"], s["smoothCgIterations"])
                if mode == "1":  # the assembled flow system of the last iteration: symmetric, on the pattern of the smoothness operator
                    A, S = al.csr(api.CSR_FLOW_SYSTEM), al.csr(api.CSR_WHITNEY_SMOOTH)
                    assert np.array_equal(A.indptr, S.indptr) and np.array_equal(A.indices, S.indices) and abs(A - A.T).max() < 1e-12 * abs(A).max()
            finally:
                al.close()
    finally:
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    assert rel(flows["1"], st.tfield) < 1e-6 and rel(flows["0"], st.tfield) < 1e-6
    assert rel(flows["1"], flows["0"]) < 1e-6
    assert iters["1"][0] < iters["0"][0] / 3 and iters["1"][1] < iters["0"][1]  # the hierarchies are in use


def test_solver_variants_give_the_same_bits(emulated, monkeypatch):
    """The PCG loop as one graph with a device-side WHILE node (default) against host-driven replays of two iterations
    (MOF_MG_WHILE=0), and the small multigrid levels as one kernel on one thread-block cluster (MOF_MG_TAIL_CELLS=6144: operation
    program, staging, slot-ordered sums, fused correction) against the stand-alone kernels: same flow, bit for bit, same counts."""
    v, t = synthetic.octahedron_sphere(5)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 4))
    runs = {}
    for name, env in (("default", {}), ("replay", {"MOF_MG_WHILE": "0"}), ("cluster", {"MOF_MG_TAIL_CELLS": "6144"}),
                      ("cluster_streamed", {"MOF_MG_TAIL_CELLS": "6144", "MOF_MG_TAIL_RESIDENT": "100"})):
        for k in ("MOF_MG_WHILE", "MOF_MG_TAIL_CELLS", "MOF_MG_TAIL_RESIDENT"):
            monkeypatch.delenv(k, raising=False)
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        al = emulated.Aligner(0)
        try:
            al.set_mesh(v, t)
            al.set_signals(a, b)
            al.iterate(2)
            s = al.stats()
            runs[name] = (al.flow(), al.array(api.ARR_SMOOTHED), s["flowCgIterations"], s["smoothCgIterations"], s["kernelLaunches"])
        finally:
            al.close()
    for name in ("replay", "cluster", "cluster_streamed"):
        assert np.array_equal(runs[name][0], runs["default"][0]) and np.array_equal(runs[name][1], runs["default"][1]), name
        assert runs[name][2:4] == runs["default"][2:4], name
    assert runs["cluster"][4] < runs["default"][4]  # fewer launches: the cluster kernel was in use


@pytest.mark.parametrize("name", sorted(VF_MODES))
def test_conformal_and_connection_bases_match_the_reference_golden(aligner, golden_modes, name):
    g = golden_modes
    vf_mode, c_mode = VF_MODES[name]
    v, t = g["input_vertices_f32"].astype(np.float64), g["triangles"]
    p = api.default_params()
    p.vfMode, p.cMode, p.vfSmooth, p.iterations = vf_mode, c_mode, (3e-6, 5e-7, 1e4)[vf_mode], 4
    al = aligner
    al.set_params(p)
    al.set_mesh(v, t)
    al.set_signals(g["input_a"].astype(np.float64), g["input_b"].astype(np.float64))
    for i in range(4):
        al.iterate(1)
        assert rel(al.flow(), g["%s.it%02d.tFlowField" % (name, i)]) < 1e-5, (name, i)


def test_conformal_basis_with_the_two_cycle_preconditioner(aligner):
    """1 026 vertices: the scalar hierarchy exists, so the Conformal solve is preconditioned by two of its cycles."""
    v, t = synthetic.octahedron_sphere(4)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 2))
    p = api.default_params()
    p.vfMode, p.vfSmooth, p.iterations = 1, 5e-7, 2
    st, ref = O.align_vertices(v, t, a, b, O.Params(iterations=2, vfMode=1))
    al = aligner
    al.set_params(p)
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.iterate(2)
    assert rel(al.flow(), st.tfield) < 1e-4
    ca, cb = al.advect_vertices(0.5)
    assert np.abs((ca + cb) / 2.0 - ref).max() < 1e-2


def test_six_channel_blend(aligner):
    v, t = synthetic.octahedron_sphere(4)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 3))
    p = api.default_params()
    p.dogWeight, p.iterations = 0.5, 2
    st, ref = O.align_vertices(v, t, a, b, O.Params(iterations=2, dogWeight=0.5))
    al = aligner
    al.set_params(p)
    al.set_mesh(v, t)
    al.set_signals(a, b)
    al.iterate(2)
    assert rel(al.flow(), st.tfield) < 1e-6
    ca, cb = al.advect_vertices(0.5)
    assert np.abs((ca + cb) / 2.0 - ref).max() < 1e-4


def test_texture_configuration_end_to_end(aligner, golden_torus):
    """--mesh m.ply --in A.png B.png as the command line drives it with MOF_GPU_TEXPREP=1: subdivision, mesh, texel map and vertex
    colours on the "device", ten iterations, texel advection — against the reference's golden fixture."""
    g = golden_torus
    ta, tb = g["input_tex_a"], g["input_tex_b"]
    v0 = g["input_vertices_f32"]
    lo, hi = v0.astype(np.float64).min(0), v0.astype(np.float64).max(0)
    e_len = float(np.float32(np.float32(0.08) * float(np.sqrt(((hi - lo) ** 2).sum()))))
    al = aligner
    v, t, uv = al.subdivide(v0, g["input_triangles"], g["input_uv"].astype(np.float64), e_len)
    assert np.array_equal(v.astype(np.float64), g["vertices"]) and np.array_equal(t, g["triangles"])
    al.set_mesh(v.astype(np.float64), t)
    srcT, srcP = al.build_texture_map(48, 48, 2, uv, ta, tb)
    assert np.array_equal(srcT, g["textureSource_tIdx"])
    ca, cb = al.sample_textures_to_vertices()
    al.set_signals(ca, cb)
    s6 = al.array(api.ARR_SIGNALS)
    assert rel(s6[:, :3], g["signals0"]) < 1e-7 and rel(s6[:, 3:], g["signals1"]) < 1e-7
    for i in range(10):
        al.iterate(1)
        assert rel(al.flow(), g["it%02d.tFlowField" % i]) < 1e-5, i
    oa, ob = al.advect_texels(0.5)
    assert colour_outliers(oa, g["advected0"], 1.0) < 2e-3 and colour_outliers(ob, g["advected1"], 1.0) < 2e-3
    pixels = O.to_uchar_png((oa + ob) / 2.0).reshape(48, 48, 3)[::-1]
    assert colour_outliers(pixels, g["output_pixels"], 1.0) < 2e-3


def test_stand_alone_solver_and_error_paths(aligner):
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(5)
    n = 1203  # not a multiple of the slice height
    rows, cols, w = rng.integers(0, n, 6 * n), rng.integers(0, n, 6 * n), rng.uniform(0.1, 1.0, 6 * n)
    W = sp.coo_matrix((w, (rows, cols)), shape=(n, n)).tocsr()
    W = W + W.T
    A = (sp.diags(np.asarray(W.sum(1)).ravel() + rng.uniform(0.01, 0.1, n)) - W).tocsr()
    b = rng.standard_normal(n)
    x, iters, relres = aligner.pcg_solve_csr(A, b, 1e-10)
    assert relres <= 1.01e-10 and iters > 0 and rel(x, spla.spsolve(A.tocsc(), b)) < 1e-7
    al = aligner
    with pytest.raises(api.MofError) as e:  # half-edge 0->1 used twice (FEM.inl:599)
        al.set_mesh(np.eye(4, 3), np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32))
    assert e.value.code == api.MOF_E_MESH and "Edge is occupied" in e.value.message
    with pytest.raises(api.MofError) as e:
        al.iterate(1)
    assert e.value.code == api.MOF_E_INVALID
    # sizes are checked before anything is read: an empty mesh, and one whose matrix indices would not fit 32 bits
    lib, one, tri = al._lib, np.zeros(9), np.zeros(3, dtype=np.int32)
    ptr_d, ptr_i = one.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), tri.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    assert lib.mof_set_mesh(al._ctx, ptr_d, 0, ptr_i, 0) == api.MOF_E_INVALID and b"empty mesh" in lib.mof_last_error(al._ctx)
    assert lib.mof_set_mesh(al._ctx, ptr_d, 70000000, ptr_i, 140000000) == api.MOF_E_INVALID and b"32-bit" in lib.mof_last_error(al._ctx)


# ------------------------------------------------------------------ the GPU tier's own assertions, on the emulated build
#
# The functions below are the GPU tier's tests, imported and called with THIS module's aligner (the `gpu` marker of their
# modules applies to collection there, not to a call from here): what will be asserted on a B200 is asserted here first.

def test_gpu_tier_every_stage_matches_the_oracle(aligner):
    import test_gpu_parity
    test_gpu_parity.test_every_stage_matches_the_oracle(aligner)


def test_gpu_tier_error_paths_and_symmetries(aligner):
    import test_gpu_parity
    test_gpu_parity.test_error_paths(aligner)
    test_gpu_parity.test_identical_signals_give_zero_flow_and_swapping_flips_it(aligner)


def test_gpu_tier_texel_frame_sequence(aligner, golden_torus):
    import test_gpu_parity
    test_gpu_parity.test_texel_frame_sequence_matches_the_oracle(aligner, golden_torus)


def test_gpu_tier_log_space_comparison(aligner, golden_modes):
    import test_gpu_modes
    test_gpu_modes.test_log_space_comparison(aligner, golden_modes)


def test_gpu_tier_switching_bases_on_one_context(aligner):
    import test_gpu_modes
    test_gpu_modes.test_switching_bases_on_one_context(aligner)


@pytest.mark.parametrize("case", ["seams_pad3_bilinear", "large_triangles_nearest", "no_padding", "vertices_on_texels"])
def test_gpu_tier_texel_map_matches_the_checker(aligner, case):
    import test_gpu_texprep
    test_gpu_texprep.test_texel_map_matches_the_checker(aligner, case)


def test_gpu_tier_texture_preparation(aligner, golden_torus):
    import test_gpu_texprep
    test_gpu_texprep.test_subdivision_matches_the_reference_golden(aligner, golden_torus)
    test_gpu_texprep.test_subdivision_matches_the_checker(aligner, 9, 5, 0.05)
    test_gpu_texprep.test_error_paths(aligner)
    test_gpu_texprep.test_texel_map_and_vertex_colours_match_the_reference_golden(aligner, golden_torus)


def test_command_line_on_the_emulated_build(emulated, tmp_path, golden_torus):
    """The drop-in command line (csrc/host/*.cpp) linked against the emulated library instead of libmof_b200.so: the texture
    configuration from files to file, with the preparation on the host (default) and on the "device" (MOF_GPU_TEXPREP=1)."""
    from PIL import Image
    host = os.path.join(ROOT, "meshopticalflow_b200", "csrc", "host")
    lib = emulated.load_library()._name
    exe = str(tmp_path / "OpticalFlow_emul")
    sources = [os.path.join(host, f) for f in ("optical_flow_main.cpp", "ply_io.cpp", "png_codec.cpp", "texture_prep.cpp", "cmdline.cpp")]
    # -DMOF_WITH_HOST_TEXPREP: the serial host restatement of the preparation (host/texture_prep.cpp, test infrastructure) as the cross-check
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DMOF_WITH_HOST_TEXPREP", "-I" + os.path.join(ROOT, "include"), "-o", exe] + sources + [lib, "-lz", "-Wl,-rpath," + os.path.dirname(lib)])
    g = golden_torus
    synthetic.write_ply_textured(str(tmp_path / "m.ply"), g["input_vertices_f32"], g["input_triangles"], g["input_uv"])
    open(tmp_path / "A.png", "wb").write(g["png_a"].tobytes())
    open(tmp_path / "B.png", "wb").write(g["png_b"].tobytes())
    pictures = []
    for mode in ("0", "1"):
        r = subprocess.run([exe, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "r%s.png" % mode, "--eLength", "0.08"], cwd=tmp_path, capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, MOF_GPU_TEXPREP=mode, MOF_SMOOTH_AHEAD="0"))
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Num vertices %d" % g["vertices"].shape[0] in r.stdout  # OpticalFlow.cpp:716
        pictures.append(np.asarray(Image.open(tmp_path / ("r%s.png" % mode))).astype(int))
    assert pictures[0].shape == (48, 48, 3) and np.array_equal(pictures[0], pictures[1])  # same preparation, bit for bit => same file
    assert colour_outliers(pictures[1], g["output_pixels"], 1.0) < 2e-3


def test_spectrum_command_line_on_the_emulated_build(emulated, tmp_path):
    """The Spectrum command line (csrc/host/spectrum_main.cpp) linked against the emulated library: eigenvector-%03d.bin files in the
    working directory (int count, count x 2 doubles: WriteVector, Src/VectorIO.h:23-31), the reference's eigenvalue listing, the usage
    text without --mesh."""
    import struct
    host = os.path.join(ROOT, "meshopticalflow_b200", "csrc", "host")
    lib = emulated.load_library()._name
    exe = str(tmp_path / "Spectrum_emul")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe, os.path.join(host, "spectrum_main.cpp"), os.path.join(host, "ply_io.cpp"),
                           lib, "-Wl,-rpath," + os.path.dirname(lib)])
    v, t = synthetic.octahedron_sphere(3)
    synthetic.write_ply_colored(str(tmp_path / "m.ply"), v, np.zeros((v.shape[0], 3), np.uint8), t)
    r = subprocess.run([exe, "--mesh", "m.ply", "--eigenVectors", "6", "--vfMode", "2", "--cMode", "1"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    vf = v.astype(np.float32).astype(np.float64)  # what the tool read from the file
    ref_ev = O.spectrum(vf, t, 6, 2, 1)[0]
    listed = np.array([float(x) for x in r.stdout.split("Eigenvalues:")[1].split()])
    assert np.abs(listed - ref_ev).max() < 2e-8 * ref_ev.max() + 1e-8
    for i in range(6):
        raw = (tmp_path / ("eigenvector-%03d.bin" % (i + 1))).read_bytes()
        assert struct.unpack("<i", raw[:4])[0] == t.shape[0] and len(raw) == 4 + 16 * t.shape[0]
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "--mesh" in r.stdout
    r = subprocess.run([exe, "--mesh", "m.ply", "--edgeMetric"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "edgeMetric" in r.stderr
