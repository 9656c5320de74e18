"""CPU tier: one mesh over several "GPUs" (include/mof_b200.h, mof_dist_*; SURVEY.md §8e, BASELINE.json configs[4]) without a GPU:
the emulated build of the whole library (tests/host_emulation, see test_library_host_emulation.py) INCLUDING csrc/dist.cu, with
every rank on its own OS thread of this process and NCCL replaced by an in-process stand-in (tests/host_emulation/nccl.h:
grouped send / receive, all-reduce in rank order, all-gather, broadcast — also inside the captured PCG iteration). The
row-partitioned flow and smoothing solves (blocks of 32-row slices / vertex rows, halo index lists from the sparsity patterns,
packed halo exchange, all-reduced dot products, every rank's rows gathered at the end) must give the single-"GPU" result on
every rank, for worlds of 1, 2, 3 and 4."""
import os
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import ROOT, rel  # noqa: F401 (golden_sphere comes from conftest)
from meshopticalflow_b200 import api, synthetic

EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
UNITS = 10  # library_emul.cpp: units 0-6, 8 and 9 as in test_library_host_emulation.py, unit 7 = dist.cu


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    out = tmp_path_factory.mktemp("dist_emul")
    extra = os.environ.get("MOF_EMUL_CXXFLAGS", "-O2").split()
    base = ["g++"] + extra + ["-std=c++17", "-fPIC", "-c", "-x", "c++", "-DMOF_HOST_EMULATION", "-fno-gnu-unique", "-DMOF_EMUL_THREADS", "-I.", "-w"]
    jobs = [base + ["-DEMUL_UNIT=%d" % u, "-o", str(out / ("unit%d.o" % u)), "library_emul.cpp"] for u in range(UNITS)]
    jobs += [base + ["-o", str(out / "runtime.o"), "emul_runtime.cpp"]]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda cmd: subprocess.check_call(cmd, cwd=EMU_DIR), jobs))
    lib = str(out / "libmof_emul_threads.so")
    subprocess.check_call(["g++", "-shared"] + [f for f in extra if f.startswith("-fsanitize")] + ["-o", lib] + [j[j.index("-o") + 1] for j in jobs] + ["-lpthread"])
    saved = (api.LIB_PATH, api._lib, os.environ.get("MOF_SMOOTH_AHEAD"))
    api.LIB_PATH, api._lib = lib, None
    os.environ["MOF_SMOOTH_AHEAD"] = "0"
    try:
        api.load_library()
        yield api
    finally:
        api.LIB_PATH, api._lib = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("MOF_SMOOTH_AHEAD", None)
        else:
            os.environ["MOF_SMOOTH_AHEAD"] = saved[2]


def _align(emulated, v, t, a, b, iterations, world=0, rank=0, uid=None, out=None):
    """One rank's alignment (world = 0: no communicator at all). ctypes releases the GIL inside every library call, so the
    ranks really run side by side and meet in the collectives."""
    try:
        al = emulated.Aligner(0)
        try:
            if world:
                al.dist_init(world, rank, uid)
            al.set_mesh(v, t)
            al.set_signals(a, b)
            al.iterate(iterations)
            flow = al.flow()
            ca, cb = al.advect_vertices(0.5)
            result = {"flow": flow, "colours": (ca + cb) / 2.0, "stats": al.stats()}
        finally:
            al.close()
    except BaseException as e:  # a rank that dies would leave the others waiting in a collective
        result = {"error": repr(e)}
    if out is not None:
        out[rank] = result
    return result


def _run_world(emulated, world, v, t, a, b, iterations):
    uid = emulated.dist_unique_id()
    out = [None] * world
    threads = [threading.Thread(target=_align, args=(emulated, v, t, a, b, iterations, world, r, uid, out)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=600)
    assert all(not th.is_alive() for th in threads), "a rank is stuck in a collective"
    for r, res in enumerate(out):
        assert res is not None and "error" not in res, (r, res)
    return out


@pytest.fixture(scope="module")
def workload(emulated):
    v, t = synthetic.octahedron_sphere(5)  # 4 098 vertices, 12 288 Whitney unknowns = 384 slices; three-level hierarchies
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 4))
    single = _align(emulated, v, t, a, b, 2)
    assert "error" not in single, single
    return v, t, a, b, single


@pytest.mark.parametrize("world,exchange", [(1, "nccl"), (2, "nccl"), (3, "nccl"), (4, "nccl"), (2, "peer"), (3, "peer"), (4, "peer")])
def test_partitioned_solves_give_the_single_gpu_result_on_every_rank(emulated, workload, world, exchange, monkeypatch):
    """exchange = "peer": halo values written straight into the peers' memory windows by one fused kernel per exchange (dist.cu:
    PeerWindow, MOF_DIST_P2P=1; here the windows are plain pointers between the ranks' OS threads) instead of NCCL send / recv."""
    monkeypatch.setenv("MOF_DIST_P2P", "1" if exchange == "peer" else "0")
    v, t, a, b, single = workload
    out = _run_world(emulated, world, v, t, a, b, 2)
    for r, res in enumerate(out):
        assert rel(res["flow"], single["flow"]) < 1e-6, (world, r)
        assert np.abs(res["colours"] - single["colours"]).max() < 1e-3, (world, r)
        assert res["stats"]["lastFlowResidual"] <= 1.01e-8 and res["stats"]["lastSmoothResidual"] <= 1.01e-10
        assert np.array_equal(res["flow"], out[0]["flow"]) and np.array_equal(res["colours"], out[0]["colours"])  # the ranks agree bit for bit
        assert (res["stats"]["haloEntries"] > 0) == (world > 1)
    # the preconditioner is the same operator however the rows are split (the sums are taken in another order, no more):
    its = [res["stats"]["flowCgIterations"] for res in out]
    assert max(its) == min(its) and abs(its[0] - single["stats"]["flowCgIterations"]) <= 4


@pytest.mark.parametrize("world,threshold", [(2, 100), (3, 100), (4, 800)])
def test_coarse_levels_dealt_to_the_ranks(emulated, world, threshold, monkeypatch):
    """16 386 vertices, four-level hierarchies. MOF_DIST_LEVEL_CELLS lowered so that the levels of more than `threshold` cells are
    dealt to the ranks in octree-aligned cell ranges (threshold 100: two levels, 800: one): stencil halos per level, the
    aggregates that straddle a row-block boundary (residual rows in, correction cells out), restriction and prolongation inside a
    rank, the first replicated level gathered once per visit — against the single-"GPU" run and across ranks."""
    monkeypatch.setenv("MOF_DIST_P2P", "1")  # the levels' halos and the gather of the first replicated level through the peer windows too
    v, t = synthetic.octahedron_sphere(6)
    a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 9))
    single = _align(emulated, v, t, a, b, 1)
    assert "error" not in single, single
    monkeypatch.setenv("MOF_DIST_LEVEL_CELLS", str(threshold))
    monkeypatch.setenv("MOF_DIST_LEVEL_CELLS_SCALAR", str(threshold))  # (replicated by default)
    out = _run_world(emulated, world, v, t, a, b, 1)
    for r, res in enumerate(out):
        assert rel(res["flow"], single["flow"]) < 1e-6, (world, r)
        assert np.abs(res["colours"] - single["colours"]).max() < 1e-3, (world, r)
        assert res["stats"]["lastFlowResidual"] <= 1.01e-8 and res["stats"]["lastSmoothResidual"] <= 1.01e-10
        assert np.array_equal(res["flow"], out[0]["flow"])
    its = [res["stats"]["flowCgIterations"] for res in out]
    assert max(its) == min(its) and abs(its[0] - single["stats"]["flowCgIterations"]) <= 4
    # fewer launches per rank than with replicated coarse levels would not show here (same kernels on fewer cells); what shows is
    # that the level-1 all-reduce is gone: the exchanged halo of the flow system's level partitions is not empty
    monkeypatch.setenv("MOF_DIST_LEVEL_CELLS", "100000000")
    monkeypatch.setenv("MOF_DIST_LEVEL_CELLS_SCALAR", "0")
    replicated = _run_world(emulated, world, v, t, a, b, 1)
    assert rel(replicated[0]["flow"], out[0]["flow"]) < 1e-9
    assert replicated[0]["stats"]["flowCgIterations"] == its[0]


def test_repeatable_and_basis_restriction(emulated, workload):
    v, t, a, b, _ = workload
    first = _run_world(emulated, 2, v, t, a, b, 1)
    again = _run_world(emulated, 2, v, t, a, b, 1)
    assert np.array_equal(first[0]["flow"], again[0]["flow"])  # same inputs, same bits
    # a partitioned mesh supports the Whitney basis only (mof_api.cu: finish_signals)
    uid = emulated.dist_unique_id()
    errors = [None, None]

    def rank(r):
        al = emulated.Aligner(0)
        try:
            p = emulated.default_params()
            p.vfMode, p.vfSmooth = 2, 1e4
            al.set_params(p)
            al.dist_init(2, r, uid)
            al.set_mesh(v, t)
            try:
                al.set_signals(a, b)
            except emulated.MofError as e:
                errors[r] = e
        finally:
            al.close()

    threads = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert all(e is not None and e.code == api.MOF_E_UNSUPPORTED for e in errors)


def test_cooperative_jacobi_pcg_on_three_ctas(emulated, golden_sphere):
    """This build runs a cooperative kernel with one OS thread per CTA (three of them, meeting in grid.sync()): the persistent
    Jacobi-PCG kernel of pcg_kernels.cu — every solve of a mesh too small for a hierarchy — with its per-CTA partial sums in
    three rotating banks, against the reference's golden flow."""
    g = golden_sphere
    al = emulated.Aligner(0)
    try:
        al.set_mesh(g["input_vertices_f32"].astype(np.float64), g["triangles"])
        al.set_signals(g["input_a"].astype(np.float64), g["input_b"].astype(np.float64))
        for i in range(4):
            al.iterate(1)
            assert rel(al.flow(), g["it%02d.tFlowField" % i]) < 1e-6, i
        s = al.stats()
        assert s["lastFlowResidual"] <= 1.01e-8 and s["lastSmoothResidual"] <= 1.01e-10
    finally:
        al.close()


def test_partitioned_solves_on_a_renumbered_mesh(emulated, workload, monkeypatch):
    """A badly numbered mesh is renumbered along a Morton curve inside mof_set_mesh (reorder.cu) BEFORE the rows are dealt to the ranks —
    contiguous ranges of the library's numbering are compact patches whatever the caller's was — and every rank hands the flow and the
    colours back in the caller's numbering: a shuffled copy of the workload on two ranks against the sorted single-"GPU" run."""
    v, t, a, b, single = workload
    rng = np.random.default_rng(2)
    vo, to = rng.permutation(v.shape[0]), rng.permutation(t.shape[0])  # shuffled index -> original index
    rank = np.empty_like(vo)
    rank[vo] = np.arange(vo.size)
    vs, ts = np.ascontiguousarray(v[vo]), np.ascontiguousarray(rank[t][to].astype(np.int32))
    monkeypatch.setenv("MOF_REORDER", "1")
    monkeypatch.setenv("MOF_DIST_P2P", "0")
    out = _run_world(emulated, 2, vs, ts, a[vo], b[vo], 2)
    for r, res in enumerate(out):
        flow = np.empty_like(res["flow"])
        flow[to] = res["flow"]
        colours = np.empty_like(res["colours"])
        colours[vo] = res["colours"]
        assert rel(flow, single["flow"]) < 1e-6, r
        assert np.abs(colours - single["colours"]).max() < 1e-3, r
        assert res["stats"]["haloEntries"] > 0
        assert np.array_equal(res["flow"], out[0]["flow"])
