"""ctypes binding of libmof_b200.so (include/mof_b200.h) for the test and bench harness.

This is plumbing over the C ABI, not a second implementation: every method is one C call. There is
no CPU path — importing works anywhere (so the CPU test tier can check the exported symbols), but
constructing an `Aligner` without the library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_ubyte, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmof_b200.so")

MOF_OK, MOF_E_INVALID, MOF_E_CUDA, MOF_E_MESH, MOF_E_NOCONVERGE, MOF_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

CSR_SCALAR_MASS, CSR_SCALAR_STIFFNESS, CSR_WHITNEY_SMOOTH, CSR_FLOW_SYSTEM = range(4)
(ARR_METRIC, ARR_AREA, ARR_OPPOSITE, ARR_XFORM_LINEAR, ARR_XFORM_CONSTANT, ARR_REDUCED_EDGE, ARR_EXPANDED_EDGE, ARR_POSITIVE_EDGE,
 ARR_PROLONGATION, ARR_SIGNALS, ARR_SMOOTHED, ARR_RESAMPLED, ARR_DATA_TERM, ARR_DATA_RHS, ARR_FLOW_RHS, ARR_FLOW_SOLUTION, ARR_SIGNALS_RAW,
 ARR_RESAMPLED_RAW) = range(18)

VF_WHITNEY, VF_CONFORMAL, VF_CONNECTION = range(3)  # --vfMode, VectorField.h:3-7

_ARRAY_SPEC = {
    ARR_METRIC: (np.float64, 3), ARR_AREA: (np.float64, None), ARR_OPPOSITE: (np.int32, None), ARR_XFORM_LINEAR: (np.float64, 4),
    ARR_XFORM_CONSTANT: (np.float64, 2), ARR_REDUCED_EDGE: (np.int32, None), ARR_EXPANDED_EDGE: (np.int32, None), ARR_POSITIVE_EDGE: (np.int32, None),
    ARR_PROLONGATION: (np.float64, 6), ARR_SIGNALS: (np.float64, 6), ARR_SMOOTHED: (np.float64, 6), ARR_RESAMPLED: (np.float64, 6),
    ARR_DATA_TERM: (np.float64, 3), ARR_DATA_RHS: (np.float64, 2), ARR_FLOW_RHS: (np.float64, None), ARR_FLOW_SOLUTION: (np.float64, None),
    ARR_SIGNALS_RAW: (np.float64, 6), ARR_RESAMPLED_RAW: (np.float64, 6),
}

# Every symbol include/mof_b200.h declares (the CPU test tier checks the library exports them all).
# mof_kernel_id (include/mof_b200.h): kernels mof_time_kernel can time on their own
KERNELS = {"flow_spmv": 0, "flow_fine_sweep": 1, "flow_update": 2, "flow_restrict": 3, "flow_prolong": 4, "flow_direction": 5, "flow_level1": 6,
           "scalar_spmv": 7, "scalar_fine_sweep": 8, "scalar_update": 9, "scalar_level1": 10, "walk": 11}

EXPORTED_SYMBOLS = [
    "mof_default_params", "mof_create", "mof_destroy", "mof_last_error", "mof_set_params", "mof_get_stats", "mof_reset_stats", "mof_synchronize",
    "mof_set_mesh", "mof_set_mesh_device", "mof_set_reorder", "mof_get_permutation", "mof_spectrum", "mof_set_signals", "mof_set_signals_device", "mof_iterate", "mof_get_flow", "mof_get_coeffs", "mof_num_edges", "mof_num_coeffs",
    "mof_advect_vertices", "mof_advect_vertices_device", "mof_set_texture_map", "mof_advect_texels", "mof_advect_texels_frames", "mof_csr_size", "mof_get_csr", "mof_array_bytes",
    "mof_get_array", "mof_pcg_solve_csr", "mof_time_flow_spmv", "mof_time_kernel", "mof_dist_unique_id", "mof_dist_init",
    "mof_subdivide", "mof_get_subdivision", "mof_build_texture_map", "mof_get_texture_map", "mof_sample_textures_to_vertices",
]


class Params(ctypes.Structure):
    _fields_ = [("iterations", c_int), ("sSmooth", c_double), ("sMultiply", c_double), ("vfSmooth", c_double), ("vMultiply", c_double),
                ("vfSThreshold", c_double), ("dogWeight", c_double), ("dogSmooth", c_double), ("flowTol", c_double), ("smoothTol", c_double),
                ("maxCgIterations", c_int), ("vfMode", c_int), ("cMode", c_int), ("logSpace", c_int)]


class Stats(ctypes.Structure):
    _fields_ = [("kernelLaunches", c_longlong), ("flowCgIterations", c_longlong), ("smoothCgIterations", c_longlong), ("flowSolves", c_int),
                ("smoothSolves", c_int), ("lastFlowResidual", c_double), ("lastSmoothResidual", c_double), ("flowSolveMs", c_float),
                ("smoothSolveMs", c_float), ("advectMs", c_float), ("setupMs", c_float), ("flowSpmvBytes", c_double), ("flowRows", c_longlong),
                ("flowNnz", c_longlong), ("haloEntries", c_longlong), ("solvesAboveTolerance", c_int)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class MofError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"mof error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load_library():
    """Loads libmof_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (make -C meshopticalflow_b200/csrc)")
    try:  # libmof_b200.so needs libnccl.so.2: let torch's own copy (if torch is here) be the one both of them bind to
        import torch  # noqa: F401
    except ImportError:
        pass
    lib = ctypes.CDLL(LIB_PATH)
    D, I = POINTER(c_double), POINTER(c_int)
    lib.mof_default_params.argtypes = [POINTER(Params)]
    lib.mof_default_params.restype = None
    lib.mof_create.argtypes = [c_int, c_void_p, POINTER(c_void_p)]
    lib.mof_destroy.argtypes = [c_void_p]
    lib.mof_destroy.restype = None
    lib.mof_last_error.argtypes = [c_void_p]
    lib.mof_last_error.restype = c_char_p
    lib.mof_set_params.argtypes = [c_void_p, POINTER(Params)]
    lib.mof_get_stats.argtypes = [c_void_p, POINTER(Stats)]
    lib.mof_reset_stats.argtypes = [c_void_p]
    lib.mof_reset_stats.restype = None
    lib.mof_synchronize.argtypes = [c_void_p]
    lib.mof_set_mesh.argtypes = [c_void_p, D, c_int, I, c_int]
    lib.mof_set_mesh_device.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int]
    lib.mof_set_reorder.argtypes = [c_void_p, c_int]
    lib.mof_get_permutation.argtypes = [c_void_p, I, I, I]
    lib.mof_spectrum.argtypes = [c_void_p, c_int, c_double, c_int, D, D, I, D]
    lib.mof_set_signals.argtypes = [c_void_p, D, D, c_int]
    lib.mof_set_signals_device.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    lib.mof_iterate.argtypes = [c_void_p, c_int]
    lib.mof_get_flow.argtypes = [c_void_p, D]
    lib.mof_get_coeffs.argtypes = [c_void_p, D]
    lib.mof_num_edges.argtypes = [c_void_p]
    lib.mof_num_coeffs.argtypes = [c_void_p]
    lib.mof_num_coeffs.restype = c_longlong
    lib.mof_advect_vertices.argtypes = [c_void_p, c_double, D, D]
    lib.mof_advect_vertices_device.argtypes = [c_void_p, c_double, c_void_p, c_void_p]
    lib.mof_set_texture_map.argtypes = [c_void_p, c_int, c_int, I, D, D, POINTER(c_ubyte), POINTER(c_ubyte)]
    lib.mof_advect_texels.argtypes = [c_void_p, c_double, c_int, D, D]
    lib.mof_advect_texels_frames.argtypes = [c_void_p, c_int, c_int, D, D]
    lib.mof_subdivide.argtypes = [c_void_p, POINTER(ctypes.c_float), c_int, I, D, c_int, c_double, I, I]
    lib.mof_get_subdivision.argtypes = [c_void_p, POINTER(ctypes.c_float), I, D]
    lib.mof_build_texture_map.argtypes = [c_void_p, c_int, c_int, c_int, D, POINTER(c_ubyte), POINTER(c_ubyte), I]
    lib.mof_get_texture_map.argtypes = [c_void_p, I, D]
    lib.mof_sample_textures_to_vertices.argtypes = [c_void_p, c_int, D, D]
    lib.mof_csr_size.argtypes = [c_void_p, c_int, I, POINTER(c_longlong)]
    lib.mof_get_csr.argtypes = [c_void_p, c_int, I, I, D]
    lib.mof_array_bytes.argtypes = [c_void_p, c_int]
    lib.mof_array_bytes.restype = c_longlong
    lib.mof_get_array.argtypes = [c_void_p, c_int, c_void_p]
    lib.mof_pcg_solve_csr.argtypes = [c_void_p, c_int, I, I, D, D, D, c_double, c_int, I, D]
    lib.mof_time_flow_spmv.argtypes = [c_void_p, c_int, POINTER(c_float)]
    lib.mof_time_kernel.argtypes = [c_void_p, c_int, c_int, POINTER(c_double), POINTER(c_double)]
    lib.mof_dist_unique_id.argtypes = [POINTER(c_ubyte)]
    lib.mof_dist_init.argtypes = [c_void_p, c_int, c_int, POINTER(c_ubyte)]
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(POINTER(c_double))


def _i(a):
    return a.ctypes.data_as(POINTER(c_int))


def dist_unique_id() -> bytes:
    """A fresh communicator id (call on ONE rank and broadcast it; see sharding.broadcast_bytes)."""
    buf = (c_ubyte * 128)()
    rc = load_library().mof_dist_unique_id(buf)
    if rc != MOF_OK:
        raise MofError(rc, "mof_dist_unique_id failed")
    return bytes(buf)


def default_params() -> Params:
    p = Params()
    load_library().mof_default_params(byref(p))
    return p


class Aligner:
    """One solver context on one GPU: the state the reference keeps in WhitneyFlowViewer's statics
    (flowData, vf, inputGeometryData, inputTextureData; OpticalFlow.cpp:562-567)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load_library()
        self._ctx = c_void_p()
        rc = self._lib.mof_create(device, c_void_p(stream) if stream else None, byref(self._ctx))
        if rc != MOF_OK:
            raise MofError(rc, "mof_create failed: no usable CUDA device (there is no CPU path)")
        self.V = self.T = 0

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.mof_destroy(self._ctx)
            self._ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != MOF_OK:
            raise MofError(rc, self._lib.mof_last_error(self._ctx).decode(errors="replace"))

    # --- parameters and counters
    def set_params(self, p: Params):
        self._check(self._lib.mof_set_params(self._ctx, byref(p)))

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.mof_get_stats(self._ctx, byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._lib.mof_reset_stats(self._ctx)

    def synchronize(self):
        self._check(self._lib.mof_synchronize(self._ctx))

    # --- numbering (include/mof_b200.h: mof_set_reorder): -1 decide by the locality of the caller's numbering, 0 never, 1 always
    def set_reorder(self, mode: int):
        self._check(self._lib.mof_set_reorder(self._ctx, mode))

    def permutation(self):
        """(reordered, vertex order, triangle order): new index -> the caller's index (identity if nothing was renumbered)."""
        flag = c_int(0)
        vo, to = np.empty(self.V, dtype=np.int32), np.empty(self.T, dtype=np.int32)
        self._check(self._lib.mof_get_permutation(self._ctx, byref(flag), _i(vo), _i(to)))
        return bool(flag.value), vo, to

    # --- the Spectrum tool (Spectrum.cpp, VectorLaplacianSpectrum.inl): lowest eigenpairs of the basis' vector Laplacian
    def spectrum(self, count: int = 20, tol: float = 1e-8, max_iterations: int = 2000):
        """(eigenvalues [count], fields [count, T, 2], iterations, residual)."""
        ev = np.empty(count, dtype=np.float64)
        fields = np.empty((count, self.T, 2), dtype=np.float64)
        its, res = c_int(0), c_double(0)
        self._check(self._lib.mof_spectrum(self._ctx, count, tol, max_iterations, _d(ev), _d(fields), byref(its), byref(res)))
        return ev, fields, its.value, res.value

    # --- one mesh over several GPUs (no reference counterpart): call before set_mesh, on every rank
    def dist_init(self, world: int, rank: int, unique_id: bytes):
        """Joins this context to a `world`-rank communicator: the flow solves of the meshes set afterwards are
        row-partitioned across the ranks. Every rank must then make the same calls with the same inputs."""
        if len(unique_id) != 128:
            raise MofError(MOF_E_INVALID, "the communicator id is 128 bytes (dist_unique_id() on rank 0, broadcast by the caller)")
        buf = (c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.mof_dist_init(self._ctx, world, rank, buf))

    # --- Init (OpticalFlow.cpp:787-871)
    def set_mesh(self, vertices, triangles):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        self.V, self.T = v.shape[0], t.shape[0]
        self._check(self._lib.mof_set_mesh(self._ctx, _d(v), self.V, _i(t), self.T))

    def set_mesh_device(self, d_vertices_ptr: int, V: int, d_triangles_ptr: int, T: int):
        self.V, self.T = V, T
        self._check(self._lib.mof_set_mesh_device(self._ctx, c_void_p(d_vertices_ptr), V, c_void_p(d_triangles_ptr), T))

    def set_signals(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        if a.shape != b.shape or a.shape[0] != self.V:
            raise MofError(MOF_E_INVALID, "signals must be V x channels")
        self._check(self._lib.mof_set_signals(self._ctx, _d(a), _d(b), a.shape[1]))

    def set_signals_device(self, d_a_ptr: int, d_b_ptr: int, channels: int = 3):
        self._check(self._lib.mof_set_signals_device(self._ctx, c_void_p(d_a_ptr), c_void_p(d_b_ptr), channels))

    # --- IterativeOptimization (OpticalFlow.cpp:1036-1056)
    def iterate(self, n: int):
        self._check(self._lib.mof_iterate(self._ctx, n))

    @property
    def num_edges(self) -> int:
        return self._lib.mof_num_edges(self._ctx)

    def flow(self):
        out = np.empty((self.T, 2))
        self._check(self._lib.mof_get_flow(self._ctx, _d(out)))
        return out

    @property
    def num_coeffs(self) -> int:
        """Unknowns of the flow basis in use: E (Whitney), 2V (Conformal), 2T (Connection)."""
        return self._lib.mof_num_coeffs(self._ctx)

    def coeffs(self):
        out = np.empty(self.num_coeffs)
        self._check(self._lib.mof_get_coeffs(self._ctx, _d(out)))
        return out

    def advect_vertices(self, alpha: float = 0.5, out_a=None, out_b=None):
        a = np.empty((self.V, 3)) if out_a is None else out_a
        b = np.empty((self.V, 3)) if out_b is None else out_b
        self._check(self._lib.mof_advect_vertices(self._ctx, alpha, _d(a), _d(b)))
        return a, b

    def advect_vertices_device(self, alpha: float, d_out_a_ptr: int, d_out_b_ptr: int):
        self._check(self._lib.mof_advect_vertices_device(self._ctx, alpha, c_void_p(d_out_a_ptr), c_void_p(d_out_b_ptr)))

    def set_texture_map(self, W, H, srcT, srcP, tri_uv, tex_a, tex_b):
        srcT = np.ascontiguousarray(srcT, dtype=np.int32)
        srcP = np.ascontiguousarray(srcP, dtype=np.float64)
        uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
        ta = np.ascontiguousarray(tex_a, dtype=np.uint8)
        tb = np.ascontiguousarray(tex_b, dtype=np.uint8)
        self._tex = (W, H)
        self._check(self._lib.mof_set_texture_map(self._ctx, W, H, _i(srcT), _d(srcP), _d(uv), ta.ctypes.data_as(POINTER(c_ubyte)), tb.ctypes.data_as(POINTER(c_ubyte))))

    # --- the texture configuration's one-time preparation on the device (MeshFlow.inl:158-467)
    def subdivide(self, vertices_f32, triangles, tri_uv, edge_length: float):
        """Subdivide: returns (vertices float32 [V',3], triangles int32 [T',3], tri_uv float64 [T',6])."""
        v = np.ascontiguousarray(vertices_f32, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        uv = np.ascontiguousarray(tri_uv, dtype=np.float64).reshape(-1, 6)
        nv, nt = c_int(), c_int()
        self._check(self._lib.mof_subdivide(self._ctx, v.ctypes.data_as(POINTER(ctypes.c_float)), v.shape[0], _i(t), _d(uv), t.shape[0], float(edge_length), byref(nv), byref(nt)))
        vo, to, uo = np.empty((nv.value, 3), dtype=np.float32), np.empty((nt.value, 3), dtype=np.int32), np.empty((nt.value, 6))
        self._check(self._lib.mof_get_subdivision(self._ctx, vo.ctypes.data_as(POINTER(ctypes.c_float)), _i(to), _d(uo)))
        return vo, to, uo

    def build_texture_map(self, W, H, pad, tri_uv, tex_a, tex_b):
        """GetTextureSource for the mesh in place; installs the map and the textures. Returns (srcT [W*H], srcP [W*H,2])."""
        uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
        ta = np.ascontiguousarray(tex_a, dtype=np.uint8)
        tb = np.ascontiguousarray(tex_b, dtype=np.uint8)
        self._tex = (W, H)
        misses = c_int()
        self._check(self._lib.mof_build_texture_map(self._ctx, W, H, pad, _d(uv), ta.ctypes.data_as(POINTER(c_ubyte)), tb.ctypes.data_as(POINTER(c_ubyte)), byref(misses)))
        return self.texture_map()

    def texture_map(self):
        W, H = self._tex
        srcT, srcP = np.empty(W * H, dtype=np.int32), np.empty((W * H, 2))
        self._check(self._lib.mof_get_texture_map(self._ctx, _i(srcT), _d(srcP)))
        return srcT, srcP

    def sample_textures_to_vertices(self, bilinear: bool = True):
        """SampleTextureToVertices of the two installed textures: (colours A [V,3], colours B [V,3])."""
        a, b = np.empty((self.V, 3)), np.empty((self.V, 3))
        self._check(self._lib.mof_sample_textures_to_vertices(self._ctx, 1 if bilinear else 0, _d(a), _d(b)))
        return a, b

    def advect_texels(self, alpha: float = 0.5, bilinear: bool = True):
        W, H = self._tex
        a, b = np.empty((W * H, 3)), np.empty((W * H, 3))
        self._check(self._lib.mof_advect_texels(self._ctx, alpha, 1 if bilinear else 0, _d(a), _d(b)))
        return a, b

    def advect_texels_frames(self, frames: int, bilinear: bool = True):
        """InputTextureData::flow(frames): two arrays [frames, W*H, 3]."""
        W, H = self._tex
        a, b = np.empty((frames, W * H, 3)), np.empty((frames, W * H, 3))
        self._check(self._lib.mof_advect_texels_frames(self._ctx, frames, 1 if bilinear else 0, _d(a), _d(b)))
        return a, b

    # --- debug taps
    def csr(self, which: int):
        import scipy.sparse as sp
        rows, nnz = c_int(), c_longlong()
        self._check(self._lib.mof_csr_size(self._ctx, which, byref(rows), byref(nnz)))
        rowptr, col, val = np.empty(rows.value + 1, dtype=np.int32), np.empty(nnz.value, dtype=np.int32), np.empty(nnz.value)
        self._check(self._lib.mof_get_csr(self._ctx, which, _i(rowptr), _i(col), _d(val)))
        return sp.csr_matrix((val, col, rowptr), shape=(rows.value, rows.value))

    def array(self, which: int):
        dtype, cols = _ARRAY_SPEC[which]
        nbytes = self._lib.mof_array_bytes(self._ctx, which)
        if nbytes <= 0:
            raise MofError(MOF_E_INVALID, "array not available yet")
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self._lib.mof_get_array(self._ctx, which, out.ctypes.data_as(c_void_p)))
        return out.reshape(-1, cols) if cols else out

    def pcg_solve_csr(self, A, b, tol=1e-8, max_iters=100000):
        A = A.tocsr()
        A.sort_indices()
        rowptr, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        iters, relres = c_int(), c_double()
        self._check(self._lib.mof_pcg_solve_csr(self._ctx, b.size, _i(rowptr), _i(col), _d(val), _d(b), _d(x), tol, max_iters, byref(iters), byref(relres)))
        return x, iters.value, relres.value

    def time_kernel(self, which: int, reps: int = 20):
        """(microseconds per launch, algorithmic bytes per launch) of one kernel of the solvers / the walk (mof_kernel_id)."""
        us, nbytes = c_double(), c_double()
        self._check(self._lib.mof_time_kernel(self._ctx, which, reps, byref(us), byref(nbytes)))
        return us.value, nbytes.value

    def time_flow_spmv(self, reps: int = 20) -> float:
        ms = c_float()
        self._check(self._lib.mof_time_flow_spmv(self._ctx, reps, byref(ms)))
        return ms.value


def align_vertices(vertices, triangles, colors_a, colors_b, params: Params | None = None, device: int = 0):
    """--in A.ply B.ply --out r.ply (OpticalFlow.cpp:1049-1055): returns (blended colours V x 3, flow T x 2, stats)."""
    al = Aligner(device)
    try:
        p = params or default_params()
        al.set_params(p)
        al.set_mesh(vertices, triangles)
        al.set_signals(colors_a, colors_b)
        al.iterate(p.iterations)
        a, b = al.advect_vertices(0.5)
        return (a + b) / 2.0, al.flow(), al.stats()
    finally:
        al.close()
