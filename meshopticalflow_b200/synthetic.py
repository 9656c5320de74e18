"""Synthetic inputs for the parity tests and bench.py.

Subdivided-octahedron unit spheres (V = 4*4^k + 2, T = 8*4^k, E = 12*4^k) with smooth random RGB
per-vertex signals; signal B is signal A's field rotated by 4 degrees about z, so the pair has a
known, smooth displacement. `numpy.random.default_rng(seed)`, seed = pair index.

Also a tiny PLY reader/writer for the three layouts the reference's CLI touches
(include/Misha/Ply.h:394-405 coloured vertices, :710-714 textured faces; output layout
OpticalFlow.cpp:147 -> Ply.inl PlyWriteTriangles), used by the test harness only — the product's own
PLY code is C++ (csrc/host/ply_io.cpp).
"""
from __future__ import annotations

import numpy as np


def octahedron_sphere(level: int, spatial_sort: bool = True):
    """Unit sphere from `level` 1-to-4 subdivisions of the octahedron; outward-facing triangles.

    Returns (vertices float64 [V,3], triangles int32 [T,3]). With spatial_sort the vertices and
    triangles are renumbered along a Morton curve of their positions / centroids, the kind of
    locality a mesh written by a modelling tool has; topology and geometry are unchanged.
    """
    v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    t = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]], dtype=np.int64)
    for _ in range(level):
        nv = v.shape[0]
        a, b, c = t[:, 0], t[:, 1], t[:, 2]
        e = np.concatenate([np.stack([a, b], 1), np.stack([b, c], 1), np.stack([c, a], 1)], 0)
        key = np.minimum(e[:, 0], e[:, 1]) * nv + np.maximum(e[:, 0], e[:, 1])
        uniq, inv = np.unique(key, return_inverse=True)
        mid = v[uniq // nv] + v[uniq % nv]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], 0)
        nt = t.shape[0]
        ab, bc, ca = nv + inv[:nt], nv + inv[nt:2 * nt], nv + inv[2 * nt:]
        t = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], 0)
    if spatial_sort:
        order = np.argsort(_morton(v), kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        v = v[order]
        t = rank[t]
        t = t[np.argsort(_morton(v[t].mean(axis=1)), kind="stable")]
    return np.ascontiguousarray(v), np.ascontiguousarray(t.astype(np.int32))


def _morton(p: np.ndarray) -> np.ndarray:
    q = np.clip(((p + 1.0) * 0.5 * 1023.0).astype(np.int64), 0, 1023)
    code = np.zeros(p.shape[0], dtype=np.int64)
    for bit in range(10):
        for axis in range(3):
            code |= ((q[:, axis] >> bit) & 1) << (3 * bit + axis)
    return code


def _field(p: np.ndarray, w: np.ndarray, freq: np.ndarray, phase: np.ndarray) -> np.ndarray:
    # per channel: sum of 6 sinusoids of (p . w), integer frequency 1..3, random phase
    proj = np.einsum("vk,cnk->vcn", p, w)
    total = np.sin(proj * freq[None] + phase[None]).sum(axis=2)
    return np.rint(127.5 + 127.5 * np.tanh(0.6 * total)).clip(0, 255).astype(np.uint8)


def smooth_rgb_pair(vertices: np.ndarray, seed: int, degrees: float = 4.0):
    """Two uint8 [V,3] signals: a smooth random field and the same field rotated about z."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((3, 6, 3))
    freq = rng.integers(1, 4, size=(3, 6)).astype(np.float64)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=(3, 6))
    th = np.deg2rad(degrees)
    rot = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    return _field(vertices, w, freq, phase), _field(vertices @ rot.T, w, freq, phase)


# --------------------------------------------------------------------------------------------- PLY

def write_ply_colored(path: str, vertices: np.ndarray, colors: np.ndarray, triangles: np.ndarray, binary: bool = True) -> None:
    """float x y z, uchar red green blue, list uchar int vertex_indices."""
    nv, nt = vertices.shape[0], triangles.shape[0]
    fmt = "binary_little_endian" if binary else "ascii"
    head = (f"ply\nformat {fmt} 1.0\nelement vertex {nv}\nproperty float x\nproperty float y\nproperty float z\n"
            f"property uchar red\nproperty uchar green\nproperty uchar blue\nelement face {nt}\n"
            f"property list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as fp:
        fp.write(head.encode())
        if binary:
            vrec = np.zeros(nv, dtype=[("p", "<f4", 3), ("c", "u1", 3)])
            vrec["p"], vrec["c"] = vertices.astype(np.float32), colors.astype(np.uint8)
            fp.write(vrec.tobytes())
            frec = np.zeros(nt, dtype=[("n", "u1"), ("i", "<i4", 3)])
            frec["n"], frec["i"] = 3, triangles
            fp.write(frec.tobytes())
        else:
            for p, c in zip(vertices.astype(np.float32), colors.astype(np.uint8)):
                fp.write(("%g %g %g %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2])).encode())
            for f in triangles:
                fp.write(("3 %d %d %d\n" % (f[0], f[1], f[2])).encode())


_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4", "float": "<f4", "double": "<f8",
              "int8": "i1", "uint8": "u1", "int16": "<i2", "uint16": "<u2", "int32": "<i4", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}


def read_ply(path: str):
    """Returns dict: 'vertex' -> {name: array}, 'face' -> {'vertex_indices': [T,3] int32, 'texcoord': [T,6] (if present)}.
    Handles ascii and binary_little_endian, triangles only."""
    with open(path, "rb") as fp:
        data = fp.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode().split("\n")
    fmt, elements = None, []
    for ln in lines:
        tok = ln.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], "list", tok[2], tok[3]))
            else:
                elements[-1][2].append((tok[2], tok[1]))
    out = {}
    if fmt == "ascii":
        rows = data[end:].decode().split("\n")
        pos = 0
        for name, count, props in elements:
            block = rows[pos:pos + count]
            pos += count
            if all(len(p) == 2 for p in props):
                arr = np.array([r.split() for r in block], dtype=np.float64).reshape(count, len(props))
                out[name] = {p[0]: arr[:, i] for i, p in enumerate(props)}
            else:
                vals = [np.array(r.split(), dtype=np.float64) for r in block]
                res, off = {}, 0
                for p in props:
                    n = int(vals[0][off])
                    res[p[0]] = np.stack([v[off + 1:off + 1 + n] for v in vals])
                    off += 1 + n
                out[name] = res
    elif fmt == "binary_little_endian":
        off = end
        for name, count, props in elements:
            if all(len(p) == 2 for p in props):
                dt = np.dtype([(p[0], _PLY_TYPES[p[1]]) for p in props])
                arr = np.frombuffer(data, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                out[name] = {p[0]: arr[p[0]] for p in props}
            else:
                fields = []
                for p in props:
                    n = data[off + sum(np.dtype(f[1]).itemsize * (f[2] if len(f) > 2 else 1) for f in fields)]
                    fields.append((p[0] + "_n", _PLY_TYPES[p[2]]))
                    fields.append((p[0], _PLY_TYPES[p[3]], int(n)))
                dt = np.dtype([(f[0], f[1], (f[2],)) if len(f) > 2 else (f[0], f[1]) for f in fields])
                arr = np.frombuffer(data, dtype=dt, count=count, offset=off)
                off += dt.itemsize * count
                out[name] = {p[0]: arr[p[0]] for p in props}
    else:
        raise ValueError(f"unsupported PLY format {fmt}")
    if "face" in out and "vertex_indices" in out["face"]:
        out["face"]["vertex_indices"] = out["face"]["vertex_indices"].astype(np.int32)
    return out
