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


# ------------------------------------------------------------------------- textured test inputs

def uv_torus(nu: int, nv: int, major: float = 0.55, minor: float = 0.45):
    """Closed torus of nu x nv quads split into triangles, with per-face-corner texture coordinates laid out
    like the reference's Example/mesh.ply: vertex (i, j) sits at uv ((i+.5)/nu, (j+.5)/nv), so the faces on the
    two seams span the texture backwards — a good exercise for the rasteriser's first-writer rule
    (MeshFlow.inl:334). Returns (vertices float32 [V,3], triangles int32 [T,3], tri_uv float32 [T,6])."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    th, ph = 2 * np.pi * i / nu, 2 * np.pi * j / nv
    r = major + minor * np.cos(ph)
    v = np.stack([r * np.cos(th), minor * np.sin(ph), r * np.sin(th)], -1).reshape(-1, 3).astype(np.float32)
    vid = lambda a, b: (a % nu) * nv + (b % nv)
    uv = lambda a, b: np.stack([(a % nu + 0.5) / nu, (b % nv + 0.5) / nv], -1)
    a, b = i.reshape(-1), j.reshape(-1)
    t1 = np.stack([vid(a, b), vid(a + 1, b), vid(a + 1, b + 1)], 1)
    t2 = np.stack([vid(a, b), vid(a + 1, b + 1), vid(a, b + 1)], 1)
    u1 = np.concatenate([uv(a, b), uv(a + 1, b), uv(a + 1, b + 1)], 1)
    u2 = np.concatenate([uv(a, b), uv(a + 1, b + 1), uv(a, b + 1)], 1)
    tri = np.concatenate([t1, t2], 0).astype(np.int32)
    tuv = np.concatenate([u1, u2], 0).astype(np.float32)
    # orientation: outward normals
    n = np.cross(v[tri[:, 1]] - v[tri[:, 0]], v[tri[:, 2]] - v[tri[:, 0]])
    c = v[tri].mean(1)
    centre = np.stack([major * c[:, 0] / np.hypot(c[:, 0], c[:, 2]), np.zeros(len(c)), major * c[:, 2] / np.hypot(c[:, 0], c[:, 2])], 1)
    flip = ((c - centre) * n).sum(1) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    tuv[flip] = tuv[flip][:, [0, 1, 4, 5, 2, 3]]
    return v, tri, tuv


def smooth_texture_pair(width: int, height: int, seed: int, shift: float = 0.02):
    """Two uint8 [H,W,3] textures: a smooth periodic random pattern and the same pattern shifted in u and v."""
    rng = np.random.default_rng(seed)
    k = rng.integers(1, 4, size=(3, 5, 2)).astype(np.float64)
    ph = rng.uniform(0, 2 * np.pi, size=(3, 5))

    def img(du, dv):
        y, x = np.meshgrid((np.arange(height) + 0.5) / height + dv, (np.arange(width) + 0.5) / width + du, indexing="ij")
        out = np.zeros((height, width, 3))
        for c in range(3):
            s = sum(np.sin(2 * np.pi * (k[c, q, 0] * x + k[c, q, 1] * y) + ph[c, q]) for q in range(5))
            out[..., c] = 127.5 + 127.5 * np.tanh(0.5 * s)
        return np.rint(out).clip(0, 255).astype(np.uint8)

    return img(0.0, 0.0), img(shift, 0.5 * shift)


def write_ply_textured(path: str, vertices: np.ndarray, triangles: np.ndarray, tri_uv: np.ndarray) -> None:
    """ASCII: float x y z; list uchar int vertex_indices; list uchar float texcoord (like Example/mesh.ply)."""
    with open(path, "w") as fp:
        fp.write(f"ply\nformat ascii 1.0\nelement vertex {len(vertices)}\nproperty float x\nproperty float y\nproperty float z\n"
                 f"element face {len(triangles)}\nproperty list uchar int vertex_indices\nproperty list uchar float texcoord\nend_header\n")
        for p in vertices:
            fp.write("%.9g %.9g %.9g \n" % tuple(p))
        for f, uv in zip(triangles, tri_uv):
            fp.write("3 %d %d %d 6 %.9g %.9g %.9g %.9g %.9g %.9g \n" % (f[0], f[1], f[2], *uv))
