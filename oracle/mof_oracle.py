"""TEST INFRASTRUCTURE ONLY — CPU oracle for the halfway-alignment solver loop.

A numpy/scipy restatement (not a copy) of the reference algorithm, function by function, each citing
the reference file:line it follows (paths relative to /root/reference). The serial triangle walks,
the uv rasteriser and the edge subdivision are plain C (oracle/walk.c, loaded through ctypes).
Direct sparse solves use scipy's SuperLU where the reference uses Eigen 3.2.7 SimplicialLLT/LDLT
(include/Misha/LinearSolvers.h:249-391); both are exact factorisations, so results agree to
round-off.

Pinning: the reference ships no tests or golden vectors for this path. This oracle is
pinned against the reference itself — oracle/_ref/OpticalFlow_ref, the unmodified reference sources
compiled by oracle/ref/build_ref.sh, run with --tap — by tests/test_oracle_vs_reference.py in this
container, and against the committed fixtures under tests/golden/ (generated from that binary by
tests/golden/make_golden.py) everywhere else.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module. The
product (meshopticalflow_b200/) never does.

Conventions: half-edge h = 3t+j is the edge of triangle t opposite corner j; 2x2 matrices are
row-major in standard notation; metric rows are (g00, g01, g11).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_walk.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "walk.c")):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_walk.so"], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


_D, _I, _U8 = ctypes.c_double, ctypes.c_int, ctypes.c_ubyte

GRAD = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # hat-function gradients, FEM.inl:441-443
CORNER = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])  # FEM.inl:552


# ------------------------------------------------------------------------------------- mesh setup

def metric_from_embedding(vertices, triangles):
    """setMetricFromEmbedding, FEM.inl:1305-1323 -> g [T,3] = (e0.e0, e0.e1, e1.e1)."""
    e0 = vertices[triangles[:, 1]] - vertices[triangles[:, 0]]
    e1 = vertices[triangles[:, 2]] - vertices[triangles[:, 0]]
    return np.stack([(e0 * e0).sum(1), (e0 * e1).sum(1), (e1 * e1).sum(1)], 1)


def det(g):
    return g[:, 0] * g[:, 2] - g[:, 1] * g[:, 1]


def make_unit_area(g):
    """makeUnitArea, FEM.inl:1283-1291: g *= 2 / sum(sqrt(det g))."""
    return g * (2.0 / np.sqrt(det(g)).sum())


def inverse_metric(g):
    """setInverseMetric, FEM.inl:1363-1369 (2x2 inverse, Geometry.inl:277-289)."""
    d = 1.0 / det(g)
    return np.stack([g[:, 2] * d, -g[:, 1] * d, g[:, 0] * d], 1)


def triangle_areas(g):
    """area(i), FEM.inl:1301."""
    return np.sqrt(det(g)) / 2.0


def _gmul(g, v):
    return np.stack([g[:, 0] * v[:, 0] + g[:, 1] * v[:, 1], g[:, 1] * v[:, 0] + g[:, 2] * v[:, 1]], 1)


def _rotate90(g, gi, v):
    """FEM::Rotate90, FEM.inl:18-24."""
    w = _gmul(gi, np.stack([-v[:, 1], v[:, 0]], 1))
    vn = (_gmul(g, v) * v).sum(1)
    wn = (_gmul(g, w) * w).sum(1)
    scale = np.where(wn != 0, np.sqrt(vn / np.where(wn != 0, wn, 1.0)), 1.0)
    return w * scale[:, None]


def opposite_half_edges(triangles):
    """setEdgeXForms, FEM.inl:592-614 (adjacency part). Returns opp [3T] (-1 on a boundary).
    Raises on a half-edge used twice ('[ERROR] Edge is occupied', FEM.inl:599)."""
    T = triangles.shape[0]
    nv = int(triangles.max()) + 1
    j = np.arange(3)
    a = triangles[:, (j + 1) % 3].astype(np.int64)  # edge j runs corner j+1 -> corner j+2
    b = triangles[:, (j + 2) % 3].astype(np.int64)
    key = (a * nv + b).reshape(-1)
    rev = (b * nv + a).reshape(-1)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    if np.any(skey[1:] == skey[:-1]):
        raise ValueError("[ERROR] Edge is occupied")
    pos = np.searchsorted(skey, rev)
    pos_c = np.minimum(pos, skey.size - 1)
    found = skey[pos_c] == rev
    return np.where(found, order[pos_c], -1).astype(np.int32)


def edge_xforms(g, opp):
    """_setEdgeXForm, FEM.inl:550-590. Returns (linear [3T,4] row-major, constant [3T,2]).
    Boundary half-edges keep the identity (CoordinateXForm ctor, FEM.h:126)."""
    n = opp.size
    gi = inverse_metric(g)
    lin = np.tile(np.array([1.0, 0.0, 0.0, 1.0]), (n, 1))
    cst = np.zeros((n, 2))
    h = np.nonzero(opp >= 0)[0]
    o = opp[h]
    t, ot = h // 3, o // 3
    v0, v1 = (h + 1) % 3, (h + 2) % 3
    ov0, ov1 = (o + 1) % 3, (o + 2) % 3
    ed = CORNER[v1] - CORNER[v0]
    oed = -(CORNER[ov1] - CORNER[ov0])
    ed = ed / np.sqrt((ed * _gmul(g[t], ed)).sum(1))[:, None]
    oed = oed / np.sqrt((oed * _gmul(g[ot], oed)).sum(1))[:, None]
    ep = _rotate90(g[t], gi[t], ed)
    oep = _rotate90(g[ot], gi[ot], oed)
    # M has columns (ed, ep), oM has columns (oed, oep); linear = oM * M^-1 (FEM.inl:575-580)
    m00, m01, m10, m11 = ed[:, 0], ep[:, 0], ed[:, 1], ep[:, 1]
    d = 1.0 / (m00 * m11 - m01 * m10)
    i00, i01, i10, i11 = m11 * d, -m01 * d, -m10 * d, m00 * d
    o00, o01, o10, o11 = oed[:, 0], oep[:, 0], oed[:, 1], oep[:, 1]
    L = np.stack([o00 * i00 + o01 * i10, o00 * i01 + o01 * i11, o10 * i00 + o11 * i10, o10 * i01 + o11 * i11], 1)
    s = CORNER[v0] + CORNER[v1]
    os_ = CORNER[ov0] + CORNER[ov1]
    c = (os_ - np.stack([L[:, 0] * s[:, 0] + L[:, 1] * s[:, 1], L[:, 2] * s[:, 0] + L[:, 3] * s[:, 1]], 1)) / 2.0
    lin[h], cst[h] = L, c
    return lin, cst


def scalar_matrices(g, triangles, nv):
    """_scalarMatrix with SetScalarMassMatrix / SetScalarStiffnessMatrix, FEM.inl:1507-1547, 439-496.
    Returns (mass, stiffness) as CSR with sorted columns (the reference stores the diagonal first,
    then ascending columns: same pattern)."""
    sq = np.sqrt(det(g))
    gi = inverse_metric(g)
    mloc = np.full((3, 3), 1.0 / 24) + np.eye(3) * (1.0 / 12 - 1.0 / 24)
    rows = np.repeat(triangles, 3, axis=1).reshape(-1)
    cols = np.tile(triangles, (1, 3)).reshape(-1)
    mvals = (sq[:, None] * mloc.reshape(-1)[None]).reshape(-1)
    gg = np.zeros((g.shape[0], 3, 3))
    for i in range(3):
        for j in range(3):
            gj = np.stack([gi[:, 0] * GRAD[j, 0] + gi[:, 1] * GRAD[j, 1], gi[:, 1] * GRAD[j, 0] + gi[:, 2] * GRAD[j, 1]], 1)
            gg[:, i, j] = (GRAD[i, 0] * gj[:, 0] + GRAD[i, 1] * gj[:, 1]) / 2.0 * sq
    M = sp.coo_matrix((mvals, (rows, cols)), shape=(nv, nv)).tocsr()
    S = sp.coo_matrix((gg.reshape(-1), (rows, cols)), shape=(nv, nv)).tocsr()
    M.sort_indices(), S.sort_indices()
    return M, S


def integral(g, triangles, x):
    """getIntegral, FEM.inl:2081-2098: sum_t sum_j x[tri[t][j]] * rowsum(mass_t)[j]."""
    return float((x[triangles] * (np.sqrt(det(g)) * (1.0 / 12 + 2.0 / 24))[:, None]).sum())


# ---------------------------------------------------------------------------- Whitney 1-form basis

@dataclass
class Whitney:
    reduced: np.ndarray      # [3T] half-edge -> edge dof          (Whitney.inl:33-51)
    expanded: np.ndarray     # [E] edge dof -> its first half-edge (Whitney.inl:42)
    positive: np.ndarray     # [3T] bool, half-edge has the dof's orientation (Whitney.inl:47)
    P: sp.csr_matrix         # 2T x E prolongation                 (Whitney.inl:65-88)
    S: sp.csr_matrix         # E x E smoothness operator           (Whitney.inl:92-180)


def whitney_numbering(opp):
    """InitializeCoefficients, Whitney.inl:28-62: dofs numbered by a first-touch scan over (t,j)."""
    h = np.arange(opp.size)
    first = (opp < 0) | (h < opp)
    expanded = h[first].astype(np.int32)
    rank = np.cumsum(first) - 1
    reduced = np.where(first, rank, rank[np.maximum(opp, 0)]).astype(np.int32)
    return reduced, expanded, first


def whitney_prolongation(g, reduced, positive):
    """InitializeProlonagtionOperator, Whitney.inl:65-88: row 2t+{0,1}, entry k = +-ginv*(grad[k+2]-grad[k+1])/3."""
    T = g.shape[0]
    gi = inverse_metric(g)
    rows, cols, vals = [], [], []
    t = np.arange(T)
    for k in range(3):
        d = (GRAD[(k + 2) % 3] - GRAD[(k + 1) % 3]) / 3.0
        gd = np.stack([gi[:, 0] * d[0] + gi[:, 1] * d[1], gi[:, 1] * d[0] + gi[:, 2] * d[1]], 1)
        sign = np.where(positive[3 * t + k], 1.0, -1.0)
        for r in range(2):
            rows.append(2 * t + r), cols.append(reduced[3 * t + k]), vals.append(gd[:, r] * sign)
    E = int(reduced.max()) + 1
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * T, E)).tocsr()


def whitney_smooth_operator(g, triangles, opp, reduced, expanded, positive, nv):
    """InitializeSmoothOperator, Whitney.inl:92-180: S = (d1^T m2 d1 + m1 d0 m0^-1 d0^T m1) / 2."""
    T, E = g.shape[0], expanded.size
    gi = inverse_metric(g)
    area = triangle_areas(g)
    et, ev = expanded // 3, expanded % 3
    e = np.arange(E)
    d0 = sp.coo_matrix((np.concatenate([-np.ones(E), np.ones(E)]),
                        (np.concatenate([e, e]), np.concatenate([triangles[et, (ev + 1) % 3], triangles[et, (ev + 2) % 3]]))), shape=(E, nv)).tocsr()
    t3 = np.repeat(np.arange(T), 3)
    d1 = sp.coo_matrix((np.where(positive, 1.0, -1.0), (t3, reduced)), shape=(T, E)).tocsr()
    m0 = np.zeros(nv)
    np.add.at(m0, triangles.reshape(-1), np.repeat(area / 3.0, 3))

    def cot_term(h):  # -area(t) * grad[v+1] . ginv grad[v+2]   (Whitney.inl:146)
        t, v = h // 3, h % 3
        a, b = GRAD[(v + 1) % 3], GRAD[(v + 2) % 3]
        gb = np.stack([gi[t, 0] * b[:, 0] + gi[t, 1] * b[:, 1], gi[t, 1] * b[:, 0] + gi[t, 2] * b[:, 1]], 1)
        return -area[t] * (a * gb).sum(1)

    m1 = cot_term(expanded)
    o = opp[expanded]
    has = o >= 0
    m1[has] += cot_term(o[has])
    rot = d1.T @ sp.diags(1.0 / area) @ d1
    div = sp.diags(m1) @ d0 @ sp.diags(1.0 / m0) @ d0.T @ sp.diags(m1)
    # Structural union of the two patterns, zeros kept (the reference's SpGEMM never drops an entry,
    # SparseMatrix.inl:358-425).
    pat = ((abs(d1).T @ abs(d1)) + (abs(d0) @ abs(d0).T)).tocoo()
    parts = [((rot + div) * 0.5).tocoo(), sp.coo_matrix((np.zeros(pat.nnz), (pat.row, pat.col)), shape=pat.shape)]
    S = sp.coo_matrix((np.concatenate([p.data for p in parts]), (np.concatenate([p.row for p in parts]), np.concatenate([p.col for p in parts]))),
                      shape=(E, E)).tocsr()  # coo->csr sums duplicates and keeps explicit zeros
    S.sort_indices()
    return S, m0, m1


# ------------------------------------------------------ Conformal and Connection bases (--vfMode 1|2)

ROT_GRAD = np.array([[1.0, -1.0], [0.0, 1.0], [-1.0, 0.0]])   # Conformal.inl:56
EDGE_DIR = np.array([[-1.0, 1.0], [0.0, -1.0], [1.0, 0.0]])   # Connection.inl:35


def conformal_field(g, triangles, nv, K):
    """ConformalVectorField, Conformal.inl:12-82. Unknowns: 2V (a potential and a co-potential per vertex).
    P (2T x 2V): row 2t+r, entry k = (ginv grad_k)[r] on vertex k, (rotGrad_k)[r]/sqrt(det g) on V + vertex k.
    S = blockdiag(B, B), B = K diag(1/m) K / 2 with m the LUMPED mass diagonal (sum of sqrt(det)/6, FEM.inl:474)
    and K the scalar stiffness matrix."""
    T = g.shape[0]
    gi = inverse_metric(g)
    sq = np.sqrt(det(g))
    t = np.arange(T)
    rows, cols, vals = [], [], []
    for k in range(3):
        pg = np.stack([gi[:, 0] * GRAD[k, 0] + gi[:, 1] * GRAD[k, 1], gi[:, 1] * GRAD[k, 0] + gi[:, 2] * GRAD[k, 1]], 1)
        for r in range(2):
            rows.append(2 * t + r), cols.append(triangles[:, k]), vals.append(pg[:, r])
            rows.append(2 * t + r), cols.append(triangles[:, k] + nv), vals.append(ROT_GRAD[k, r] / sq)
    P = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * T, 2 * nv)).tocsr()
    m = np.zeros(nv)
    np.add.at(m, triangles.reshape(-1), np.repeat(sq / 6.0, 3))
    Bl = (K @ sp.diags(1.0 / m) @ K) * 0.5
    S = sp.block_diag([Bl, Bl]).tocsr()
    S.sort_indices()
    return P, S


def connection_weights(g, opp, lin, cst, mode):
    """The per-half-edge weight l of ConnectionVectorField::InitializeSmoothOperator, Connection.inl:55-70.
    mode 0: |edge|^2 / (4 (area_i + area_ii) / 3); 1: ((area_i + area_ii) / 3) / |centroid_i - centroid_ii|^2
    (the neighbour's centroid unfolded into i's chart); 2: 1 / (cot_i + cot_ii)."""
    area = triangle_areas(g)
    h = np.arange(opp.size)
    i, j = h // 3, h % 3
    o = opp
    ii, jj = o // 3, o % 3

    def gdot(gt, a, b):
        return a[:, 0] * (gt[:, 0] * b[:, 0] + gt[:, 1] * b[:, 1]) + a[:, 1] * (gt[:, 1] * b[:, 0] + gt[:, 2] * b[:, 1])

    if mode == 0:
        return gdot(g[i], EDGE_DIR[j], EDGE_DIR[j]) / (4.0 * (area[i] + area[ii]) / 3.0)
    if mode == 1:
        c = 1.0 / 3
        xc = np.stack([lin[o, 0] * c + lin[o, 1] * c + cst[o, 0], lin[o, 2] * c + lin[o, 3] * c + cst[o, 1]], 1)
        d = np.array([c, c])[None] - xc
        return ((area[i] + area[ii]) / 3.0) / gdot(g[i], d, d)
    if mode == 2:
        ci = gdot(g[i], -EDGE_DIR[(j + 1) % 3], EDGE_DIR[(j + 2) % 3]) / (2.0 * area[i])
        cii = gdot(g[ii], -EDGE_DIR[(jj + 1) % 3], EDGE_DIR[(jj + 2) % 3]) / (2.0 * area[ii])
        return 1.0 / (ci + cii)
    raise ValueError("Undefined Connection Mode")


def connection_field(g, opp, lin, cst, mode):
    """ConnectionVectorField, Connection.inl:22-104. Unknowns: 2T (one tangent vector per triangle, in the
    triangle's chart), P = identity. Row block i of S: sum_j l_j g_i on the diagonal, -l_j g_i L_j towards the
    neighbour across edge j, L_j = the linear part of the transform from the neighbour's chart into i's."""
    T = g.shape[0]
    l = connection_weights(g, opp, lin, cst, mode)
    h = np.arange(3 * T)
    i, o = h // 3, opp
    ii = o // 3
    G = np.stack([g[:, 0], g[:, 1], g[:, 1], g[:, 2]], 1)  # row-major 2x2
    Lo = lin[o]
    X = np.stack([G[i, 0] * Lo[:, 0] + G[i, 1] * Lo[:, 2], G[i, 0] * Lo[:, 1] + G[i, 1] * Lo[:, 3],
                  G[i, 2] * Lo[:, 0] + G[i, 3] * Lo[:, 2], G[i, 2] * Lo[:, 1] + G[i, 3] * Lo[:, 3]], 1)  # g_i L, row-major
    rows, cols, vals = [], [], []
    for r in range(2):
        for c in range(2):
            rows.append(2 * i + r), cols.append(2 * i + c), vals.append(l * G[i, 2 * r + c])
            rows.append(2 * i + r), cols.append(2 * ii + c), vals.append(-l * X[:, 2 * r + c])
    S = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * T, 2 * T)).tocsr()
    S.sort_indices()
    return sp.identity(2 * T, format="csr"), S


# ------------------------------------------------------------------------------------ Spectrum tool

def vector_field_mass(g, P):
    """massOperator = restriction * vfMass * prolongation, VectorLaplacianSpectrum.inl:9-19: vfMass = blockdiag(g_t * area_t)."""
    area = triangle_areas(g)
    T = g.shape[0]
    t = np.arange(T)
    rows = np.concatenate([2 * t, 2 * t, 2 * t + 1, 2 * t + 1])
    cols = np.concatenate([2 * t, 2 * t + 1, 2 * t, 2 * t + 1])
    vals = np.concatenate([g[:, 0] * area, g[:, 1] * area, g[:, 1] * area, g[:, 2] * area])
    vf_mass = sp.coo_matrix((vals, (rows, cols)), shape=(2 * T, 2 * T)).tocsr()
    return (P.T @ vf_mass @ P).tocsc()


def spectrum(vertices, triangles, count=20, vf_mode=0, c_mode=0, shift=1e-8):
    """ComputeSpectrum, VectorLaplacianSpectrum.inl:5-39, as Spectrum.cpp:177-184 calls it: the `count` eigenpairs of S x = lambda M x
    nearest `shift` by shift-invert Lanczos (scipy's eigsh is the same ARPACK driver the reference calls through ARPACK++,
    EigenvalueSolver.h:177-219, with SuperLU where the reference uses Eigen's LDLT), prolonged to per-triangle fields.
    Returns (eigenvalues ascending [count], fields [count, T, 2], coefficient vectors [count, N], S, M)."""
    import scipy.sparse.linalg as spla

    triangles = np.asarray(triangles)
    nv = vertices.shape[0]
    g = make_unit_area(metric_from_embedding(vertices, triangles))
    opp = opposite_half_edges(triangles)
    if vf_mode == 0:
        reduced, expanded, positive = whitney_numbering(opp)
        P = whitney_prolongation(g, reduced, positive)
        S = whitney_smooth_operator(g, triangles, opp, reduced, expanded, positive, nv)[0]
    elif vf_mode == 1:
        K = scalar_matrices(g, triangles, nv)[1]
        P, S = conformal_field(g, triangles, nv, K)
    else:
        lin, cst = edge_xforms(g, opp)
        P, S = connection_field(g, opp, lin, cst, c_mode)
    M = vector_field_mass(g, P)
    # Conformal basis: the constants of either potential are in the null space of S AND of M (P maps them to the zero field), where
    # the shift-invert problem is not defined (the reference's own answer there is whatever its LDLT makes of a singular pencil). The
    # checker works on the quotient by the constants: one vertex of each potential is pinned (rows / columns 0 and V removed), which
    # leaves the eigenvalues and the prolonged fields unchanged.
    n_full = S.shape[0]
    keep = np.setdiff1d(np.arange(n_full), [0, nv]) if vf_mode == 1 else np.arange(n_full)
    Sk, Mk = S.tocsr()[keep][:, keep].tocsc(), M.tocsr()[keep][:, keep].tocsc()
    # A single-vector Lanczos process can miss copies of a multiple eigenvalue (a sphere's come in clusters of 3-6): ask for a few
    # more pairs than wanted, in a Krylov space four times that, from a fixed start vector, and keep the lowest `count`.
    n = Sk.shape[0]
    k = min(count + 8, n - 2)
    v0 = np.random.default_rng(0).standard_normal(n)
    vals, vk = spla.eigsh(Sk, k=k, M=Mk, sigma=shift, which="LM", ncv=min(n - 1, max(4 * k, 80)), v0=v0)
    order = np.argsort(vals)[:count]
    vals = vals[order]
    vecs = np.zeros((n_full, count))
    vecs[keep] = vk[:, order]
    T = triangles.shape[0]
    fields = np.stack([(P @ vecs[:, i]).reshape(T, 2) for i in range(count)])
    return vals, fields, vecs.T.copy(), S, M


# -------------------------------------------------------------------------------- solver pieces

def smooth_signal(M, S, signal, weight):
    """FlowData::smoothSignal, OpticalFlow.cpp:351-370: X = (M + w S)^-1 M B, channel by channel."""
    solve = spla.factorized((M + S * weight).tocsc())
    return np.stack([solve(M @ signal[:, c]) for c in range(signal.shape[1])], 1)


def dog_preprocess(M, S, g, triangles, signal, weight):
    """Difference-of-Gaussians normalisation, OpticalFlow.cpp:822-857 (Channels == 3 branch)."""
    solve = spla.factorized((M + S * weight).tocsc())
    out = np.empty_like(signal)
    for c in range(signal.shape[1]):
        x = signal[:, c].copy()
        b = M @ x
        old_avg = integral(g, triangles, x)
        old_var = float(x @ b) - old_avg * old_avg
        x = signal[:, c] - solve(b)
        b = M @ x
        new_avg = integral(g, triangles, x)
        new_var = float(x @ b) - new_avg * new_avg
        out[:, c] = (x - new_avg) * np.sqrt(old_var / new_var) + old_avg
    return out


def dog_blend(signal, dog, weight):
    """The Channels == 6 branch, OpticalFlow.cpp:849-855: (raw * (1 - w), DoG * w) side by side."""
    return np.hstack([signal * (1.0 - weight), dog * weight])


def data_term(triangles, areas, a, b):
    """SetDataTerm, OpticalFlow.cpp:395-421 with the k<2 semantics of the optimised reference build
    (SURVEY.md §8a a11). Returns D [T,3] = (d00, d01, d11) and rhs [T,2]."""
    va, vb = a[triangles], b[triangles]          # [T,3 corners,C]
    f = (va + vb) / 2.0
    mean_diff = (va - vb).sum(1) / 3.0            # [T,C]
    g0 = f[:, 1] - f[:, 0]
    g1 = f[:, 2] - f[:, 0]
    D = np.stack([(g0 * g0).sum(1), (g0 * g1).sum(1), (g1 * g1).sum(1)], 1) * areas[:, None]
    rhs = np.stack([(g0 * mean_diff).sum(1), (g1 * mean_diff).sum(1)], 1) * areas[:, None]
    return D, rhs


def block_diag_data(D):
    T = D.shape[0]
    t = np.arange(T)
    rows = np.concatenate([2 * t, 2 * t, 2 * t + 1, 2 * t + 1])
    cols = np.concatenate([2 * t, 2 * t + 1, 2 * t, 2 * t + 1])
    vals = np.concatenate([D[:, 0], D[:, 1], D[:, 1], D[:, 2]])
    return sp.coo_matrix((vals, (rows, cols)), shape=(2 * T, 2 * T)).tocsr()


def flow_system(w: Whitney, D, rhs, weight):
    """VectorField::UpdateOpticalFlow, VectorField.h:51-67: (A, b, Dt, scale) with
    Dt = s*R D P, b = s*R rhs, s = 1/||R D P||_F, A = Dt + weight*S."""
    Dt = (w.P.T @ block_diag_data(D) @ w.P).tocsr()
    b = w.P.T @ rhs.reshape(-1)
    scale = 1.0 / np.sqrt((Dt.data ** 2).sum())
    Dt = Dt * scale
    b = b * scale
    A = (Dt + w.S * weight).tocsr()
    A.sort_indices()
    return A, b, Dt, scale


def update_optical_flow(w: Whitney, coeffs, D, rhs, weight):
    """VectorField.h:46-104: solve, optimal step, coeffs += step*x; returns (coeffs, tField [T,2], x)."""
    A, b, Dt, _ = flow_system(w, D, rhs, weight)
    x = spla.spsolve(A.tocsc(), b)
    denom = float(x @ (Dt @ x))
    num = float(x @ b)
    step = num / denom if denom else 0.0
    if step:
        coeffs = coeffs + x * step
    return coeffs, (w.P @ coeffs).reshape(-1, 2), x


# ------------------------------------------------------------------------------------ the walks

def resample_signal(triangles, opp, lin, cst, g, tfield, signal, length):
    """ResampleSignal, OpticalFlow.cpp:198-216 (serial C walk)."""
    T, V, C = triangles.shape[0], signal.shape[0], signal.shape[1]
    tri = np.ascontiguousarray(triangles, dtype=np.int32)
    sig = np.ascontiguousarray(signal, dtype=np.float64)
    tf = np.ascontiguousarray(tfield, dtype=np.float64)
    out = np.empty_like(sig)
    _lib().mof_oracle_resample(_I(T), _I(V), _p(tri, _I), _p(opp, _I), _p(lin, _D), _p(cst, _D), _p(g, _D), _p(tf, _D), _p(sig, _D), _p(out, _D), _I(C), _D(length))
    return out


def texture_source(tri_uv, W, H, pad, opp, lin, cst, g):
    """GetTextureSource, MeshFlow.inl:411-467. tri_uv [T,6]. Returns (srcT [H*W], srcP [H*W,2])."""
    T = tri_uv.shape[0]
    uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
    srcT = np.empty(W * H, dtype=np.int32)
    srcP = np.empty((W * H, 2), dtype=np.float64)
    fails = _lib().mof_oracle_texture_source(_I(T), _p(uv, _D), _I(W), _I(H), _I(pad), _p(opp, _I), _p(lin, _D), _p(cst, _D), _p(g, _D), _p(srcT, _I), _p(srcP, _D))
    if fails:
        raise RuntimeError("[ERROR] FEM::Mesh::exp: ray does not intersect triangle (%d texels)" % fails)
    return srcT, srcP


def advect_texels(W, H, srcT, srcP, opp, lin, cst, g, tfield, tri_uv, texture, length, bilinear=True, init=None):
    """InputTextureData::flow, OpticalFlow.cpp:501-515. texture uint8 [H,W,3] top row first. Texels
    outside the map keep `init` (the viewer initialises them to the flipped input, :889)."""
    tex = np.ascontiguousarray(texture, dtype=np.uint8)
    if init is None:
        init = tex[::-1].reshape(-1, 3).astype(np.float64)
    out = np.ascontiguousarray(init, dtype=np.float64).copy()
    uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
    tf = np.ascontiguousarray(tfield, dtype=np.float64)
    _lib().mof_oracle_advect_texels(_I(W), _I(H), _p(srcT, _I), _p(srcP, _D), _p(opp, _I), _p(lin, _D), _p(cst, _D), _p(g, _D), _p(tf, _D), _p(uv, _D), _p(tex, _U8), _D(length), _I(1 if bilinear else 0), _p(out, _D))
    return out


def advect_texels_frames(W, H, frames, srcT, srcP, opp, lin, cst, g, tfield, tri_uv, texture, sign, bilinear=True):
    """InputTextureData::flow(frames), OpticalFlow.cpp:517-539: [frames, W*H, 3]; sign -1 for the first signal, +1 for the second."""
    tex = np.ascontiguousarray(texture, dtype=np.uint8)
    out = np.empty((frames, W * H, 3), dtype=np.float64)
    uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
    tf = np.ascontiguousarray(tfield, dtype=np.float64)
    _lib().mof_oracle_advect_texels_frames(_I(W), _I(H), _I(frames), _p(srcT, _I), _p(srcP, _D), _p(opp, _I), _p(lin, _D), _p(cst, _D), _p(g, _D), _p(tf, _D), _p(uv, _D),
                                           _p(tex, _U8), _D(sign), _I(1 if bilinear else 0), _p(out, _D))
    return out


def sample_texture(texture, uv, bilinear=True):
    """Sample, MeshFlow.inl:66-84."""
    tex = np.ascontiguousarray(texture, dtype=np.uint8)
    H, W = tex.shape[0], tex.shape[1]
    q = np.ascontiguousarray(uv, dtype=np.float64)
    out = np.empty((q.shape[0], 3))
    _lib().mof_oracle_sample_texture(_p(tex, _U8), _I(W), _I(H), _I(q.shape[0]), _p(q, _D), _I(1 if bilinear else 0), _p(out, _D))
    return out


def subdivide(vertices_f32, triangles, tri_uv, edge_length):
    """Subdivide, MeshFlow.inl:158-232: repeat passes until no edge is longer than edge_length."""
    v = np.ascontiguousarray(vertices_f32, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    uv = np.ascontiguousarray(tri_uv, dtype=np.float64)
    while True:
        V, T = v.shape[0], t.shape[0]
        vout = np.empty((V + 3 * T, 3), dtype=np.float32)
        tout = np.empty((4 * T, 3), dtype=np.int32)
        uvout = np.empty((4 * T, 6), dtype=np.float64)
        nT = _I(0)
        added = _lib().mof_oracle_subdivide_pass(_I(V), _I(T), _p(v, ctypes.c_float), _p(t, _I), _p(uv, _D), _D(edge_length), _p(vout, ctypes.c_float), _p(tout, _I), _p(uvout, _D), ctypes.byref(nT))
        if not added:
            return v, t, uv
        v, t, uv = vout[:V + added].copy(), tout[:nT.value].copy(), uvout[:nT.value].copy()


def sample_texture_to_vertices(triangles, tri_uv, nv, texture, bilinear=True):
    """SampleTextureToVertices, MeshFlow.inl:252-266: wedge-averaged vertex colours."""
    c = sample_texture(texture, tri_uv.reshape(-1, 2), bilinear)
    out = np.zeros((nv, 3))
    cnt = np.zeros(nv)
    np.add.at(out, triangles.reshape(-1), c)
    np.add.at(cnt, triangles.reshape(-1), 1.0)
    return out / cnt[:, None]


# --------------------------------------------------------------------------------- the whole path

@dataclass
class Params:
    """Defaults of OpticalFlow.cpp:56-63 and _main, :1062-1069 (Whitney mode)."""
    iterations: int = 10
    sSmooth: float = float(np.float32(3e-3))
    sMultiply: float = 0.25
    vfSmooth: float | None = None   # None = the mode's default, OpticalFlow.cpp:1067-1069
    vMultiply: float = 1.0
    vfSThreshold: float = float(np.float32(1e-8))
    dogWeight: float = 1.0
    dogSmooth: float = float(np.float32(1e-4))
    eLength: float = float(np.float32(0.006))
    pad: int = 2
    vfMode: int = 0   # 0 Whitney, 1 Conformal, 2 Connection (VectorField.h:3-7); vfSmooth defaults 3e-6 / 5e-7 / 1e4 (:1067-1069)
    cMode: int = 0    # Connection.inl:1-5
    logSpace: bool = False  # --log, OpticalFlow.cpp:821: the comparison signals only

    def __post_init__(self):
        if self.vfSmooth is None:
            self.vfSmooth = (3e-6, 5e-7, 1e4)[self.vfMode]


@dataclass
class State:
    vertices: np.ndarray
    triangles: np.ndarray
    g: np.ndarray
    opp: np.ndarray
    lin: np.ndarray
    cst: np.ndarray
    area: np.ndarray
    M: sp.csr_matrix
    S: sp.csr_matrix
    whitney: Whitney
    signals: list
    coeffs: np.ndarray
    tfield: np.ndarray
    taps: dict = field(default_factory=dict)


def init(vertices, triangles, sig_a, sig_b, params: Params) -> State:
    """WhitneyFlowViewer::Init from the mesh onwards, OpticalFlow.cpp:787-871."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32)
    nv = vertices.shape[0]
    g = np.ascontiguousarray(make_unit_area(metric_from_embedding(vertices, triangles)))
    opp = opposite_half_edges(triangles)
    if np.any(opp < 0):
        raise ValueError("[ERROR] Boundary edge (TriangleMesh::unfold)")
    lin, cst = edge_xforms(g, opp)
    lin, cst = np.ascontiguousarray(lin), np.ascontiguousarray(cst)
    M, S = scalar_matrices(g, triangles, nv)
    signals = [np.asarray(sig_a, dtype=np.float64).copy(), np.asarray(sig_b, dtype=np.float64).copy()]
    if params.logSpace:  # OpticalFlow.cpp:821 (before the DoG; the colours advected at the end, :482-489, stay raw)
        signals = [np.log(np.maximum(1.0, s)) * 255.0 / np.log(255.0) for s in signals]
    if params.dogWeight > 0:
        dog = [dog_preprocess(M, S, g, triangles, s, params.dogSmooth) for s in signals]
        signals = dog if params.dogWeight >= 1 else [dog_blend(s, d, params.dogWeight) for s, d in zip(signals, dog)]
    reduced, expanded, positive = whitney_numbering(opp)
    if params.vfMode == 0:
        P = whitney_prolongation(g, reduced, positive)
        Sw, _, _ = whitney_smooth_operator(g, triangles, opp, reduced, expanded, positive, nv)
    elif params.vfMode == 1:
        P, Sw = conformal_field(g, triangles, nv, S)
    elif params.vfMode == 2:
        P, Sw = connection_field(g, opp, lin, cst, params.cMode)
    else:
        raise ValueError("ERROR: Unsupported vector field!")
    w = Whitney(reduced, expanded, positive, P, Sw)
    T = triangles.shape[0]
    return State(vertices, triangles, g, opp, lin, cst, triangle_areas(g), M, S, w, signals, np.zeros(P.shape[1]), np.zeros((T, 2)))


def update_flow(st: State, s_weight, vf_weight, tap_prefix=None):
    """UpdateFlow, OpticalFlow.cpp:424-474 (SMOOTH_FIRST)."""
    smoothed = [smooth_signal(st.M, st.S, s, s_weight) if s_weight else s for s in st.signals]
    resampled = [resample_signal(st.triangles, st.opp, st.lin, st.cst, st.g, st.tfield, smoothed[s], -0.5 if s == 0 else 0.5) for s in range(2)]
    D, rhs = data_term(st.triangles, st.area, resampled[0], resampled[1])
    st.coeffs, st.tfield, x = update_optical_flow(st.whitney, st.coeffs, D, rhs, vf_weight)
    st.tfield = np.ascontiguousarray(st.tfield)
    if tap_prefix is not None:
        st.taps.update({tap_prefix + "smoothed0": smoothed[0], tap_prefix + "smoothed1": smoothed[1], tap_prefix + "resampled0": resampled[0],
                        tap_prefix + "resampled1": resampled[1], tap_prefix + "dataTerm": D, tap_prefix + "rhs": rhs, tap_prefix + "x": x,
                        tap_prefix + "coeffs": st.coeffs.copy(), tap_prefix + "tFlowField": st.tfield.copy()})


def iterate(st: State, params: Params, taps=False):
    """IterativeOptimization's loop, OpticalFlow.cpp:1037-1043."""
    sw = params.sSmooth
    vw = params.vfSmooth
    for i in range(params.iterations):
        update_flow(st, sw, vw, "it%02d." % i if taps else None)
        sw *= params.sMultiply
        vw = vw * params.vMultiply if vw * params.vMultiply > params.vfSThreshold else vw


def advect_vertices(st: State, col_a, col_b, alpha=0.5):
    """InputGeometryData::flow, OpticalFlow.cpp:482-489: raw colours along -alpha and 1-alpha."""
    a = resample_signal(st.triangles, st.opp, st.lin, st.cst, st.g, st.tfield, np.asarray(col_a, dtype=np.float64), -alpha)
    b = resample_signal(st.triangles, st.opp, st.lin, st.cst, st.g, st.tfield, np.asarray(col_b, dtype=np.float64), 1.0 - alpha)
    return a, b


def align_vertices(vertices, triangles, col_a, col_b, params: Params | None = None, taps=False):
    """--in A.ply B.ply --out r.ply: returns (State, blended colours [V,3] before the uchar cast)."""
    params = params or Params()
    st = init(vertices, triangles, col_a, col_b, params)
    iterate(st, params, taps)
    a, b = advect_vertices(st, col_a, col_b)
    return st, (a + b) / 2.0


def prepare_texture_mesh(vertices_f32, triangles, tri_uv, e_length_param=float(np.float32(0.006))):
    """The texture branch of Init up to the signals, OpticalFlow.cpp:704-726: bounding-box diagonal, edge
    length = float(eLength * diagonal) (:713), Subdivide (:714)."""
    v = np.ascontiguousarray(vertices_f32, dtype=np.float32)
    lo, hi = v.astype(np.float64).min(0), v.astype(np.float64).max(0)
    diagonal = float(np.sqrt(((hi - lo) ** 2).sum()))
    e_len = float(np.float32(np.float32(e_length_param) * diagonal))
    if e_len > 0:
        return subdivide(v, triangles, np.asarray(tri_uv, dtype=np.float32).astype(np.float64), e_len)
    return v, np.ascontiguousarray(triangles, dtype=np.int32), np.asarray(tri_uv, dtype=np.float64)


def align_texture(vertices_f32, triangles, tri_uv, tex_a, tex_b, params: Params | None = None, taps=False, bilinear=True):
    """--mesh m.ply --in A.png B.png --out r.png (OpticalFlow.cpp:686-751, 818, 1044-1047). Textures are uint8
    [H,W,3], top row first. Returns (State, dict with srcT, srcP, advected [2][W*H,3], pixels uint8 [H,W,3] as written)."""
    params = params or Params()
    v, t, uv = prepare_texture_mesh(vertices_f32, triangles, tri_uv, params.eLength)
    H, W = tex_a.shape[0], tex_a.shape[1]
    sig = [sample_texture_to_vertices(t, uv, v.shape[0], tex, bilinear) for tex in (tex_a, tex_b)]
    st = init(v.astype(np.float64), t, sig[0], sig[1], params)
    srcT, srcP = texture_source(uv, W, H, params.pad, st.opp, st.lin, st.cst, st.g)
    iterate(st, params, taps)
    adv = [advect_texels(W, H, srcT, srcP, st.opp, st.lin, st.cst, st.g, st.tfield, uv, tex, length, bilinear) for tex, length in ((tex_a, -0.5), (tex_b, 0.5))]
    blended = (adv[0] + adv[1]) / 2.0
    pixels = to_uchar_png(blended).reshape(H, W, 3)[::-1]  # flipY, OpticalFlow.cpp:131
    return st, {"vertices": v, "triangles": t, "tri_uv": uv, "srcT": srcT, "srcP": srcP, "advected": adv, "pixels": np.ascontiguousarray(pixels)}


def to_uchar_ply(colors):
    """OutputMesh, OpticalFlow.cpp:139-148: float cast, clamp to [0,255], uchar by truncation
    (PlyFile.inl:2309-2313)."""
    c = np.clip(colors.astype(np.float32), 0.0, 255.0)
    return c.astype(np.uint8)


def to_uchar_png(pixels):
    """OutputImage, OpticalFlow.cpp:126-137: (int) truncation then clamp."""
    return np.clip(np.trunc(pixels), 0, 255).astype(np.uint8)
