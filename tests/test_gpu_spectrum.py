"""GPU tier: the Spectrum tool — mof_spectrum (csrc/spectrum.cu: LOBPCG on the device operators, a multigrid cycle as
preconditioner for the Whitney basis) and the `Spectrum` command line against ComputeSpectrum as the checker restates it
(include/Src/VectorLaplacianSpectrum.inl:5-39 through scipy's ARPACK shift-invert driver, the same Lanczos the reference calls).

Gates: eigenvalues to 1e-7 of the largest; prolonged eigenvectors compared as SPANS over whole clusters (a sphere's eigenvalues are
multiple, so single vectors are not defined): both sets are orthonormal in <f, h> = sum_t f_t^T (g_t area_t) h_t, so the cross-Gram
matrix of a whole cluster is orthogonal — asserted to 1e-5."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import CLI_BIN
from meshopticalflow_b200 import api, synthetic
from oracle import mof_oracle as O

pytestmark = pytest.mark.gpu
SPECTRUM_BIN = os.path.join(os.path.dirname(CLI_BIN), "Spectrum")


@pytest.fixture()
def aligner():
    al = api.Aligner(0)
    yield al
    al.close()


def cross_gram(v, t, fa, fb):
    g = O.make_unit_area(O.metric_from_embedding(v, t))
    area = O.triangle_areas(g)

    def weighted(f):
        return np.stack([(g[:, 0] * f[:, 0] + g[:, 1] * f[:, 1]) * area, (g[:, 1] * f[:, 0] + g[:, 2] * f[:, 1]) * area], 1)

    return np.array([[np.sum(fa[i] * weighted(fb[j])) for j in range(len(fb))] for i in range(len(fa))])


@pytest.mark.parametrize("mode,cmode,level,count,cluster", [(0, 0, 5, 12, 6), (0, 0, 6, 6, 6), (2, 0, 4, 12, 12), (2, 1, 4, 6, 6), (2, 2, 4, 6, 6)])
def test_spectrum_matches_the_checker(aligner, mode, cmode, level, count, cluster):
    v, t = synthetic.octahedron_sphere(level)
    al = aligner
    p = api.default_params()
    p.vfMode, p.cMode = mode, cmode
    al.set_params(p)
    al.set_mesh(v, t)
    ev, fields, its, res = al.spectrum(count, 1e-8, 20000)
    ref_ev, ref_fields, _, _, _ = O.spectrum(v, t, count, mode, cmode)
    assert res <= 1e-8 and its > 0
    assert np.abs(ev - ref_ev).max() <= 1e-7 * np.abs(ref_ev).max(), (ev, ref_ev)
    G = cross_gram(v, t, fields[:cluster], ref_fields[:cluster])
    assert np.abs(G @ G.T - np.eye(cluster)).max() < 1e-5
    # an alignment on the same context afterwards: mof_spectrum borrowed the data term's buffers and the flow hierarchy
    if mode == 0 and level == 5:
        a, b = (x.astype(np.float64) for x in synthetic.smooth_rgb_pair(v, 1))
        with pytest.raises(api.MofError):
            al.iterate(1)  # signals have to be set again
        p.vfMode = 0
        al.set_signals(a, b)
        al.iterate(2)
        st, _ = O.align_vertices(v, t, a, b, O.Params(iterations=2))
        assert np.linalg.norm(al.flow() - st.tfield) <= 1e-3 * np.linalg.norm(st.tfield)


@pytest.mark.parametrize("level,count", [(3, 6), (6, 6)])
def test_conformal_spectrum_on_the_complement_of_the_constants(aligner, level, count):
    """Conformal basis: constants of either potential are in the null space of S and of M, where ARPACK's shift-invert answer is
    not defined; the checker works on the quotient (oracle.mof_oracle.spectrum). 258 vertices: inverse-diagonal preconditioner;
    16 386 vertices: the two-cycle preconditioner over the scalar hierarchy (without it the iteration does not converge there)."""
    v, t = synthetic.octahedron_sphere(level)
    al = aligner
    p = api.default_params()
    p.vfMode = 1
    al.set_params(p)
    al.set_mesh(v, t)
    ev, fields, its, res = al.spectrum(count, 1e-8, 3000)
    ref_ev, ref_fields, _, _, _ = O.spectrum(v, t, count, 1, 0)
    assert res <= 1e-8 and np.abs(ev - ref_ev).max() <= 1e-7 * ref_ev[-1], (ev, ref_ev)
    G = cross_gram(v, t, fields, ref_fields)
    assert np.abs(G @ G.T - np.eye(count)).max() < 1e-5


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


def test_spectrum_of_a_renumbered_mesh(aligner):
    """A shuffled 65 538-vertex sphere is renumbered inside mof_set_mesh (reorder.cu): same eigenvalues, fields back in the caller's
    triangle order (the span of the first cluster equals that of the sorted mesh's)."""
    v, t = synthetic.octahedron_sphere(7)
    rng = np.random.default_rng(3)
    vo, to = rng.permutation(v.shape[0]), rng.permutation(t.shape[0])
    rank = np.empty_like(vo)
    rank[vo] = np.arange(vo.size)
    vs, ts = np.ascontiguousarray(v[vo]), np.ascontiguousarray(rank[t][to].astype(np.int32))
    al = aligner
    al.set_mesh(v, t)
    ev0, f0, _, _ = al.spectrum(6, 1e-8, 5000)
    al.set_mesh(vs, ts)
    assert al.permutation()[0]
    ev1, f1, _, _ = al.spectrum(6, 1e-8, 5000)
    assert np.abs(ev1 - ev0).max() <= 1e-8 * ev0.max()
    back = np.empty_like(f1)
    back[:, to] = f1
    G = cross_gram(v, t, back, f0)
    assert np.abs(G @ G.T - np.eye(6)).max() < 1e-5


def test_spectrum_command_line(tmp_path):
    """`Spectrum --mesh m.ply --eigenVectors 6`: eigenvector-%03d.bin files in the working directory (Spectrum.cpp:185-189), int count
    then count x 2 doubles (WriteVector, Src/VectorIO.h:23-31), and the reference's eigenvalue listing on stdout."""
    v, t = synthetic.octahedron_sphere(4)
    mesh = str(tmp_path / "m.ply")
    synthetic.write_ply_colored(mesh, v, np.zeros((v.shape[0], 3), np.uint8), t)
    r = subprocess.run([SPECTRUM_BIN, "--mesh", mesh, "--eigenVectors", "6"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    vf = v.astype(np.float32).astype(np.float64)  # what the tool read from the file
    ref_ev, ref_fields, _, _, _ = O.spectrum(vf, t, 6, 0, 0)
    listed = [float(x) for x in r.stdout.split("Eigenvalues:")[1].split()]
    assert np.abs(np.array(listed) - ref_ev).max() < 2e-8 * ref_ev.max() + 1e-8
    fields = []
    for i in range(6):
        raw = (tmp_path / ("eigenvector-%03d.bin" % (i + 1))).read_bytes()
        n = struct.unpack("<i", raw[:4])[0]
        assert n == t.shape[0] and len(raw) == 4 + 16 * n
        fields.append(np.frombuffer(raw[4:], dtype="<f8").reshape(n, 2))
    G = cross_gram(vf, t, np.stack(fields), ref_fields)
    assert np.abs(G @ G.T - np.eye(6)).max() < 1e-5
    r = subprocess.run([SPECTRUM_BIN], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "--mesh" in r.stdout
