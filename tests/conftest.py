import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "OpticalFlow_ref")
CLI_BIN = os.path.join(ROOT, "meshopticalflow_b200", "OpticalFlow")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def colour_outliers(a, b, tol=1.0):
    """Fraction of values further apart than tol. Colours are compared, not walk paths: a sample point that
    sits exactly on an edge can be carried to either side by last-bit differences."""
    return float((np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) > tol).mean())


@pytest.fixture(scope="session")
def golden_sphere():
    return dict(np.load(os.path.join(GOLDEN, "sphere3_vertex.npz")))


@pytest.fixture(scope="session")
def golden_torus():
    return dict(np.load(os.path.join(GOLDEN, "torus_texture.npz")))


@pytest.fixture(scope="session")
def golden_modes():
    return dict(np.load(os.path.join(GOLDEN, "sphere3_modes.npz")))


# name in tests/golden/sphere3_modes.npz -> (vfMode, cMode)
VF_MODES = {"conformal": (1, 0), "connection0": (2, 0), "connection1": (2, 1), "connection2": (2, 2)}


def csr_from_golden(g, name, shape=None):
    import scipy.sparse as sp
    return sp.csr_matrix((g[name + ".val"], g[name + ".col"], g[name + ".rowptr"]), shape=shape)


def sorted_csr(m):
    m = m.tocsr().copy()
    m.sort_indices()
    return m
