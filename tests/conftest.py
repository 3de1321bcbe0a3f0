import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def planted_counts(nrows, ncols, density=0.06, n_clusters=8, seed=0, dtype=np.float64):
    """Small planted-cluster Poisson count matrix (scipy CSR) for parity tests: decaying spectrum, so
    the leading subspace is well conditioned (flat spectra are poor parity inputs)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    base = np.exp(rng.normal(np.log(density), 1.0, ncols))
    lfc = rng.normal(0, 1.0, (n_clusters, ncols)) * (rng.random((n_clusters, ncols)) < 0.2)
    cl = rng.integers(0, n_clusters, nrows)
    sf = rng.gamma(5.0, 0.2, nrows)
    lam = base[None, :] * np.exp(lfc[cl]) * sf[:, None]
    X = rng.poisson(lam).astype(dtype)
    A = sp.csr_matrix(X)
    A.sort_indices()
    return A


@pytest.fixture(scope="session")
def salg():
    import single_algebra_b200 as s
    return s


@pytest.fixture(scope="session")
def ctx(salg):
    if salg.device_count() == 0:
        pytest.skip("no CUDA device")
    return salg.default_context()
