"""Generates tests/golden/pca_small.npz: a small planted-cluster count matrix, the host Gaussian test
matrix and the oracle's fit results (f64), so the GPU parity tests also run against committed vectors.
The reference itself (Rust, un-vendored single-svdlib) cannot be run in this image; the fixture is the
oracle's output — see oracle/oracle.py header ("parity unpinned")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import planted_counts  # noqa: E402
from oracle import oracle as O  # noqa: E402

A = planted_counts(1500, 300, density=0.06, n_clusters=8, seed=123)
k, p, q = 20, 10, 7
omega = np.random.Generator(np.random.PCG64(42)).standard_normal((300, k + p))
r = O.sparse_pca_fit(A, k, omega=omega, n_oversamples=p, n_power_iterations=q)
mask = np.zeros(300, bool)
mask[np.random.Generator(np.random.PCG64(7)).choice(300, 120, replace=False)] = True
omega_m = np.random.Generator(np.random.PCG64(42)).standard_normal((120, k + p))
rm = O.sparse_pca_fit(A, k, omega=omega_m, n_oversamples=p, n_power_iterations=q, mask=mask)
np.savez_compressed(
    os.path.join(HERE, "pca_small.npz"),
    shape=np.array(A.shape), indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int64), data=A.data,
    k=k, p=p, q=q, omega=omega, singular_values=r.singular_values, components=r.components,
    explained_variance=r.explained_variance, mean=r.mean, total_var=r.total_var,
    scores=O.transform(A, r.components, r.mean, mode=O.EXACT),
    mask=mask, omega_m=omega_m, singular_values_m=rm.singular_values, components_m=rm.components,
    total_var_m=rm.total_var)
print("written", A.shape, A.nnz)
