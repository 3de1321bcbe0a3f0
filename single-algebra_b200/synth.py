"""Synthetic single-cell count matrices (host side of the generator).

The device generator (csrc/synth.cu, `salg_csr_synth_*`) and this numpy generator produce
bit-identical CSR matrices for the same `SynthSpec`: all per-cell work is 64-bit integer
hashing plus integer comparisons against tables that are built ONCE here (in f64) and
handed to the device verbatim, so any row range of a 1M x 30k matrix generated on a B200
can be regenerated on the host for a spot check.

Model: genes j have log-normal base rates beta_j; cells belong to one of `n_clusters`
planted clusters with per-cluster log-fold changes on a subset of marker genes; each cell
has one of 16 size-factor levels; x_ij ~ Poisson(lambda_ij), sampled by inverse CDF with
lambda quantised to 256 log-spaced levels.  Zeros are not stored.
"""
from __future__ import annotations

import dataclasses

import numpy as np

N_LEVELS = 256
K_CDF = 40          # thresholds per level: x = #{j : u >= cdf[level][j]}  in 0..K_CDF
N_SF = 16
LAM_MIN = 1e-5
LAM_MAX = 16.0

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_COLK = np.uint64(0xD1B54A32D192ED03)
_COLA = np.uint64(0x8CB92BA72F3D8DD7)
_ROWSALT = np.uint64(0xA5A5A5A5DEADBEEF)


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * _M1
        x = x ^ (x >> np.uint64(27))
        x = x * _M2
        x = x ^ (x >> np.uint64(31))
    return x


@dataclasses.dataclass
class SynthSpec:
    nrows: int
    ncols: int
    seed: int
    n_clusters: int
    base_level: np.ndarray      # uint8  [n_clusters, ncols]
    sf_offset: np.ndarray       # int32  [N_SF]   added to the level per cell
    cdf: np.ndarray             # uint32 [N_LEVELS, K_CDF]
    density: float              # expected density realised by the tables


def _poisson_cdf_table():
    lam = LAM_MIN * (LAM_MAX / LAM_MIN) ** (np.arange(N_LEVELS) / (N_LEVELS - 1))
    # P(X <= j) for j = 0..K_CDF-1, as u32 thresholds
    j = np.arange(K_CDF)
    from scipy.stats import poisson
    c = poisson.cdf(j[None, :], lam[:, None])
    t = np.floor(c * 4294967296.0)
    t = np.minimum(t, 4294967295.0).astype(np.uint32)
    return lam, t


def _level_of(lam):
    x = np.log(np.maximum(lam, LAM_MIN) / LAM_MIN) / np.log(LAM_MAX / LAM_MIN) * (N_LEVELS - 1)
    return np.clip(np.rint(x), 0, N_LEVELS - 1).astype(np.int64)


def make_spec(nrows, ncols, density=0.07, seed=42, n_clusters=30, sigma=1.2,
              marker_frac=0.10, lfc_sigma=0.8) -> SynthSpec:
    """Build the generator tables; the lognormal location is bisected so that the expected
    density over (cluster, gene, size-factor level) hits `density`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    z = rng.standard_normal(ncols) * sigma
    lfc = rng.standard_normal((n_clusters, ncols)) * lfc_sigma
    lfc *= rng.random((n_clusters, ncols)) < marker_frac
    # 16 size-factor levels, roughly Gamma(5, 0.2) quantiles, as integer level offsets
    q = (np.arange(N_SF) + 0.5) / N_SF
    from scipy.stats import gamma
    sfac = gamma.ppf(q, a=5.0, scale=0.2)
    step = np.log(LAM_MAX / LAM_MIN) / (N_LEVELS - 1)
    sf_off = np.rint(np.log(sfac) / step).astype(np.int32)
    lam_levels, cdf = _poisson_cdf_table()
    p_nz = 1.0 - cdf[:, 0].astype(np.float64) / 4294967296.0

    def dens(mu):
        lvl = _level_of(np.exp(mu + z[None, :] + lfc))
        tot = 0.0
        for o in sf_off:
            tot += p_nz[np.clip(lvl + o, 0, N_LEVELS - 1)].mean()
        return tot / N_SF

    lo, hi = -14.0, 3.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if dens(mid) < density:
            lo = mid
        else:
            hi = mid
    mu = 0.5 * (lo + hi)
    base_level = _level_of(np.exp(mu + z[None, :] + lfc)).astype(np.uint8)
    return SynthSpec(nrows, ncols, seed, n_clusters, base_level, sf_off, cdf, dens(mu))


def row_meta(spec: SynthSpec, rows):
    """(cluster, size-factor level) of each row — must match synth.cu::row_meta."""
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        hr = mix64((np.uint64(spec.seed) ^ _ROWSALT) + rows * _GOLD)
    cluster = ((hr & np.uint64(0xFFFF)) % np.uint64(spec.n_clusters)).astype(np.int64)
    sf = ((hr >> np.uint64(16)) & np.uint64(N_SF - 1)).astype(np.int64)
    return cluster, sf


def expected_row_nnz(spec: SynthSpec, row0: int, row1: int):
    """Expected stored entries of rows [row0, row1) under the generator's model (f64): a row's (cluster, size-factor)
    pair fixes its per-gene Poisson levels.  Used to cut row shards balanced by entries without
    generating the matrix."""
    p_nz = 1.0 - spec.cdf[:, 0].astype(np.float64) / 4294967296.0
    table = np.empty((spec.n_clusters, N_SF), dtype=np.float64)
    for sf in range(N_SF):
        lvl = np.clip(spec.base_level.astype(np.int64) + int(spec.sf_offset[sf]), 0, N_LEVELS - 1)
        table[:, sf] = p_nz[lvl].sum(axis=1)
    cluster, sf = row_meta(spec, np.arange(row0, row1, dtype=np.uint64))
    return table[cluster, sf]


def generate_rows(spec: SynthSpec, row0: int, row1: int, dtype=np.float32, block=512):
    """CSR (indptr, indices[int64], data[dtype]) of rows [row0, row1) of the spec's matrix."""
    indptr = [0]
    idx_parts, val_parts = [], []
    cols = np.arange(spec.ncols, dtype=np.uint64)
    with np.errstate(over="ignore"):
        colterm = cols * _COLK + _COLA
    total = 0
    for b0 in range(row0, row1, block):
        b1 = min(row1, b0 + block)
        rows = np.arange(b0, b1, dtype=np.uint64)
        cluster, sf = row_meta(spec, rows)
        with np.errstate(over="ignore"):
            hrow = mix64(np.uint64(spec.seed) + rows * _GOLD)
        h = mix64(hrow[:, None] ^ colterm[None, :])
        u = (h >> np.uint64(32)).astype(np.uint32)
        lvl = spec.base_level[cluster].astype(np.int64) + spec.sf_offset[sf][:, None]
        lvl = np.clip(lvl, 0, N_LEVELS - 1)
        c0 = spec.cdf[:, 0][lvl]
        nz = u >= c0
        r_i, c_i = np.nonzero(nz)
        uu = u[r_i, c_i]
        ll = lvl[r_i, c_i]
        x = (uu[:, None] >= spec.cdf[ll]).sum(axis=1)
        cnt = np.bincount(r_i, minlength=b1 - b0)
        for c in cnt:
            total += int(c)
            indptr.append(total)
        idx_parts.append(c_i.astype(np.int64))
        val_parts.append(x.astype(dtype))
    indices = np.concatenate(idx_parts) if idx_parts else np.zeros(0, np.int64)
    data = np.concatenate(val_parts) if val_parts else np.zeros(0, dtype)
    return np.asarray(indptr, dtype=np.int64), indices, data


def generate(spec: SynthSpec, dtype=np.float32):
    return generate_rows(spec, 0, spec.nrows, dtype=dtype)


def make_omega(n_eff, l, seed=42, dtype=np.float32):
    """Host-generated Gaussian test matrix shared by the oracle and the GPU path
    (BASELINE.md section 4): Generator(PCG64(seed)).standard_normal((n_eff, l))."""
    return np.random.Generator(np.random.PCG64(seed)).standard_normal((n_eff, l)).astype(dtype)


def make_mask(ncols, n_keep, seed=7):
    """Sorted random subset of `n_keep` columns as a bool mask."""
    ids = np.random.Generator(np.random.PCG64(seed)).choice(ncols, n_keep, replace=False)
    m = np.zeros(ncols, dtype=bool)
    m[ids] = True
    return m
