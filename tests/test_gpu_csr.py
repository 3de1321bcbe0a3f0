"""CSR container, mask compaction (bit-exact) and the device generator — through the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_upload_download_roundtrip_bit_exact(salg, ctx, dtype):
    A = planted_counts(257, 93, seed=1, dtype=dtype)
    m = salg.CsrMatrix.from_scipy(A, ctx)
    off, idx, val = m.to_device().download()
    assert np.array_equal(off, A.indptr.astype(np.uint64))
    assert np.array_equal(idx, A.indices.astype(np.uint64))
    assert np.array_equal(val, A.data)
    # scipy / AnnData layout: int32 indices + int64 offsets
    m2 = salg.CsrMatrix(A.shape[0], A.shape[1], A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data, ctx)
    off2, idx2, val2 = m2.to_device().download()
    assert np.array_equal(off2, off) and np.array_equal(idx2, idx) and np.array_equal(val2, val)


def test_empty_and_ragged_matrices(salg, ctx):
    E = sp.csr_matrix((5, 7), dtype=np.float64)
    d = salg.CsrMatrix.from_scipy(E, ctx).to_device()
    assert (d.nrows, d.ncols, d.nnz) == (5, 7, 0)
    assert d.sum_col().tolist() == [0.0] * 7 and d.sum_row().tolist() == [0.0] * 5
    R = sp.csr_matrix(np.array([[0, 0, 0.0], [1, 2, 3], [0, 0, 0], [0, 4, 0]]))
    d = salg.CsrMatrix.from_scipy(R, ctx).to_device()
    assert d.sum_row().tolist() == [0.0, 6.0, 0.0, 4.0]
    assert d.sum_col().tolist() == [1.0, 6.0, 3.0]


def test_invalid_csr_is_rejected(salg, ctx):
    off = np.array([0, 2, 3], np.uint64)
    val = np.array([1.0, 2.0, 3.0])
    with pytest.raises(salg.SalgError) as e:      # unsorted row
        salg.CsrMatrix(2, 4, off, np.array([2, 1, 0], np.uint64), val, ctx).to_device()
    assert e.value.code == 1 and "strictly increasing" in str(e.value)
    with pytest.raises(salg.SalgError) as e:      # duplicate column
        salg.CsrMatrix(2, 4, off, np.array([1, 1, 0], np.uint64), val, ctx).to_device()
    assert e.value.code == 1
    with pytest.raises(salg.SalgError) as e:      # column out of range
        salg.CsrMatrix(2, 4, off, np.array([0, 4, 0], np.uint64), val, ctx).to_device()
    assert e.value.code == 1 and "out of range" in str(e.value)
    with pytest.raises(salg.SalgError) as e:      # offsets not ending at nnz
        salg.CsrMatrix(2, 4, np.array([0, 2, 2], np.uint64), np.array([0, 1, 2], np.uint64), val, ctx).to_device()
    assert e.value.code == 1 and "offsets" in str(e.value)
    with pytest.raises(salg.SalgError):           # decreasing offsets
        salg.CsrMatrix(2, 4, np.array([0, 3, 2], np.uint64), np.array([0, 1, 2], np.uint64), val, ctx).to_device()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_select_columns_bit_exact(salg, ctx, dtype):
    A = planted_counts(1200, 700, seed=3, dtype=dtype)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    for frac, seed in ((0.07, 1), (0.5, 2), (0.0, 3), (1.0, 4)):
        mask = np.random.default_rng(seed).random(700) < frac
        c = d.select_columns(mask)
        off, idx, val = c.download()
        ip, ix, dv = O.mask_compact(A.indptr, A.indices, A.data, mask)
        assert c.ncols == int(mask.sum())
        assert np.array_equal(off.astype(np.int64), ip)
        assert np.array_equal(idx.astype(np.int64), ix)
        assert np.array_equal(val, dv)


def test_select_columns_mask_length_error(salg, ctx):
    A = planted_counts(50, 30, seed=4)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    with pytest.raises(salg.SalgError) as e:
        d.select_columns(np.ones(29, bool))
    assert e.value.code == 2
    assert str(e.value) == "The mask vector length and the number of features (columns) have to be the same!"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_generator_equals_host_generator(salg, ctx, dtype):
    spec = salg.synth.make_spec(5000, 1500, density=0.07, seed=42)
    for row0, n in ((0, 300), (4700, 300)):
        d = salg.synth_device(spec, row0, n, dtype=dtype, ctx=ctx)
        off, idx, val = d.download()
        ip, ix, dv = salg.synth.generate_rows(spec, row0, row0 + n, dtype=dtype)
        assert np.array_equal(off.astype(np.int64), ip)
        assert np.array_equal(idx.astype(np.int64), ix)
        assert np.array_equal(val, dv)
    full = salg.synth_device(spec, 0, 5000, ctx=ctx)
    assert abs(full.nnz / (5000 * 1500) - 0.07) < 0.01
