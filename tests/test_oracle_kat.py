"""The oracle against the reference's own known-answer tests (tests/golden/kat.json)."""
import json
import os

import numpy as np
import scipy.sparse as sp

from oracle import oracle as O

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def _csr_from_triplets(k):
    A = sp.coo_matrix((k["vals"], (k["rows"], k["cols"])), shape=(k["nrows"], k["ncols"])).tocsr()
    A.sort_indices()
    return A


def test_kat_n1_csr_normalize_column_and_row():
    k = KAT["KAT-N1"]
    A = _csr_from_triplets(k)
    out = O.normalize(A.indptr, A.indices, A.data, np.array(k["column"]["sums"]), k["column"]["target"], O.COLUMN)
    assert np.max(np.abs(out - np.array(k["column"]["expected"]))) < k["tol"]
    out = O.normalize(A.indptr, A.indices, A.data, np.array(k["row"]["sums"]), k["row"]["target"], O.ROW)
    assert np.max(np.abs(out - np.array(k["row"]["expected"]))) < k["tol"]


def test_kat_s1_sums_exact():
    k = KAT["KAT-S1"]
    A = sp.csr_matrix(np.array(k["dense"]))
    assert O.sum_col(A.indptr, A.indices, A.data, 3).tolist() == k["col_sums"]
    assert O.sum_row(A.indptr, A.indices, A.data, 3).tolist() == k["row_sums"]
    assert O.sum_col_squared(A.indptr, A.indices, A.data, 3).tolist() == [17.0, 9.0, 29.0]


def test_kat_n2_normalize_sums_to_target():
    k = KAT["KAT-N2"]
    A = sp.csr_matrix(np.array(k["dense"]))
    out = O.normalize(A.indptr, A.indices, A.data, np.array(k["col_sums"]), k["target"], O.COLUMN)
    B = sp.csr_matrix((out, A.indices, A.indptr), shape=A.shape)
    assert np.max(np.abs(np.asarray(B.sum(axis=0)).ravel() - k["target"])) < k["tol"]
    out = O.normalize(A.indptr, A.indices, A.data, np.array(k["row_sums"]), k["target"], O.ROW)
    B = sp.csr_matrix((out, A.indices, A.indptr), shape=A.shape)
    assert np.max(np.abs(np.asarray(B.sum(axis=1)).ravel() - k["target"])) < k["tol"]


def test_kat_l1_log1p_of_zero():
    k = KAT["KAT-L1"]
    out = O.log1p_normalize(np.array(k["vals"]))
    assert np.max(np.abs(out)) < k["tol"]


def test_normalize_leaves_nonpositive_sums_untouched():
    # src/sparse/csr.rs:1024-1029, 1041, 1055
    A = sp.csr_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))
    out = O.normalize(A.indptr, A.indices, A.data, np.array([0.0, 4.0]), 1.0, O.ROW)
    assert out.tolist() == [1.0, 2.0, 0.75, 1.0]
    out = O.normalize(A.indptr, A.indices, A.data, np.array([-1.0, 2.0]), 1.0, O.COLUMN)
    assert out.tolist() == [1.0, 1.0, 3.0, 2.0]


def test_log1p_is_ln_of_rounded_sum_not_log1p():
    # src/sparse/csr.rs:1074-1075
    x = np.array([1e-10], dtype=np.float32)
    assert O.log1p_normalize(x)[0] == np.log(np.float32(1) + x)[0] == 0.0
    assert np.log1p(x)[0] != 0.0


def test_empty_and_ragged():
    A = sp.csr_matrix((4, 5), dtype=np.float64)
    assert O.sum_col(A.indptr, A.indices, A.data, 5).tolist() == [0.0] * 5
    assert O.sum_row(A.indptr, A.indices, A.data, 4).tolist() == [0.0] * 4
    A = sp.csr_matrix(np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [0, 4, 0.0]]))
    assert O.sum_row(A.indptr, A.indices, A.data, 4).tolist() == [0.0, 6.0, 0.0, 4.0]
