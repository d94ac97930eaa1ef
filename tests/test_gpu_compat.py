"""-m gpu: the operator-level boundary (string_grouper_b200/sparse_dot_topn_compat.py) — CUDA versions of the two
callables the reference imports at string_grouper.py:12 — against the CPU oracle of the same callables, and inside the
reference's own block loop (oracle/pipeline.build_matches restates string_grouper.py:709-752 verbatim)."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix, random as sprandom

from parity import compare_triples, csr_triples, row_cutoffs
from synth_corpus import make_names

pytestmark = pytest.mark.gpu


def _rows_as_sets(m):
    m = m.tocsr()
    return [dict(zip(m.indices[m.indptr[i]:m.indptr[i + 1]].tolist(), m.data[m.indptr[i]:m.indptr[i + 1]].tolist()))
            for i in range(m.shape[0])]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("threshold,sort", [(0.3, True), (0.3, False), (0.0, True), (None, False), (0.75, True)])
def test_sp_matmul_topn_equals_oracle(dtype, threshold, sort):
    from oracle import pipeline as P
    from oracle import sdt
    from string_grouper_b200 import sparse_dot_topn_compat as C
    left = make_names(700, seed=31)
    right = make_names(400, seed=32) + left[:150]
    m, d, _ = P.tf_idf_matrices(left, right, dtype=dtype)
    ref = sdt.sp_matmul_topn(m, d.T, top_n=7, threshold=threshold, sort=sort, n_threads=2)
    got = C.sp_matmul_topn(m, d.T, top_n=7, threshold=threshold, sort=sort, n_threads=2)
    assert got.shape == ref.shape and got.dtype == ref.dtype and got.indices.dtype == np.int32
    thr = -1.0 if threshold is None else threshold
    cut = row_cutoffs(ref.indptr, ref.data, 7, m.shape[0])
    tol = 1e-5 if dtype == np.float32 else 1e-12
    compare_triples(csr_triples(ref), csr_triples(got), d.shape[0], thr, tol=max(tol, 1e-5), cutoff_row=cut,
                    label="sp_matmul_topn")
    if sort:      # rows value-descending
        for i in range(got.shape[0]):
            v = got.data[got.indptr[i]:got.indptr[i + 1]]
            assert np.all(v[1:] <= v[:-1])


def test_sp_matmul_topn_argument_errors():
    from string_grouper_b200 import sparse_dot_topn_compat as C
    a = csr_matrix(np.eye(3))
    with pytest.raises(TypeError):
        C.sp_matmul_topn(a, np.eye(3), 2)
    with pytest.raises(ValueError):
        C.sp_matmul_topn(a, csr_matrix(np.eye(4)), 2)
    with pytest.raises(TypeError):
        C.sp_matmul_topn(a, csr_matrix(np.eye(3, dtype=np.float32)), 2)
    neg = csr_matrix(np.array([[1.0, -1.0], [0.5, 0.5]]))
    with pytest.raises(NotImplementedError):
        C.sp_matmul_topn(neg, neg.T.tocsr(), 2, threshold=None)
    out = C.sp_matmul_topn(a, a.T.tocsr(), 5, threshold=0.5)          # top_n larger than the right side
    np.testing.assert_array_equal(out.toarray(), np.eye(3))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_zip_equals_oracle_and_global_topn(dtype):
    from oracle import sdt
    from string_grouper_b200 import sparse_dot_topn_compat as C
    rng = np.random.default_rng(4)
    blocks = []
    for w in (50, 1, 33, 120):
        b = sprandom(200, w, density=0.2, format="csr", random_state=rng, dtype=np.float64)
        b.data = np.round(b.data, 2)              # many exact ties, some exact zeros
        blocks.append(b.astype(dtype))
    for top_n in (1, 4, 300):
        ref = sdt.zip_sp_matmul_topn(top_n, blocks)
        got = C.zip_sp_matmul_topn(top_n, blocks)
        assert got.shape == ref.shape == (200, 204) and got.dtype == ref.dtype
        cut = row_cutoffs(ref.indptr, ref.data.astype(np.float64), top_n, 200)
        compare_triples(csr_triples(ref), csr_triples(got), 204, 0.0, tol=1e-9, cutoff_row=cut, label="zip")
        assert got.nnz == ref.nnz
        for i in range(200):
            v = got.data[got.indptr[i]:got.indptr[i + 1]]
            assert np.all(v[1:] <= v[:-1]) and np.all(v > 0)


@pytest.mark.parametrize("n_blocks", [None, (1, 1), (1, 4), (2, 3), (3, 7)])
def test_reference_block_loop_with_cuda_operators(n_blocks, monkeypatch):
    """string_grouper.py:709-752 restated verbatim (oracle/pipeline.build_matches) with the two operators swapped
    for the CUDA ones: same matrix as with the CPU operators for every block split."""
    from oracle import pipeline as P
    from string_grouper_b200 import sparse_dot_topn_compat as C
    names = make_names(2500, seed=41)
    dupes = make_names(900, seed=42) + names[:300]
    m, d, _ = P.tf_idf_matrices(names, dupes)
    ref = P.build_matches(m, d, n_blocks, 6, 0.6, n_threads=2)
    monkeypatch.setattr(P, "sp_matmul_topn", C.sp_matmul_topn)
    monkeypatch.setattr(P, "zip_sp_matmul_topn", C.zip_sp_matmul_topn)
    got = P.build_matches(m, d, n_blocks, 6, 0.6, n_threads=2)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    cut = row_cutoffs(ref.indptr, ref.data, 6, len(names))
    st = compare_triples(csr_triples(ref), csr_triples(got), len(dupes), 0.6, tol=1e-12, cutoff_row=cut,
                         label="block loop %r" % (n_blocks,))
    assert st["common"] > 500
