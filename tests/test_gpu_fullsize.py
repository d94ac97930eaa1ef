"""-m gpu: BASELINE.json's full-size configuration (663 000 names, self-match @0.8, top 20) checked through
size-independent properties plus an oracle spot check on sampled rows."""
import numpy as np
import pandas as pd
import pytest

from synth_corpus import make_names

pytestmark = pytest.mark.gpu
N = 663_000


@pytest.fixture(scope="module")
def fitted():
    from string_grouper_b200 import StringGrouper
    names = pd.Series(make_names(N, seed=0))
    sg = StringGrouper(names).fit()
    return names, sg


def test_full_size_structure(fitted):
    names, sg = fitted
    ml = sg._matches_list
    r, c, s = ml.master_side.to_numpy(), ml.dupe_side.to_numpy(), ml.similarity.to_numpy()
    key = r * N + c
    assert np.all(np.diff(key) > 0)                               # (row, col) ascending, no duplicates
    assert np.all(s > 0.8) and np.all(s <= 1.0 + 1e-9)            # strict threshold
    diag = r == c
    assert diag.sum() == N and np.all(s[diag] == 1.0)             # _fix_diagonal: every row, exactly 1
    rev = np.searchsorted(key, c * N + r)                         # _symmetrize_matrix: (c, r) present, same score
    assert np.array_equal(key[rev], c * N + r)
    assert np.array_equal(s[rev][~diag], s[~diag])
    assert sg._true_max_n_matches == 20                           # clusters of identical names fill top-n
    assert len(ml) > 4_000_000


def test_full_size_is_tile_and_order_invariant(fitted):
    """sum of checksums over two different tilings (block invariance, reference tests :191-336)."""
    from string_grouper_b200 import _device as D
    names, sg = fitted
    A, _ = sg._get_tf_idf_matrices()
    a = D.cossim_topn(A, A, 20, 0.8, tile_w=768, warps=32)
    b = D.cossim_topn(A, A, 20, 0.8, tile_w=1536, warps=16)
    ta, tb = a.host_triples(), b.host_triples()
    assert a.nnz == b.nnz
    for x, y in zip(ta, tb):
        assert np.array_equal(x, y)


def test_full_size_sampled_rows_equal_oracle(fitted):
    from oracle.sdt import sp_matmul_topn
    from parity import compare_triples, row_cutoffs
    from string_grouper_b200 import _device as D
    names, sg = fitted
    A, _ = sg._get_tf_idf_matrices()
    full = A.to_scipy()
    rng = np.random.default_rng(7)
    rows = np.sort(rng.choice(N, size=1500, replace=False))
    ref = sp_matmul_topn(full[rows], full.T.tocsr(), top_n=20, threshold=0.8, sort=True, n_threads=16)
    got = D.cossim_topn(A, A, 20, 0.8)
    gr, gc, gs = got.host_triples()
    sel = np.isin(gr, rows)
    pos = np.searchsorted(rows, gr[sel])
    cut = row_cutoffs(ref.indptr, ref.data, 20, len(rows))
    rr = np.repeat(np.arange(len(rows)), np.diff(ref.indptr))
    st = compare_triples((rr, ref.indices, ref.data), (pos, gc[sel], gs[sel]), N, 0.8, tol=1e-9, cutoff_row=cut,
                         label="663k sample")
    assert st["common"] >= 0.97 * st["pairs_ref"]


def test_two_series_large_sampled_rows_equal_oracle():
    """config-4-shaped run (master x duplicates, min_similarity 0.7) at 400k x 150k."""
    from oracle.sdt import sp_matmul_topn
    from parity import compare_triples, row_cutoffs
    from string_grouper_b200 import StringGrouper
    base = make_names(480_000, seed=3)
    master = pd.Series(base[:400_000])
    dupes = pd.Series(base[380_000:] + make_names(50_000, seed=4))          # 20k shared + 80k + 50k new
    sg = StringGrouper(master, dupes, min_similarity=0.7).fit()
    ml = sg._matches_list
    r, c, s = ml.master_side.to_numpy(), ml.dupe_side.to_numpy(), ml.similarity.to_numpy()
    assert np.all(np.diff(r) >= 0) and np.all(s > 0.7)
    same = r[1:] == r[:-1]
    assert np.all(s[1:][same] <= s[:-1][same])                       # sort=True inside a row
    assert np.bincount(r).max() <= 20 and c.max() < len(dupes)
    A, B = sg._get_tf_idf_matrices()
    fa, fb = A.to_scipy(), B.to_scipy()
    rows = np.sort(np.random.default_rng(1).choice(len(master), size=1500, replace=False))
    ref = sp_matmul_topn(fa[rows], fb.T.tocsr(), top_n=20, threshold=0.7, sort=True, n_threads=16)
    sel = np.isin(r, rows)
    pos = np.searchsorted(rows, r[sel])
    cut = row_cutoffs(ref.indptr, ref.data, 20, len(rows))
    rr = np.repeat(np.arange(len(rows)), np.diff(ref.indptr))
    st = compare_triples((rr, ref.indices, ref.data), (pos, c[sel], s[sel]), len(dupes), 0.7, tol=1e-9,
                         cutoff_row=cut, label="400k x 150k sample")
    assert st["common"] >= 0.97 * st["pairs_ref"] and st["pairs_ref"] > 500
