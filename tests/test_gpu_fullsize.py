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
