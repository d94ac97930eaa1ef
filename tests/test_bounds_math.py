"""not-gpu: the two inequalities the pruned K2 traversal relies on (csrc/sg_prune.cu, csrc/sg_cossim.cu), restated in
numpy on the oracle's TF-IDF matrix.  The CUDA kernels are checked end to end by the -m gpu tests (every pruning
level returns bit-identical results); this file pins the MATH, so that a change of the rules that breaks exactness
fails on CPU already.

  (1) pruning:    x.y <= |x_P| * |y_H| + x_S.y         for P a subset of the heavy features H
  (2) block max:  x_S.y_j <= sum_f |x_f| * max_{j' in tile(j)} |y_j'f|
  (3) grouped:    x_P.y <= sum_g |x_P,g| * |y_H,g|     over the 16 groups of heavy ranks (sg_rescore_refined)
"""
import numpy as np
import scipy.sparse as sp

from oracle import pipeline as P
from synth_corpus import make_names

N_HEAVY, TILE_W, THRESHOLD, BUDGET = 64, 256, 0.8, 0.9 * (0.8 - 1.5e-3)


def _setup(n=6000, seed=5):
    A, _, _ = P.tf_idf_matrices(make_names(n, seed=seed))
    A = A.tocsr()
    V = A.shape[1]
    df = np.bincount(A.indices, minlength=V)
    heavy = np.zeros(V, bool)
    heavy[np.argsort(-df, kind="stable")[:N_HEAVY]] = True
    return A, df, heavy


def _prune(A, df, heavy):
    """Host statement of prune_rows_kernel: per row, heavy features ranked by df / w^2, pruned while the norm of
    the pruned part stays within the budget.  Returns (kept matrix, norm of the pruned part per row)."""
    n = A.shape[0]
    keep = np.ones(A.nnz, bool)
    xp = np.zeros(n)
    for i in range(n):
        lo, hi = A.indptr[i], A.indptr[i + 1]
        f, w = A.indices[lo:hi], A.data[lo:hi]
        key = np.where(heavy[f], df[f] / (w * w), 0.0)
        order = np.lexsort((np.arange(hi - lo), -key))
        acc = 0.0
        for k in order:
            if key[k] <= 0 or acc + w[k] ** 2 > BUDGET ** 2:
                break
            acc += w[k] ** 2
            keep[lo + k] = False
        xp[i] = np.sqrt(acc)
    S = sp.csr_matrix((A.data * keep, A.indices.copy(), A.indptr.copy()), shape=A.shape)
    S.eliminate_zeros()
    return S, xp


def test_pruning_and_block_max_bounds_never_lose_a_match():
    A, df, heavy = _setup()
    n = A.shape[0]
    S, xp = _prune(A, df, heavy)
    assert S.nnz < 0.9 * A.nnz and xp.max() <= BUDGET + 1e-12
    AH = sp.csr_matrix((A.data * heavy[A.indices], A.indices.copy(), A.indptr.copy()), shape=A.shape)
    y_heavy = np.sqrt(np.asarray(AH.multiply(AH).sum(axis=1)).ravel())
    # right rows in processing order: quantised heavy norm first (the signature part does not matter here)
    perm = np.argsort(np.minimum(31, np.ceil(y_heavy * 31)), kind="stable")
    tile_of = np.empty(n, dtype=np.int64)
    tile_of[perm] = np.arange(n) // TILE_W
    T = int(tile_of.max()) + 1
    tile_bound = np.zeros(T)
    np.maximum.at(tile_bound, tile_of, y_heavy)
    coo = A.tocoo()
    maxw = np.zeros((T, A.shape[1]))
    np.maximum.at(maxw, (tile_of[coo.row], coo.col), np.abs(coo.data))

    full = (A @ A.T).tocoo()
    part = (S @ A.T).tocsr()
    hit = full.data > THRESHOLD
    rows, cols, score = full.row[hit], full.col[hit], full.data[hit]
    partial = np.asarray(part[rows, cols]).ravel()
    # (1) every true match is reported by the pruned traversal: partial score above thr - |x_P| * tile bound
    assert np.all(partial + xp[rows] * tile_bound[tile_of[cols]] >= score - 1e-12)
    assert np.all(partial > THRESHOLD - xp[rows] * tile_bound[tile_of[cols]] - 1e-12)
    # (2) and its tile survives the block-max test
    Sabs = abs(S).tocsr()
    ub_hit = np.array([Sabs[r].multiply(maxw[tile_of[c]]).sum() for r, c in zip(rows[:4000], cols[:4000])])
    assert np.all(ub_hit >= partial[:4000] - 1e-12)
    # the test has teeth: most (row, tile) pairs are skippable
    sample = np.arange(0, n, 40)
    ub_all = (Sabs[sample] @ sp.csr_matrix(maxw).T).toarray()
    thr = THRESHOLD - 1.5e-3 - xp[sample][:, None] * tile_bound[None, :]
    assert (ub_all <= thr).mean() > 0.5


def _half_up(x):
    """fp16 rounded towards +inf (what __float2half_ru does) as float64."""
    h = np.asarray(x, dtype=np.float64).astype(np.float16)
    low = h.astype(np.float64) < x
    h = np.where(low, np.nextafter(h, np.float16(np.inf)), h)
    return h.astype(np.float64)


def _norm_up(s2):
    """norm_up() of csrc/sg_prune.cu in fp32: sqrt(s) * (1 + 1e-5) + 1e-6, 0 stays 0."""
    s2 = np.asarray(s2, dtype=np.float32)
    out = np.sqrt(s2) * np.float32(1.0 + 1e-5) + np.float32(1e-6)
    return np.where(s2 > 0, out, np.float32(0)).astype(np.float64)


def _heavy_group(rank):
    return np.where(rank < 14, rank, np.where(rank < 39, 14, 15))


def test_grouped_bound_of_the_pruned_part_is_an_upper_bound_and_tight():
    """(3) sg_rescore_refined:  x_P . y  <=  sum_g |x_P,g| |y_H,g|  over the 16 groups of heavy ranks (0..13 alone,
    14..38, 39..63), with the group norms computed in fp32 from the fp32 weights, rounded up to fp16, and the sum taken
    in fp32 with the kernel's slack.  Never below the exact fp64 pruned part; tighter than |x_P| |y_H|."""
    A, df, heavy = _setup()
    n, V = A.shape
    S, xp = _prune(A, df, heavy)
    order = np.argsort(-df, kind="stable")[:N_HEAVY]
    rank = np.full(V, -1)
    rank[order] = np.arange(N_HEAVY)
    grp = _heavy_group(rank)
    G = 16
    A32 = A.astype(np.float32)                       # the kernels see the fp32 copy of the weights

    def group_norms(M):                              # [rows, G] fp16-rounded-up norms over the heavy entries of M
        coo = M.tocoo()
        h = rank[coo.col] >= 0
        s2 = np.zeros((M.shape[0], G), dtype=np.float32)
        np.add.at(s2, (coo.row[h], grp[coo.col[h]]), (coo.data[h].astype(np.float32)) ** 2)
        return _half_up(_norm_up(s2))

    P32 = (A32 - S.astype(np.float32)).tocsr()       # pruned part, fp32
    P32.eliminate_zeros()
    assert P32.nnz > 1000 and np.all(rank[P32.indices] >= 0)
    xg, yg = group_norms(P32), group_norms(A32)
    P64 = (A - S).tocsr()
    rows = np.flatnonzero(xp > 0)[:1500]
    exact = (P64[rows] @ A.T).toarray()              # exact pruned part of every (row, column) pair, fp64
    bound = (xg[rows].astype(np.float32) @ yg.T.astype(np.float32)).astype(np.float64) * (1 + 1e-5) + 1e-6
    assert np.all(bound >= exact)
    AH = sp.csr_matrix((A.data * heavy[A.indices], A.indices.copy(), A.indptr.copy()), shape=A.shape)
    y_heavy = np.sqrt(np.asarray(AH.multiply(AH).sum(axis=1)).ravel())
    single = xp[rows][:, None] * y_heavy[None, :]
    live = exact > 0.02
    assert live.sum() > 1000
    # the slack of the single-group Cauchy-Schwarz bound over the exact value more than halves (6000 rows; at 663k
    # rows the candidates per row go from 458 to 45, tests/gpu_bound_stats.py)
    assert (bound[live] - exact[live]).mean() < 0.5 * (single[live] - exact[live]).mean()
