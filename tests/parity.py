"""Parity checker between the CUDA path and the CPU oracle.

Pairs must be identical and scores equal within `tol`, except "boundary ties":
pairs whose score lies within `tol` of min_similarity, or within `tol` of the
top-n cut-off score of a row that is full — there the reference's own choice
depends on heap / traversal order and on n_blocks.
"""
import numpy as np


def row_cutoffs(indptr, data, top_n, n_rows):
    """score of the last kept entry for rows holding exactly top_n entries, else -inf."""
    cut = np.full(n_rows, -np.inf)
    cnt = np.diff(indptr)
    full = np.nonzero(cnt >= top_n)[0]
    for r in full:
        cut[r] = data[indptr[r]:indptr[r + 1]].min()
    return cut


def compare_triples(ref, got, n_cols, threshold, tol=1e-5, cutoff_row=None, cutoff_col=None, label=""):
    """ref/got: (row, col, score) arrays.  Returns stats; raises AssertionError on a real mismatch."""
    rr, rc, rs = (np.asarray(x) for x in ref)
    gr, gc, gs = (np.asarray(x) for x in got)
    rk = rr.astype(np.int64) * n_cols + rc.astype(np.int64)
    gk = gr.astype(np.int64) * n_cols + gc.astype(np.int64)
    assert len(np.unique(rk)) == len(rk), label + ": duplicate pairs in reference"
    assert len(np.unique(gk)) == len(gk), label + ": duplicate pairs in result"
    ro, go = np.argsort(rk), np.argsort(gk)
    rk, rs, rr, rc = rk[ro], rs[ro], rr[ro], rc[ro]
    gk, gs, gr, gc = gk[go], gs[go], gr[go], gc[go]
    common, ri, gi = np.intersect1d(rk, gk, assume_unique=True, return_indices=True)
    max_err = float(np.abs(rs[ri] - gs[gi]).max()) if len(common) else 0.0
    assert max_err <= tol, "%s: score mismatch %.3g > %.3g" % (label, max_err, tol)

    def exempt(rows, cols, scores):
        ok = np.abs(scores - threshold) <= tol
        if cutoff_row is not None:
            ok |= np.abs(scores - cutoff_row[rows]) <= tol
        if cutoff_col is not None:
            ok |= np.abs(scores - cutoff_col[cols]) <= tol
        return ok

    only_r = np.setdiff1d(np.arange(len(rk)), ri)
    only_g = np.setdiff1d(np.arange(len(gk)), gi)
    ex_r = exempt(rr[only_r], rc[only_r], rs[only_r])
    ex_g = exempt(gr[only_g], gc[only_g], gs[only_g])
    bad_r, bad_g = only_r[~ex_r], only_g[~ex_g]
    assert len(bad_r) == 0 and len(bad_g) == 0, (
        "%s: %d pairs only in reference, %d only in result (first: ref %s got %s)" % (
            label, len(bad_r), len(bad_g),
            [(int(rr[i]), int(rc[i]), float(rs[i])) for i in bad_r[:5]],
            [(int(gr[i]), int(gc[i]), float(gs[i])) for i in bad_g[:5]]))
    return {"pairs_ref": int(len(rk)), "pairs_got": int(len(gk)), "common": int(len(common)),
            "max_abs_err": max_err, "boundary_ties": int(len(only_r) + len(only_g))}


def csr_triples(m):
    m = m.tocsr()
    r = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    return r, m.indices.copy(), m.data.copy()
