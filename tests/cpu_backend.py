"""TEST INFRASTRUCTURE: an oracle-backed stand-in for string_grouper_b200._device so that the HOST logic
(ingest, option handling, result shaping) can be exercised without a GPU (`-m "not gpu"`).

Never imported by the product.  The -m gpu tests run the same cases through the real CUDA path.
"""
import contextlib

import numpy as np
from scipy.sparse import csr_matrix
from sklearn.feature_extraction.text import TfidfVectorizer

from oracle import pipeline as P
from oracle.sdt import sp_matmul_topn
from string_grouper_b200 import _device, _lib


class FakeCSR:
    def __init__(self, m):
        self.m = m.tocsr()
        self.shape = self.m.shape
        self.dtype = self.m.dtype
        self.nnz = self.m.nnz

    def toarray(self):
        return self.m.toarray()

    def to_scipy(self):
        return self.m


class FakeMatches:
    def __init__(self, m, max_row=None, out_dtype=np.float64):
        self.m = m.tocsr()
        self.shape = self.m.shape
        self.nnz = self.m.nnz
        self.max_row = int(np.diff(self.m.indptr).max()) if max_row is None and self.m.shape[0] else int(max_row or 0)
        self.out_dtype = np.dtype(out_dtype)
        self.pending_fix_diagonal = False
        self.pending_mirror = False

    def with_pending(self, fix_diagonal=False, mirror=False):
        self.pending_fix_diagonal |= fix_diagonal
        self.pending_mirror |= mirror
        return self

    def host_triples(self):
        r = np.repeat(np.arange(self.shape[0]), np.diff(self.m.indptr))
        return r, self.m.indices.copy(), self.m.data.astype(np.float64)

    def toarray(self):
        return self.m.astype(self.out_dtype).toarray()


def _device_analyzer(s, ngram, flags):
    """Python model of csrc/sg_tfidf.cu for ASCII input."""
    if flags & _lib.SG_FLAG_IGNORE_CASE:
        s = "".join(chr(ord(c) | 0x20) if "A" <= c <= "Z" else c for c in s)
    if flags & _lib.SG_FLAG_STRIP_DEFAULT:
        s = "".join(c for c in s if not (0x2c <= ord(c) <= 0x2f or 0x09 <= ord(c) <= 0x0d or 0x1c <= ord(c) <= 0x20))
    return [s[i:i + ngram] for i in range(len(s) - ngram + 1)]


def tfidf(data, offsets, n_master, ngram, flags, dtype, device=None, stats=None):
    if np.asarray(data).dtype == np.uint32:         # code points of text that keeps non-ASCII characters (flags == 0)
        docs = [np.asarray(data[offsets[i]:offsets[i + 1]], dtype=np.uint32).tobytes().decode("utf-32-le", "surrogatepass")
                for i in range(len(offsets) - 1)]
    else:
        raw = bytes(np.asarray(data, dtype=np.uint8))
        docs = [raw[offsets[i]:offsets[i + 1]].decode("ascii") for i in range(len(offsets) - 1)]
    vec = TfidfVectorizer(min_df=1, analyzer=lambda s: _device_analyzer(s, ngram, flags), dtype=dtype)
    if docs and any(len(_device_analyzer(d, ngram, flags)) for d in docs):
        vec.fit(docs)
        m = vec.transform(docs)
    else:
        m = csr_matrix((len(docs), 0), dtype=dtype)
    master = FakeCSR(m[:n_master])
    dup = FakeCSR(m[n_master:]) if n_master < len(docs) else None
    return master, dup, None


def as_device_csr(m):
    return m if isinstance(m, FakeCSR) else FakeCSR(m)


def cossim_topn(A, B, top_n, threshold, row_begin=0, row_end=None, **kw):
    if A.shape[1] == 0:
        return FakeMatches(csr_matrix((A.shape[0], B.shape[0])), 0)
    row_end = A.shape[0] if row_end is None else row_end
    C = sp_matmul_topn(A.m[row_begin:row_end], B.m.T, top_n=top_n, threshold=threshold, sort=True, n_threads=1)
    full = csr_matrix((C.data, C.indices, np.concatenate([np.zeros(row_begin, C.indptr.dtype), C.indptr,
                                                          np.full(A.shape[0] - row_end, C.indptr[-1], C.indptr.dtype)])),
                      shape=(A.shape[0], B.shape[0]))
    return FakeMatches(full, max_row=int(np.diff(C.indptr).max()) if C.shape[0] else 0)


def gather_shards(m):
    import torch
    from string_grouper_b200 import _dist
    r, c, s = m.host_triples()
    row, col, score, nnz, max_row = _dist.gather_matches(
        m.shape, torch.from_numpy(r.astype(np.int32)), torch.from_numpy(c.astype(np.int32)),
        torch.from_numpy(s.astype(np.float64)), len(r), m.max_row)
    indptr = np.zeros(m.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(row.numpy(), minlength=m.shape[0]), out=indptr[1:])
    return FakeMatches(csr_matrix((score.numpy(), col.numpy(), indptr), shape=m.shape), max_row=max_row)


def as_device_matches(m):
    return m if isinstance(m, FakeMatches) else FakeMatches(m, out_dtype=m.dtype)


def symmetrize(m, fix_diagonal=True, mirror=True):
    m.pending_fix_diagonal, m.pending_mirror = fix_diagonal, mirror
    return apply_pending(m)


def apply_pending(m):
    lil = m.m.astype(np.float64).tolil()
    if m.pending_fix_diagonal:
        r = np.arange(lil.shape[0])
        lil[r, r] = 1
    if m.pending_mirror:
        r, c = lil.nonzero()
        lil[c, r] = lil[r, c]
    return FakeMatches(lil.tocsr(), m.max_row)


def rowwise_dot(A, B):
    return np.asarray(A.m.multiply(B.m).sum(axis=1)).squeeze(axis=1)


@contextlib.contextmanager
def oracle_device():
    names = ["tfidf", "as_device_csr", "cossim_topn", "as_device_matches", "apply_pending", "rowwise_dot",
             "DeviceMatches", "symmetrize", "gather_shards"]
    saved = {n: getattr(_device, n) for n in names}
    try:
        _device.tfidf, _device.as_device_csr, _device.cossim_topn = tfidf, as_device_csr, cossim_topn
        _device.as_device_matches, _device.apply_pending, _device.rowwise_dot = as_device_matches, apply_pending, rowwise_dot
        _device.DeviceMatches = FakeMatches
        _device.symmetrize = symmetrize
        _device.gather_shards = gather_shards
        yield
    finally:
        for n, v in saved.items():
            setattr(_device, n, v)
