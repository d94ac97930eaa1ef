"""Ad-hoc timing of the K2 stages on the GPU box (not a test, not the bench)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from synth_corpus import make_names
from oracle import pipeline as P
from string_grouper_b200 import _device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
names = make_names(n, 0)
t = time.time(); m, d, _ = P.tf_idf_matrices(names, dtype=np.float64); print("oracle tfidf %.2fs nnz=%d V=%d" % (time.time() - t, m.nnz, m.shape[1]))
macs = P.hot_path_macs(m, m); print("MACs %.4g  (%.3f per pair)" % (macs, macs / n / n))
A = D.DeviceCSR.from_scipy(m)
cfgs = [(1536, 32), (3072, 16), (768, 32), (1024, 32), (2048, 24), (1536, 32)]
if len(sys.argv) > 3:
    cfgs = [(int(sys.argv[2]), int(sys.argv[3]))]
for tile_w, warps in cfgs:
    A._postings.clear()
    torch.cuda.synchronize(); t = time.time()
    A.postings(tile_w); torch.cuda.synchronize(); tp = time.time() - t
    st = {}
    t = time.time(); got = D.cossim_topn(A, A, 20, 0.8, tile_w=tile_w, warps=warps, stats=st); torch.cuda.synchronize(); tk = time.time() - t
    print("tile_w=%d warps=%d: postings %.1f ms, cossim_topn %.1f ms, cand=%d nnz=%d  -> %.3g MAC/s, %.1f GB/s algorithmic" % (
        tile_w, warps, tp * 1e3, tk * 1e3, st["n_candidates"], got.nnz, macs / tk, 8 * macs / tk / 1e9))
t = time.time(); sym = D.symmetrize(got); torch.cuda.synchronize(); print("symmetrize %.1f ms nnz=%d" % ((time.time() - t) * 1e3, sym.nnz))
