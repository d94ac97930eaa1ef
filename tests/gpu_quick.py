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
# (unused, tile_w, warps, unused)
cfgs = [(1, 768, 32, 0), (1, 896, 32, 0), (1, 1536, 32, 0), (1, 768, 32, 0)]
if len(sys.argv) > 5:
    cfgs = [(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))]
for algo, tile_w, warps, rows in cfgs:
    st = {"time_kernels": True}
    torch.cuda.synchronize(); t = time.time()
    got = D.cossim_topn(A, A, 20, 0.8, tile_w=tile_w, warps=warps, stats=st)
    torch.cuda.synchronize(); tk = time.time() - t
    kms = sum(a.elapsed_time(b) for a, b in st["candidate_events"])
    print("algo=%d tile_w=%d warps=%d rows=%d: cossim_topn %.1f ms (candidates kernel %.1f ms), cand=%d nnz=%d -> kernel %.3g MAC/s, %.0f GB/s algorithmic" % (
        algo, tile_w, st["warps"], rows, tk * 1e3, kms, st["n_candidates"], got.nnz, macs / (kms / 1e3), 8 * macs / (kms / 1e3) / 1e9))
t = time.time(); sym = D.symmetrize(got); torch.cuda.synchronize(); print("symmetrize %.1f ms nnz=%d" % ((time.time() - t) * 1e3, sym.nnz))
