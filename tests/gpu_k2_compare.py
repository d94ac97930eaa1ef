"""Ad-hoc: tile-centric K2 vs row K2 on the GPU box — times, candidates, walked postings, identical results
(not a test, not the bench).   python tests/gpu_k2_compare.py N [tiles|row|both] [reps]"""
import sys, time
import numpy as np
import pandas as pd
import torch
sys.path.insert(0, '.')
from synth_corpus import make_names
from string_grouper_b200 import _device as D, _ingest

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
which = sys.argv[2] if len(sys.argv) > 2 else "both"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
names = make_names(n, 0)
data, offsets, flags, _ = _ingest.pack_strings([pd.Series(names)])
A, _, _ = D.tfidf(data, offsets, n, 3, flags, np.float64)
print("n=%d nnz=%d V=%d" % (n, A.nnz, A.shape[1]), flush=True)
res = {}
for kernel in (["tiles", "row"] if which == "both" else [which]):
    for rep in range(reps):
        A._order = None; A._postings2 = {}; A._tiles = None
        st = {"time_kernels": True}
        torch.cuda.synchronize(); t = time.time()
        got = D.cossim_topn(A, A, 20, 0.8, stats=st, kernel=kernel)
        torch.cuda.synchronize(); tk = time.time() - t
        kms = sum(a.elapsed_time(b) for a, b in st["candidate_events"])
        print("%s rep %d: cossim_topn %.1f ms, candidates launch(es) %.2f ms, cand=%d above=%d nnz=%d pairs=%s postings=%s stage=%s prune=%s select=%s" % (
            kernel, rep, tk * 1e3, kms, st["n_candidates"], st["n_above_threshold"], got.nnz, st.get("pairs_walked"),
            st.get("postings_walked"), st.get("stage_bytes"), st.get("prune"), st.get("select")), "refined=%s" % st.get("n_refined"), flush=True)
        print("   phases ms:", {k: round(v, 2) for k, v in D.phases_ms(st).items()}, flush=True)
    res[kernel] = got.host_triples()
if len(res) == 2:
    a, b = res["tiles"], res["row"]
    same = all(np.array_equal(x, y) for x, y in zip(a, b))
    print("results identical:", same, len(a[0]), len(b[0]), flush=True)
    if not same:
        ka = set(zip(a[0].tolist(), a[1].tolist())); kb = set(zip(b[0].tolist(), b[1].tolist()))
        print("only tiles:", len(ka - kb), list(ka - kb)[:10], "only row:", len(kb - ka), list(kb - ka)[:10])
        sys.exit(1)
