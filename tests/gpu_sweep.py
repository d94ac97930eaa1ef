"""Ad-hoc sweep of the K2 candidate kernel on the GPU box (not a test, not the bench):
python tests/gpu_sweep.py ROWS "prune:acc:tile_w:warps" ...   -> one line per configuration."""
import sys, time
import numpy as np
import pandas as pd
import torch
sys.path.insert(0, '.')
from synth_corpus import make_names
from string_grouper_b200 import _device as D, _ingest

n = int(sys.argv[1])
cfgs = [c.split(":") for c in sys.argv[2:]] or [["0.7", "u16", "0", "32"]]
names = make_names(n, 0)
data, offsets, flags, _ = _ingest.pack_strings([pd.Series(names)])
A, _, _ = D.tfidf(data, offsets, n, 3, flags, np.float64)
torch.cuda.synchronize()
ref = None
for cfg in cfgs:
    prune, acc, tile_w, warps = cfg[:4]
    thr = float(cfg[4]) if len(cfg) > 4 else 0.8
    for rep in range(2):
        st = {"time_kernels": True, "count_macs": rep == 0}
        torch.cuda.synchronize(); t = time.time()
        got = D.cossim_topn(A, A, 20, thr, tile_w=int(tile_w) or None, warps=int(warps), stats=st, prune=float(prune), acc=acc)
        torch.cuda.synchronize(); tk = time.time() - t
        if rep == 0:
            macs = st.get("macs_walked")
    evs = st["candidate_events"]
    kms = [a.elapsed_time(b) for a, b in evs]
    trip = got.host_triples()
    chk = (int(trip[0].astype(np.int64).sum()), int(trip[1].astype(np.int64).sum()), float(trip[2].sum()))
    if ref is None or thr != 0.8:
        ref = chk
    print("thr=%s prune=%s acc=%s tile_w=%d warps=%d: topn %.1f ms, kernel %s ms, cand=%d (est %s) nnz=%d walked_macs=%s same=%s" % (
        thr, prune, st["acc"], st["tile_w"], st["warps"], tk * 1e3, ["%.1f" % k for k in kms], st["n_candidates"],
        st.get("n_candidates_estimate"), got.nnz, macs, chk == ref), flush=True)
