"""Ad-hoc (GPU box): candidates per left row under different bounds on the pruned part of the score
(tile-wide Cauchy-Schwarz as shipped, per column, grouped per column, exact).   python tests/gpu_bound_stats.py N [tile_w]"""
import sys
import numpy as np
import pandas as pd
import torch
sys.path.insert(0, '.')
from synth_corpus import make_names
from string_grouper_b200 import _device as D, _ingest

n = int(sys.argv[1]) if len(sys.argv) > 1 else 663_000
tile_w = int(sys.argv[2]) if len(sys.argv) > 2 else 128
thr = 0.8
names = make_names(n, 0)
data, offsets, flags, _ = _ingest.pack_strings([pd.Series(names)])
A, _, _ = D.tfidf(data, offsets, n, 3, flags, np.float64)
hrank, perm_b, rank, bucket_dir, maxw, post, T, tile_bound = D.right_side(A, tile_w)
V = A.shape[1]
dev = A.device
l_idx, l_val, l_len, l_thr, l_xp, _ = D.prune_left(A, A, hrank, 0, n, thr, D.CAND_MARGIN, D.U16_MARGIN_PER_FEATURE, 0.9)
indptr = A.d_indptr[:n + 1]
idx = A.d_indices[:A.nnz].long()
val = A.d_val32[:A.nnz]
Bcsr = torch.sparse_csr_tensor(indptr, idx, val, size=(n, V))
row_of = torch.repeat_interleave(torch.arange(n, device=dev), indptr[1:] - indptr[:-1])
hr = hrank.long()
heavy = hr >= 0
tile_of = (rank.long() // tile_w)
bound_col = tile_bound[tile_of]                      # tile-wide bound seen by every column

def group_norms(gid_of_rank, G):
    g = torch.where(heavy, gid_of_rank(hr.clamp(min=0)), torch.zeros_like(hr))
    e_heavy = heavy[idx]
    Y = torch.zeros(n * G, device=dev)
    Y.index_add_(0, (row_of * G + g[idx])[e_heavy], (val * val)[e_heavy])
    Gm = torch.zeros(V, G, device=dev)
    Gm[torch.arange(V, device=dev)[heavy], g[heavy]] = 1.0
    return Y.view(n, G).sqrt() * (1 + 1e-6), Gm

def singles(k, rest):          # ranks 0..k-1 alone, the others in `rest` equal ranges
    width = -(-(64 - k) // rest)
    return lambda r: torch.where(r < k, r, k + (r - k) // width)

schemes = {"8 by rank range": (lambda r: r // 8, 8), "16 by rank range": (lambda r: r // 4, 16),
           "8 = 6 singles + 2": (singles(6, 2), 8), "8 = 7 singles + 1": (singles(7, 1), 8),
           "8 = 4 singles + 4": (singles(4, 4), 8),
           "16 = 12 singles + 4": (singles(12, 4), 16), "16 = 14 singles + 2": (singles(14, 2), 16),
           "16 = 8 singles + 8": (singles(8, 8), 16), "32 = 24 singles + 8": (singles(24, 8), 32),
           "64 (exact support)": (lambda r: r, 64)}
norms = {k: group_norms(f, G) for k, (f, G) in schemes.items()}
R = 256
tot = {}
rows_seen = 0
for start in range(0, n - R, max((n - R) // 16, 1)):
    rows = perm_b[start:start + R].long()
    Xf = torch.zeros(R, V, device=dev)
    Xk = torch.zeros(R, V, device=dev)
    for i, r in enumerate(rows.tolist()):
        p0, p1 = int(indptr[r]), int(indptr[r + 1])
        Xf[i, idx[p0:p1]] = val[p0:p1]
        k = int(l_len[r])
        Xk[i, l_idx[p0:p0 + k].long()] = l_val[p0:p0 + k]
    Xp = Xf - Xk
    Xp[Xp.abs() < 1e-12] = 0
    S_kept = torch.sparse.mm(Bcsr, Xk.t()).t()          # [R, n]
    S_full = torch.sparse.mm(Bcsr, Xf.t()).t()
    thr_r, xp = l_thr[rows][:, None], l_xp[rows][:, None]
    rows_seen += R
    def add(name, mask):
        tot[name] = tot.get(name, 0) + int(mask.sum())
    add("above threshold (truth)", S_full > thr)
    add("shipped: tile-wide bound", S_kept > (thr_r - xp * bound_col[None, :]).clamp(min=0))
    for name, (Y, Gm) in norms.items():
        Xg = (Xp * Xp @ Gm).sqrt() * (1 + 1e-6)        # [R, G]
        add("per column, " + name, S_kept > (thr_r - Xg @ Y.t()).clamp(min=0))
    add("exact pruned part", S_kept + (S_full - S_kept) > thr_r)
print("n=%d tile_w=%d rows sampled=%d" % (n, tile_w, rows_seen))
for k, v in tot.items():
    print("  %-36s %8.1f per row" % (k, v / rows_seen))
