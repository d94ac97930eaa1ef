"""Ad-hoc (GPU box): how many super-tiles (S consecutive column tiles, block maximum = max of the S) survive the
block-max test of the candidates kernel, against the tiles themselves.   python tests/gpu_supertile_stats.py N [tile_w]"""
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
V1 = A.shape[1] + 1
Tp = maxw.numel() // V1
maxw = maxw.view(V1, Tp)[:, :T]
margin = D.CAND_MARGIN
l_idx, l_val, l_len, l_thr, l_xp, _ = D.prune_left(A, A, hrank, 0, n, thr, margin, D.U16_MARGIN_PER_FEATURE, 0.9)
indptr = A.d_indptr
tb = tile_bound[:T]
R = 512
tot = {}
rows_seen = 0
long_rows = 0
for start in range(0, n - R, max((n - R) // 12, 1)):
    rows = perm_b[start:start + R].long()
    nf = l_len[rows].long()
    long_rows += int((nf > 32).sum())
    K = 32
    pos = indptr[rows][:, None] + torch.arange(K, device=rows.device)[None, :]
    ok = torch.arange(K, device=rows.device)[None, :] < nf[:, None]
    pos = torch.where(ok, pos, torch.zeros_like(pos))
    f = torch.where(ok, l_idx[pos].long(), torch.full_like(pos, V1 - 1))
    w = torch.where(ok, l_val[pos], torch.zeros_like(pos, dtype=torch.float32))
    thr_r, xp = l_thr[rows], l_xp[rows]
    slack = 5e-4 * nf.clamp(max=32).float() + 1e-4
    keep = nf <= 32
    def survive(mw, bound):
        ub = torch.zeros(R, mw.shape[1], device=rows.device)
        for k in range(K):
            ub += w[:, k:k + 1] * mw[f[:, k]].float()
        thr_t = (thr_r[:, None] - xp[:, None] * bound[None, :]).clamp(min=0)
        return (ub + slack[:, None] > thr_t) & keep[:, None]
    fine = survive(maxw, tb)
    rows_seen += int(keep.sum())
    tot.setdefault("tiles", [0, 0])
    tot["tiles"][0] += int(fine.sum()); tot["tiles"][1] += int(keep.sum()) * T
    for S in (4, 8, 16, 32, 64):
        Ts = -(-T // S)
        pad = Ts * S - T
        mw = torch.nn.functional.pad(maxw.float(), (0, pad)).view(V1, Ts, S).amax(2)
        bd = torch.nn.functional.pad(tb, (0, pad)).view(Ts, S).amax(1)
        coarse = survive(mw, bd)
        truth = torch.nn.functional.pad(fine, (0, pad)).view(R, Ts, S).any(2)
        assert bool((coarse | ~truth).all())
        e = tot.setdefault(S, [0, 0, 0])
        e[0] += int(coarse.sum()); e[1] += int(truth.sum()); e[2] += int(keep.sum()) * Ts
print("n=%d tile_w=%d T=%d rows sampled=%d (rows with > 32 kept features: %d)" % (n, tile_w, T, rows_seen, long_rows))
a, b = tot["tiles"]
print("tiles surviving: %.4f (%.1f per row)" % (a / b, a / rows_seen))
for S in (4, 8, 16, 32, 64):
    c, tr, al = tot[S]
    Ts = -(-T // S)
    # warp steps of 64 bounds each: coarse over all super-tiles + fine over the tiles of the surviving ones
    steps_now = T / 64
    steps_two = Ts / 64 + (c / rows_seen) * S / 64
    print("S=%2d: super-tiles surviving %.4f (holding a surviving tile: %.4f); bound steps per row %.1f -> %.1f" % (
        S, c / al, tr / al, steps_now, steps_two))
