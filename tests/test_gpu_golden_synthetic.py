"""-m gpu: the seeded synthetic goldens the UNMODIFIED reference produced (oracle/make_golden.py: synthetic()) against
the CUDA path — match lists, group representatives (reference _deduplicate) and nearest masters
(reference _get_nearest_matches)."""
import os

import numpy as np
import pandas as pd
import pytest

from parity import compare_triples, row_cutoffs
from synth_corpus import make_names

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "synthetic.npz"), allow_pickle=True)


def test_self3000_match_list_and_group_representatives():
    from string_grouper_b200 import StringGrouper
    names = pd.Series(make_names(3000, seed=11), name="name")
    sg = StringGrouper(names, min_similarity=0.8).fit()
    ml = sg._matches_list
    assert np.array_equal(ml.master_side.to_numpy(), GOLD["self3000_master_side"])
    assert np.array_equal(ml.dupe_side.to_numpy(), GOLD["self3000_dupe_side"])
    np.testing.assert_allclose(ml.similarity.to_numpy(), GOLD["self3000_similarity"], rtol=0, atol=1e-12)
    assert sg._true_max_n_matches == int(GOLD["self3000_true_max"][0])
    assert ml.master_side.dtype == np.int64 and ml.dupe_side.dtype == np.int64 and ml.similarity.dtype == np.float64
    grp = sg.get_groups()
    assert np.array_equal(grp["group_rep_index"].to_numpy(), GOLD["self3000_group_rep_index"])
    # group_rep='first' against the oracle's restatement of _deduplicate on the reference's own match list
    from oracle import pipeline as P
    ref_ml = pd.DataFrame({"master_side": GOLD["self3000_master_side"], "dupe_side": GOLD["self3000_dupe_side"],
                           "similarity": GOLD["self3000_similarity"]})
    first = StringGrouper(names, min_similarity=0.8, group_rep="first").fit().get_groups()
    assert np.array_equal(first["group_rep_index"].to_numpy(), P.deduplicate(ref_ml, len(names), "first"))
    assert np.array_equal(GOLD["self3000_group_rep_index"], P.deduplicate(ref_ml, len(names), "centroid"))


def test_two2000_match_list_and_nearest_masters():
    from string_grouper_b200 import StringGrouper
    master = pd.Series(make_names(2000, seed=12))
    dupes = pd.Series(make_names(1200, seed=12)[:600] + make_names(200, seed=13))
    sg = StringGrouper(master, dupes, min_similarity=0.7, max_n_matches=5).fit()
    ml = sg._matches_list
    ref = (GOLD["two2000_master_side"], GOLD["two2000_dupe_side"], GOLD["two2000_similarity"])
    indptr = np.zeros(len(master) + 1, dtype=np.int64)
    np.cumsum(np.bincount(ref[0], minlength=len(master)), out=indptr[1:])
    cut = row_cutoffs(indptr, ref[2], 5, len(master))          # the reference list is stored row by row
    st = compare_triples(ref, (ml.master_side.to_numpy(), ml.dupe_side.to_numpy(), ml.similarity.to_numpy()),
                         len(dupes), 0.7, tol=1e-12, cutoff_row=cut, label="two2000")
    assert st["common"] >= 0.99 * st["pairs_ref"]
    near = sg.get_groups()
    got = near["most_similar_master"].tolist()
    want = GOLD["two2000_nearest"].tolist()
    same = sum(a == b for a, b in zip(got, want))
    assert same >= len(want) - st["boundary_ties"], (same, len(want), st)
    gi = near["most_similar_index"].to_numpy(dtype=np.float64)
    wi = GOLD["two2000_nearest_index"]
    agree = (gi == wi) | (np.isnan(gi) & np.isnan(wi))
    assert agree.sum() >= len(wi) - st["boundary_ties"]
