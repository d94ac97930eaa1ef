"""not-gpu: the N>1 host logic (row sharding, variable-length all-gather, rank-order concatenation) with
world_size=2 over gloo on CPU; the per-shard product is the CPU oracle standing in for the device."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from string_grouper_b200 import _dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_follow_define_chunks():
    # define_chunks(length, n): ceil(length / n)-sized consecutive ranges (reference string_grouper.py:714-722)
    for n, w in [(10, 3), (7, 8), (0, 2), (663000, 8), (5, 1)]:
        got = [_dist.shard_range(n, r, w) for r in range(w)]
        chunk = int(np.ceil(n / w)) if n else 0
        want = [(min(i * chunk, n), min((i + 1) * chunk, n)) for i in range(w)]
        assert got == want
        assert sum(hi - lo for lo, hi in got) == n


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import numpy as np, pandas as pd, torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
    import string_grouper_b200 as api
    from cpu_backend import oracle_device
    from synth_corpus import make_names
    names = pd.Series(make_names(1501, seed=77))
    dupes = pd.Series(make_names(400, seed=78) + make_names(1501, seed=77)[:99])
    with oracle_device():
        a = api.match_strings(names, min_similarity=0.7)
        b = api.match_strings(names, dupes, min_similarity=0.6, max_n_matches=4)
        g = api.group_similar_strings(names, min_similarity=0.7)
    a.to_pickle("%(out)s.self.%%s.pkl" %% sys.argv[1]); b.to_pickle("%(out)s.two.%%s.pkl" %% sys.argv[1])
    g.to_pickle("%(out)s.grp.%%s.pkl" %% sys.argv[1])
    # row-sharded CSR all-gather and df all-reduce used by the sharded vectoriser
    import torch, scipy.sparse as sp
    from string_grouper_b200 import _dist
    rank = int(sys.argv[1])
    full = sp.random(37, 11, density=0.3, format="csr", random_state=5, dtype=np.float64)
    lo, hi = _dist.shard_range(37, rank, 2)
    part = full[lo:hi]
    indptr, idx, (val,) = _dist.allgather_csr_rows(torch.from_numpy(np.diff(part.indptr).astype(np.int64)),
                                                   torch.from_numpy(part.indices.astype(np.int32)),
                                                   (torch.from_numpy(part.data.copy()),))
    assert np.array_equal(indptr.numpy(), full.indptr) and np.array_equal(idx.numpy(), full.indices)
    assert np.array_equal(val.numpy(), full.data)
    df = torch.full((8,), rank + 1, dtype=torch.int32)
    assert _dist.allreduce_sum_(df).tolist() == [3] * 8
    dist.destroy_process_group()
''')


def test_two_rank_gloo_run_equals_single_process(tmp_path):
    import pandas as pd
    import string_grouper_b200 as api
    from cpu_backend import oracle_device
    from synth_corpus import make_names
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "res")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port, "out": out})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    names = pd.Series(make_names(1501, seed=77))
    dupes = pd.Series(make_names(400, seed=78) + make_names(1501, seed=77)[:99])
    with oracle_device():
        a = api.match_strings(names, min_similarity=0.7)
        b = api.match_strings(names, dupes, min_similarity=0.6, max_n_matches=4)
        g = api.group_similar_strings(names, min_similarity=0.7)
    for r in range(2):
        pd.testing.assert_frame_equal(pd.read_pickle("%s.self.%d.pkl" % (out, r)), a)
        pd.testing.assert_frame_equal(pd.read_pickle("%s.two.%d.pkl" % (out, r)), b)
        pd.testing.assert_frame_equal(pd.read_pickle("%s.grp.%d.pkl" % (out, r)), g)
