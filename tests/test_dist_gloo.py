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
    os.environ["SG_B200_DISTRIBUTED"] = "1"       # sharding is opt-in
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


WORKER_GUARDS = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import pandas as pd, torch.distributed as dist
    rank = int(sys.argv[1])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2)
    import string_grouper_b200 as api
    from string_grouper_b200 import _dist
    from cpu_backend import oracle_device
    # 1. an initialised default group alone does not shard: rank-local data stays rank-local
    assert _dist.world() == (0, 1)
    local = pd.Series(["foo inc %%d" %% rank, "foo inc. %%d" %% rank, "bar llc"])
    with oracle_device():
        out = api.match_strings(local, min_similarity=0.5)
    assert set(out.left_side) <= set(local)
    # 2. opted in, but the ranks hold different Series: every rank raises instead of mixing match lists
    _dist.enable(True)
    assert _dist.world() == (rank, 2)
    try:
        with oracle_device():
            api.match_strings(local, min_similarity=0.5)
        raise SystemExit("expected ValueError")
    except ValueError as e:
        assert "different input" in str(e)
    # 3. a rank failing in its rank-local part takes the other one down before the collective
    def work():
        if rank == 1:
            raise OverflowError("boom")
        return 7
    try:
        _dist.guarded(work)
        raise SystemExit("expected an exception")
    except OverflowError:
        assert rank == 1
    except RuntimeError as e:
        assert rank == 0 and "another rank failed" in str(e)
    # 4. SG_B200_RESULT=rank0: only rank 0 ends up with the match list
    os.environ["SG_B200_RESULT"] = "rank0"
    same = pd.Series(["foo inc", "foo inc.", "bar llc", "bar l.l.c"])
    with oracle_device():
        out = api.match_strings(same, min_similarity=0.5)
    assert (len(out) > 0) == (rank == 0), (rank, len(out))
    dist.destroy_process_group()
''')


def test_sharding_is_opt_in_and_guarded(tmp_path):
    port = 31500 + os.getpid() % 2000
    script = tmp_path / "worker_guards.py"
    script.write_text(WORKER_GUARDS % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
