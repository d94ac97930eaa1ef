"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): torchrun with one rank per GPU, NCCL all-gather of the
per-rank row blocks; every rank must return exactly what a single GPU returns."""
import os
import subprocess
import sys

import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_equals_single_gpu(tmp_path):
    import torch
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n_gpu < 4 else 4
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import string_grouper_b200 as api
    from synth_corpus import make_names
    names = pd.Series(make_names(30011, seed=91))
    dupes = pd.Series(make_names(7000, seed=92) + make_names(30011, seed=91)[:999])
    a = api.match_strings(names, min_similarity=0.8)
    b = api.match_strings(names, dupes, min_similarity=0.7, max_n_matches=5)
    g = api.group_similar_strings(names)
    for rank in range(world):
        pd.testing.assert_frame_equal(pd.read_pickle("%s.self.%d.pkl" % (out, rank)), a)
        pd.testing.assert_frame_equal(pd.read_pickle("%s.two.%d.pkl" % (out, rank)), b)
        pd.testing.assert_frame_equal(pd.read_pickle("%s.grp.%d.pkl" % (out, rank)), g)
