"""Worker for tests/test_multi_gpu.py: run under torchrun, one rank per GPU (NCCL)."""
import os
import sys

os.environ["SG_B200_DISTRIBUTED"] = "1"           # sharding is opt-in
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pandas as pd
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import string_grouper_b200 as api
from synth_corpus import make_names

out = sys.argv[1]
names = pd.Series(make_names(30011, seed=91))
dupes = pd.Series(make_names(7000, seed=92) + make_names(30011, seed=91)[:999])
a = api.match_strings(names, min_similarity=0.8)
b = api.match_strings(names, dupes, min_similarity=0.7, max_n_matches=5)
g = api.group_similar_strings(names)
os.environ["SG_B200_SHARD_VECTORISE"] = "1"      # sharded K1: df all-reduce + all-gather of the duplicate matrix
sg = api.StringGrouper(names, dupes, min_similarity=0.7, max_n_matches=5).fit()
assert sg._last_stats.get("sharded_vectorise")
b2 = sg.get_matches()
m2 = api.match_most_similar(names, dupes, min_similarity=0.7)
os.environ["SG_B200_SHARD_VECTORISE"] = "0"
m1 = api.match_most_similar(names, dupes, min_similarity=0.7)
pd.testing.assert_frame_equal(b, b2)
pd.testing.assert_frame_equal(m1, m2) if isinstance(m1, pd.DataFrame) else pd.testing.assert_series_equal(m1, m2)
a.to_pickle("%s.self.%d.pkl" % (out, rank))
b.to_pickle("%s.two.%d.pkl" % (out, rank))
g.to_pickle("%s.grp.%d.pkl" % (out, rank))
dist.barrier()
dist.destroy_process_group()
