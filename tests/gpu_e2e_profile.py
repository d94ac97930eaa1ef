"""Ad-hoc host-side profile of the end-to-end path on the GPU box (not a test, not the bench)."""
import cProfile, pstats, sys, time, io
import numpy as np
import pandas as pd
import torch
sys.path.insert(0, '.')
from synth_corpus import make_names
import string_grouper_b200 as api

import os
n = int(sys.argv[1]) if len(sys.argv) > 1 else 663000
rank = int(os.environ.get("RANK", "0"))
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
series = pd.Series(make_names(n, 0))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sg = api.StringGrouper(series); t1 = time.perf_counter()
    sg.fit(); torch.cuda.synchronize(); t2 = time.perf_counter()
    out = sg.get_matches(); torch.cuda.synchronize(); t3 = time.perf_counter()
    if rank == 0: print("rep %d: init %.3f fit %.3f get_matches %.3f total %.3f rows %d" % (rep, t1 - t0, t2 - t1, t3 - t2, t3 - t0, len(out)), flush=True)
pr = cProfile.Profile()
pr.enable()
sg = api.StringGrouper(series); sg.fit(); out = sg.get_matches(); torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
if rank == 0: print(s.getvalue()[:9000])
