#!/usr/bin/env python
"""Benchmark of the string_grouper hot path on B200 — one JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (oracle port), bounded sample

Workload (BASELINE.json metric / configs[2]): match_strings self-match of a 663 000-name sec__edgar-shaped
synthetic corpus (synth_corpus.make_names(seed=0)), 3-grams, min_similarity 0.8, max_n_matches 20, float64.
One step = one pass of the hot path: K1 vectorise -> K2 top-n cosine product -> K4 symmetrise.

  value : matched pairs / s with the packed strings already resident in HBM (CUDA events around K1..K4)
  e2e   : same metric through string_grouper_b200.match_strings(pandas Series) - host buffers in, DataFrame out
  roofline : the dominant kernel (cossim_candidates): algorithmic bytes / CUDA-event time
  cpu_baseline : the oracle port of the reference CPU path on this box's cores, bounded sample, extrapolated
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "match_strings matched pairs/sec (663k names self-match @0.8, top 20, 3-gram)"
UNIT = "pairs/s"
TOP_N, MIN_SIM = 20, 0.8


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, power) if p > 250] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference arm
def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota (os.cpu_count()
    reports the machine, and OpenMP threads beyond the quota only spin against each other)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path, parse in (("/sys/fs/cgroup/cpu.max", lambda txt: txt.split()),
                        ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", None)):
        try:
            if parse:
                q, per = parse(open(path).read())
                if q != "max":
                    n = min(n, max(1, int(float(q) / float(per))))
            else:
                q = int(open(path).read())
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


_THREADS = {}


def best_thread_count(full_matrix):
    """The OpenMP thread count at which the oracle's block product runs fastest on this box (probed once on
    8000 left rows x 48000 right rows; candidates: the usable cores and a few fractions of them)."""
    if "n" in _THREADS:
        return _THREADS["n"], _THREADS["probe"]
    from oracle import pipeline as P
    cores = usable_cores()
    cand = sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32), min(cores, 16), min(cores, 8)})
    nl, nr = min(8000, full_matrix.shape[0]), min(48000, full_matrix.shape[0])
    left, right = full_matrix[:nl], full_matrix[:nr]
    probe = {}
    for c in cand:
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            P.build_matches(left, right, (1, 12), TOP_N, MIN_SIM, c)
            best = min(best, time.perf_counter() - t0)
        probe[c] = round(best, 4)
    _THREADS["n"] = min(probe, key=probe.get)
    _THREADS["probe"] = probe
    return _THREADS["n"], probe


def cpu_reference_sample(names, full_matrix, n_threads, sample_left=(6000, 30000), sample_self=20000):
    """Bounded sample of the reference CPU path (oracle port), extrapolated to the whole job.

    (a) the oracle's fit() on the first `sample_self` names: analyzer + TfidfVectorizer (2 of the reference's 3
        analyzer passes), block product, LIL symmetrise, match list -> per-string and per-match host costs;
    (b) the block product of the first s1 and the first s2 left rows against ALL right rows with the reference's
        own block heuristic (string_grouper.py:387-389): t(s) = fixed + per_row * s.  `fixed` (slicing and
        transposing the right blocks, one OpenMP region per block) is paid once per job, only `per_row` scales
        with the left rows, so the job estimate is fixed + per_row * n (NOT t(s) * n / s).
    Returns (estimated seconds for the full job, estimated pairs, detail dict).
    """
    from oracle import pipeline as P
    n = len(names)
    sample_self = min(sample_self, n)
    t0 = time.perf_counter()
    m, d, _ = P.tf_idf_matrices(names[:sample_self])
    t_vec = time.perf_counter() - t0
    t0 = time.perf_counter()
    C = P.build_matches(m, d, P.guess_blocks(sample_self, sample_self), TOP_N, MIN_SIM, n_threads)
    t_mm_small = time.perf_counter() - t0
    t0 = time.perf_counter()
    S = P.fix_diagonal_and_symmetrize(C)
    ml = P.matches_list(S)
    t_post = time.perf_counter() - t0
    per_string = 1.5 * t_vec / sample_self            # fit + transform measured; the reference also fits in __init__
    per_match = t_post / max(len(ml), 1)
    blocks = (1, P.guess_blocks(n, n)[1])
    s1, s2 = (min(x, n) for x in sample_left)
    times, nnz_s2 = [], 0
    for sl in (s1, s2):
        t0 = time.perf_counter()
        Cs = P.build_matches(full_matrix[:sl], full_matrix, blocks, TOP_N, MIN_SIM, n_threads)
        times.append(time.perf_counter() - t0)
        nnz_s2 = Cs.nnz
    if s2 > s1:
        per_row = max((times[1] - times[0]) / (s2 - s1), 0.0)
        fixed = max(times[0] - per_row * s1, 0.0)
    else:
        per_row, fixed = times[0] / max(s1, 1), 0.0
    t_product = fixed + per_row * n
    est_pairs = (nnz_s2 / s2) * n * (len(ml) / max(C.nnz, 1))      # symmetrisation growth from (a)
    est = per_string * n + t_product + per_match * est_pairs
    detail = {"t_vectorise_sample_s": round(t_vec, 3), "t_product_s1_s2_s": [round(x, 3) for x in times],
              "left_rows_s1_s2": [s1, s2], "product_fixed_s": round(fixed, 3),
              "product_per_left_row_us": round(per_row * 1e6, 3), "est_product_s": round(t_product, 2),
              "t_post_sample_s": round(t_post, 3), "t_product_small_s": round(t_mm_small, 3),
              "est_vectorise_s": round(per_string * n, 2), "est_post_s": round(per_match * est_pairs, 2),
              "est_total_s": round(est, 2), "est_pairs": int(est_pairs), "n_blocks": list(blocks),
              "threads": n_threads}
    return est, est_pairs, detail


def run_reference_arm(args, names):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the reference tree has
    no native code and its Python cannot travel to this box), all host threads, bounded sample per step."""
    from oracle import pipeline as P
    from oracle import sdt
    sdt.build()
    t0 = time.perf_counter()
    full, _, _ = P.tf_idf_matrices(names)           # setup, untimed: the sample needs the full right matrix
    setup = time.perf_counter() - t0
    cores, probe = best_thread_count(full)
    vals, last = [], None
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        est, est_pairs, detail = cpu_reference_sample(names, full, cores)
        wall = time.perf_counter() - t0
        if step >= args.warmup:
            vals.append((est_pairs / est, est, wall))
        last = detail
    value = float(np.mean([v[0] for v in vals]))
    sample = ("per step: oracle fit() on 20000 names + block products of %s left rows x all %d right rows, "
              "n_blocks=%s, job estimate = fixed + per-left-row cost x rows; right matrix built once before "
              "timing (%.0f s); threads = fastest of %s on this box (os.cpu_count() = %s, usable = %d)"
              % (last["left_rows_s1_s2"], len(names), last["n_blocks"], setup, probe, os.cpu_count(),
                 usable_cores()))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([v[1] for v in vals])) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(len(names), args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "detail": last},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample_wall_s_per_step": float(np.mean([v[2] for v in vals]))}
    print(json.dumps(line), flush=True)


def workload_config(n, gpus):
    return {"workload": "match_strings self-match, %d synthetic sec__edgar-shaped names (synth_corpus seed 0), "
                        "ngram 3, min_similarity %.1f, max_n_matches %d, tfidf float64" % (n, MIN_SIM, TOP_N),
            "rows": n, "parallelism": "left-row shards x%d, right matrix replicated" % gpus,
            "l2": "256 MiB memset between steps (untimed); postings + bucket table exceed what stays L2-resident "
                  "across steps"}


# ----------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=663_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"           # keep stdout to the one JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from synth_corpus import corpus_sha256, make_names
    if args.impl == "reference":
        if rank != 0:
            return 0
        run_reference_arm(args, make_names(args.rows, seed=0))
        return 0

    import pandas as pd
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import string_grouper_b200 as api
    from string_grouper_b200 import _device as D
    from string_grouper_b200 import _dist, _ingest, _lib
    _lib.load()

    names = make_names(args.rows, seed=0)
    n = len(names)
    series = pd.Series(names)
    data, offsets, flags, _ = _ingest.pack_strings([series])
    d_bytes, d_off, total = D.upload_strings(data, offsets, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lo, hi = _dist.shard_range(n, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    info = {}

    def resident_step(stats):
        A, _, _ = D.tfidf_resident(d_bytes, d_off, n, total, n, 3, flags, np.float64, stats=stats)
        M = D.cossim_topn(A, A, TOP_N, MIN_SIM, row_begin=lo, row_end=hi, stats=stats)
        info["k2_nnz_local"] = M.nnz
        if world > 1:
            M = D.gather_shards(M)
        S = D.symmetrize(M)
        info["A"] = A
        return S

    # ---- device-resident throughput ("value") ----
    step_ms, cand_ms, pairs = [], [], 0
    launches0 = None
    sampler = ClockSampler(local_rank)
    for step in range(args.warmup + args.steps):
        flush.zero_()
        barrier()
        timed = step >= args.warmup
        if timed and launches0 is None:
            launches0 = dict(D.LAUNCH_COUNTS)
            if rank == 0:
                sampler.start()
        stats = {"time_kernels": True}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S = resident_step(stats)
        e1.record()
        barrier()
        if timed:
            step_ms.append(e0.elapsed_time(e1))
            cand_ms.append(sum(a.elapsed_time(b) for a, b in stats["candidate_events"]))
            pairs = S.nnz
            info["stats"] = stats
    launches = sum(D.LAUNCH_COUNTS.values()) - sum(launches0.values())
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step / 1e3)

    # ---- end to end through the public API ("e2e") ----
    e2e_s, e2e_rows, h2d, d2h = [], 0, 0, 0
    for step in range(args.warmup + args.steps):
        flush.zero_()
        barrier()
        d2h0 = D.TRANSFER_BYTES["d2h"]
        t0 = time.perf_counter()
        sg = api.StringGrouper(series)
        t1 = time.perf_counter()
        sg.fit()
        t2 = time.perf_counter()
        out = sg.get_matches()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            e2e_s.append(dt)
            e2e_parts = {"validate_s": t1 - t0, "fit_s": t2 - t1, "get_matches_s": t0 + dt - t2}
            e2e_rows = len(out)
            h2d = int(sg._last_stats.get("h2d_bytes", 0))
            d2h = int(D.TRANSFER_BYTES["d2h"] - d2h0)      # match list + gathered strings (small read-backs not counted)
    t_e2e = torch.tensor([sum(e2e_s)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_rows / (float(t_e2e.item()) / args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (cossim_candidates), SURVEY.md §8d ----
    A = info["A"]
    idx = A.d_indices[:A.nnz].long()
    df = torch.bincount(idx, minlength=A.shape[1])
    ip = A.d_indptr
    nnz_a_local = int((ip[hi] - ip[lo]).item())
    macs_local = int(df[A.d_indices[int(ip[lo].item()):int(ip[hi].item())].long()].sum().item())
    macs_total = int((df * df).sum().item())
    alg_bytes = 8 * macs_local + 8 * nnz_a_local + 4 * (hi - lo + 1) + 8 * info["k2_nnz_local"] + 4 * (hi - lo + 1)
    k2_ms = float(np.mean(cand_ms))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    achieved = alg_bytes / (k2_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:      # the ncu capture is a 1-GPU launch over all rows
        traffic = json.load(open(tpath)).get("%d" % n)
    # what the pruned traversal really walks (one extra, untimed launch with the counting switched on)
    st_count = {"count_macs": True}
    D.cossim_topn(A, A, TOP_N, MIN_SIM, row_begin=lo, row_end=hi, stats=st_count)
    walked = st_count.get("macs_walked")
    pipe = None
    ppath = os.path.join(ROOT, "profiles", "pipe.json")
    if os.path.exists(ppath) and world == 1:
        pipe = json.load(open(ppath)).get("%d" % n)
    roofline = {"bound": "hbm", "kernel": "sg::cossim_candidates_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "macs_per_launch": macs_local, "macs_total": macs_total,
                "kernel_ms": k2_ms, "kernel_share_of_step": k2_ms / float(np.mean(step_ms)),
                "note": "achieved/frac follow SURVEY.md §8d: bytes of the FULL Gustavson traversal / kernel time; "
                        "exact threshold pruning walks only `walked.macs` postings (4 B each, L2-resident), so frac > 1 "
                        "is an algorithmic gain, not HBM utilisation; the measured limiter is in `pipe` (ncu)",
                "walked": {"macs": walked, "share_of_full": (walked / macs_local) if walked else None,
                           "posting_bytes": 4 * walked if walked else None,
                           "posting_GBps": (4 * walked / (k2_ms / 1e3) / 1e9) if walked else None,
                           "posting_frac_of_peak": (4 * walked / (k2_ms / 1e3) / 1e9 / peak) if walked else None,
                           "prune": st_count.get("prune"), "accumulator": st_count.get("acc")},
                "pipe": pipe,
                "tile": {k: info["stats"].get(k) for k in ("tile_w", "warps", "n_tiles", "n_candidates",
                                                            "n_above_threshold")}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import sdt
        sdt.build()
        full = A.to_scipy()          # parity-checked equal to the sklearn matrix (tests/test_gpu_tfidf.py)
        cores, probe = best_thread_count(full)
        est, est_pairs, detail = cpu_reference_sample(names, full, cores)
        cpu_baseline = {"value": est_pairs / est, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "oracle fit() on 20000 names + block products of %s left rows x all %d right "
                                  "rows (n_blocks=%s); job estimate = fixed + per-left-row cost x rows; threads = "
                                  "fastest of %s (os.cpu_count() = %s, usable = %d)"
                                  % (detail["left_rows_s1_s2"], n, detail["n_blocks"], probe, os.cpu_count(),
                                     usable_cores()),
                        "detail": detail}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(n, world), corpus_sha256=corpus_sha256(names), pairs_per_step=pairs,
                           nnz=A.nnz, vocab=A.shape[1]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "s_per_step": float(np.mean(e2e_s)), "s_each_step": [round(x, 4) for x in e2e_s],
                    "rows": e2e_rows, "last_step_parts": e2e_parts},
            "gpu_launches": launches // args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
