"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py rep  gpurun_out/prof.ncu-rep            > profiles/<name>.txt
    python profiles/summarize.py list gpurun_out/launches.csv            > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print("kernel:", name[:110])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-62s %s %s" % (m, vals[i], units[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
        tot_s = sum(int(r[ix["# Samples"]]) for r in data)
        print("  hottest SASS (by stall samples; total %d samples, %d warp instructions):" % (tot_s, tot_i))
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:14]:
            print("    %6.2f%% samples  %6.2f%% inst  wavefronts=%-12s %s" % (
                100.0 * int(r[ix["# Samples"]]) / max(tot_s, 1), 100.0 * int(r[ix["Instructions Executed"]]) / max(tot_i, 1),
                r[ix["L1 Wavefronts Shared"]] if "L1 Wavefronts Shared" in ix else "-", r[ix["Source"]].strip()[:70]))


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1e6 if u == "ns" else (v / 1e3 if u.startswith("us") else v)
        a = agg.setdefault(r[ix["Kernel Name"]][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("launch list (ncu --metrics gpu__time_duration.sum --clock-control none): %d launches, %.1f ms total" % (
        sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("  %-92s n=%4d %11.3f ms %6.2f%%" % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else launch_list)(sys.argv[2])
