#!/usr/bin/env python3
"""Summaries of ncu captures for profiles/ (run here, no GPU needed).

    ncu_summary.py rep <file.ncu-rep> [title]     -> key metrics of every kernel in a `--set full` report
    ncu_summary.py launches <launches.csv>        -> per-kernel share of a `--metrics gpu__time_duration.sum` list
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def rep(path, title=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# %s" % (title or path))
    print("# source: ncu --set full --clock-control none --import-source on (%s)" % path)
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print()
        print("Kernel Name = %s" % d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("%s [%s] = %s" % (k, u[k], d[k]))
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.1:
                        print("%s = %s" % (k, d[k]))
                except ValueError:
                    pass


def launches(path):
    rows = list(csv.reader(open(path)))
    h = None
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            h = i
            break
    hdr = rows[h]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
        agg[d["Kernel Name"]][0] += 1
        agg[d["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel share of %s (serialised, cold-cache launches; compare shares, not absolutes)" % path)
    print("# total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%12.3f ms %6.2f%% launches=%-4d %s" % (v[1], 100 * v[1] / tot, v[0], k[:110]))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], " ".join(sys.argv[3:]))
    else:
        launches(sys.argv[2])
