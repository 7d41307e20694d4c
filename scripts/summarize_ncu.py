#!/usr/bin/env python
"""Turn gpurun_out/*.csv (ncu launch list) and *.ncu-rep (ncu --set full) into the text summaries kept under profiles/."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__inst_executed_op_global_red.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list {path}: device time per kernel (us), share of all launches captured"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{a[1]:12.1f} us {100 * a[1] / tot:5.1f}%  launches={a[0]:5d}  avg={a[1] / a[0]:9.1f} us  {k}")
    return "\n".join(out)


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full {path}"]
    for r in rows[2:]:
        out.append(f"## {r[idx['Kernel Name']][:80]}")
        for k in KEYS:
            if k in idx:
                out.append(f"   {k} = {r[idx[k]]} {units[idx[k]]}")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(full(p) if p.endswith(".ncu-rep") else launches(p))
        print()
