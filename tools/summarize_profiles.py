#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv,prof_*.ncu-rep} into the tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag>      e.g. r01_v1
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(nsteps=3):
    src = os.path.join(GP, "launches.csv")
    if not os.path.exists(src):
        return
    with open(src) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"].split("(")[0] for r in rows]
    # a step starts at its first sort kernel: k_cell_count / k_scan_onepass (k_scan_tile_sums before round 2's single-pass scan), or the first radix pass (k_hash before it);
    # with the moment sums forked onto the side stream, k_moments / k_sum_partials_par may be listed just before it
    starts = [i for i, n in enumerate(names) if "k_scan_tile_sums" in n or "k_scan_onepass" in n]
    starts = [i - 1 if i > 0 and "k_cell_count" in names[i - 1] else i for i in starts]  # (pass B files the counts on one GPU)
    if not starts:
        starts = [i for i, n in enumerate(names) if "k_radix_pass" in n and (i == 0 or "k_radix_pass" not in names[i - 1])]
        starts = [i - 1 if i > 0 and "k_hash" in names[i - 1] else i for i in starts]
    while_moved = []
    for i in starts:
        while i > 0 and ("k_moments" in names[i - 1] or "k_sum_partials" in names[i - 1] or "k_sm_solve" in names[i - 1]):
            i -= 1
        while_moved.append(i)
    starts = while_moved
    first = starts[-nsteps]
    agg = collections.OrderedDict()
    tot = 0.0
    for r, nm in zip(rows[first:], names[first:]):
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(nm, [0.0, 0, r["Grid Size"], r["Block Size"]])
        a[0] += v
        a[1] += 1
        tot += v
    path = os.path.join(OUT, f"{tag}_launches.csv")
    with open(path, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tools/profile_step.py --workload 8m --steps 3 --warmup 2\n")
        fh.write("# per-launch times are cold-cache and serialised: compare SHARES; the last 3 steps averaged\n")
        fh.write("kernel,launches_per_step,us_per_step,share_pct,grid,block\n")
        for k, (v, c, g, b) in agg.items():
            fh.write(f"{k},{c / nsteps:.2f},{v / nsteps:.1f},{100 * v / tot:.1f},\"{g}\",\"{b}\"\n")
        fh.write(f"TOTAL,,{tot / nsteps:.1f},100.0,,\n")
    print(open(path).read())


def report(name):
    rep = os.path.join(GP, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = {"kernel": d.get("Kernel Name", ("?", ""))[0],
           "source": f"ncu --set full --clock-control none --import-source on -k regex:k_{name} -s 2 -c 1 python tools/profile_step.py --workload 8m"}
    for k in KEYS:
        if k in d:
            out[k] = {"value": d[k][0], "unit": d[k][1]}

    def num(k):
        v, u = d[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    out["dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    path = os.path.join(OUT, f"{tag}_{name}.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(path, "dram bytes/launch %.1f MB" % (out["dram_bytes_per_launch"] / 1e6))
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launches()
    b = report("pass_b")
    report("pass_a")
    if b:
        with open(os.path.join(OUT, "pass_b_traffic.json"), "w") as fh:
            json.dump({"workload": "8m", "dram_bytes_per_launch": b["dram_bytes_per_launch"], "from": f"profiles/{tag}_pass_b.json"}, fh)
