#!/usr/bin/env python
"""Summarise one GPU round (tools/gpu_round.sh TAG) into profiles/:
   TAG_launch_summary.txt  per-kernel totals of the launch list (gpu__time_duration.sum)
   TAG_ncu_full_summary.txt selected `ncu --set full` metrics of one steady-state tick
   traffic.json            DRAM bytes per launch of the load / solve kernels (bench.py reads it)
Usage: python tools/ncu_summary.py TAG [trials_in_full_capture]"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
cap_trials = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

METRICS = """gpu__time_duration.sum launch__registers_per_thread launch__grid_size launch__block_size
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
lts__t_sectors.sum l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
smsp__inst_executed.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__warps_eligible.avg.per_cycle_active""".split()


def launch_summary():
    path = os.path.join(G, f"{tag}_launches.csv")
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = collections.OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
        k = r["Kernel Name"].split("(")[0]
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    s = sum(a[1] for a in tot.values())
    out = [f"ncu launch list summary, capture {tag} (tools/gpu_round.sh: ncu --metrics gpu__time_duration.sum "
           "--clock-control none --launch-skip 3000 -c 400 python bench.py --kernel-only --steps 1 --warmup 1)",
           "RO-501 x 10000 trials, 400 consecutive steady-state launches; per-launch times are cold-cache and serialised",
           f"{'kernel':34s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}"]
    for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:34s} {a[0]:8d} {a[1]:12.1f} {a[1]/a[0]:10.1f} {100*a[1]/s:6.1f}%")
    open(os.path.join(P, f"{tag}_launch_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full_summary():
    rep = os.path.join(G, f"{tag}_prof.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"ncu --set full --clock-control none, capture {tag}: one steady-state tick of RO-501 x {cap_trials} trials "
           "(python bench.py --kernel-only --trials %d)" % cap_trials]
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        out.append(f"== {name[:60]} id {r[col['ID']]}")
        for m in METRICS:
            if m in col:
                out.append(f"   {m:86s} {r[col[m]]} {units[col[m]]}")
        key = name.split("(")[0].split("<")[0].replace("void ", "").strip()

        def gb(metric):
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]]
            return v*{"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        if key in ("k_load_mos", "k_lu", "k_assemble", "k_finish") and key not in traffic:
            traffic[key] = {"source": f"profiles/{tag}_ncu_full_summary.txt (ncu --set full --clock-control none, "
                                      "RO-501, one steady-state tick)",
                            "trials": cap_trials, "dram_bytes_read": gb("dram__bytes_read.sum"),
                            "dram_bytes_write": gb("dram__bytes_write.sum"),
                            "us": float(r[col["gpu__time_duration.sum"]].replace(",", ""))}
    open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


launch_summary()
full_summary()
