"""Copy the judged summaries of one gpurun capture into profiles/ (tracked):
    python tools/save_profile.py <tag>     e.g. r1g  (expects gpurun_out/bench_<tag>.json, launches_<tag>.csv, prof_<tag>.ncu-rep)
"""
import csv
import io
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
for src, dst in (("bench_%s.json" % tag, "%s_bench.json" % tag), ("bench_%s_ref.json" % tag, "%s_bench_reference_arm.json" % tag),
                 ("launches_%s.csv" % tag, "%s_launches.csv" % tag)):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    keys += [h for h in hdr if ("sm__inst_executed_pipe_" in h and h.endswith(".avg.pct_of_peak_sustained_active"))
             or ("smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio"))]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(P, "%s_ncu_full_summary.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[ix["Kernel Name"]][:48] for r in rows[2:]])
        for k in keys:
            if k in ix:
                w.writerow([k, units[ix[k]]] + [r[ix[k]] for r in rows[2:]])
    with open(os.path.join(P, "%s_by_line.txt" % tag), "w") as f:
        for kre in ("ahd_select", "median_stage"):
            f.write(subprocess.check_output([sys.executable, "tools/ncu_by_line.py", rep, kre, "--top", "30"]).decode() + "\n")
print("saved", sorted(x for x in os.listdir(P) if x.startswith(tag)))
