"""CPU tool: turn what tools/capture_profiles.sh brought back in gpurun_out/ into the tracked files under
profiles/ (usage: python tools/summarize_profiles.py r02)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src, dst = "gpurun_out", "profiles"
for f in (f"{tag}_BENCH_n1.json", f"{tag}_BENCH_reference.json", f"{tag}_launches_bench_step.txt"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))

# DRAM traffic per tc_filter launch
p = os.path.join(src, f"{tag}_tc_traffic.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 14 and r[0].isdigit()]
    per = {}
    for r in rows:
        per.setdefault(int(r[0]), {})[r[12]] = float(r[14])
    launches = []
    for i in sorted(per):
        m = per[i]
        ms = m["gpu__time_duration.sum"] / 1e6
        launches.append({"launch": i, "dram_read_bytes": m["dram__bytes_read.sum"], "dram_write_bytes": m["dram__bytes_write.sum"],
                         "ms_under_ncu": ms, "dram_gbs": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / ms / 1e6})
    out = {"note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tc_filter on one "
                   "bench step (tools/capture_profiles.sh); durations under ncu are cold-cache and serialised",
           "launches": launches,
           "dram_bytes_per_launch": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / max(1, len(launches))}
    json.dump(out, open(os.path.join(dst, f"{tag}_tc_traffic.json"), "w"), indent=1)

# key metrics of the full-section capture
rep = os.path.join(src, f"{tag}_tc_scan.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
    tens = [h for h in hdr if "tensor" in h and "pct_of_peak" in h and "ops_path" not in h and ".min." not in h and ".max." not in h and ".sum." not in h]
    cols = sorted(set(want + tens))
    with open(os.path.join(dst, f"{tag}_tc_scan_ncu.txt"), "w") as f:
        f.write("ncu --set full --clock-control none, one error-bounded search of the bench workload "
                "(tools/capture_profiles.sh); kernels in launch order\n")
        f.write("tensor metric columns present: " + ", ".join(tens) + "\n")
        ki = hdr.index("Kernel Name")
        for r in data:
            f.write(f"\n== {r[ki].split('(')[0]}  id {r[0]}\n")
            for c in cols:
                if c in hdr:
                    j = hdr.index(c)
                    f.write(f"   {c:90s} {r[j]} {units[j]}\n")
print("profiles/ updated for", tag)
