#!/bin/bash
# GPU tool: everything profiles/ needs from one box (usage: tools/capture_profiles.sh <tag>)
tag=${1:-r02}
out=gpurun_out
# 1. bench lines (no profiler)
python bench.py --steps 20 --warmup 5 > $out/${tag}_BENCH_n1.json 2> $out/${tag}_BENCH_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_BENCH_reference.json 2> $out/${tag}_BENCH_reference.err
# 2. launch list of one search step
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches.csv python tools/exp_profile.py > $out/${tag}_launches.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv -1 -v > $out/${tag}_launches_bench_step.txt
# 3. DRAM traffic of every tc_filter launch of that step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -k regex:tc_filter --csv --log-file $out/${tag}_tc_traffic.csv \
    python tools/exp_profile.py > $out/${tag}_tc_traffic.log 2>&1
# 4. full sections of the tensor-core filter and the exact scan
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"tc_filter|scan_kernel" -c 6 -f \
    -o $out/${tag}_tc_scan python tools/exp_profile.py > $out/${tag}_tc_scan_ncu.log 2>&1
ls -la $out/${tag}_*
python - <<PY
import json
l=json.load(open("$out/${tag}_BENCH_n1.json"))
print("N=1", l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], "frac", l["roofline"]["frac"])
r=json.load(open("$out/${tag}_BENCH_reference.json"))
print("ref", r["value"], r["cpu_baseline"]["sample"])
PY
