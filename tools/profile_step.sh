#!/bin/bash
# GPU tool: launch lists (ncu gpu__time_duration) of one default bench step and of small batches.
# usage: tools/profile_step.sh <tag>
tag=${1:-x}
python bench.py --steps 3 --warmup 2 --no-cpu --no-extras > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/${tag}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv -1 -v > gpurun_out/${tag}_launches.txt
python - <<PY
import json
l=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms_per_step", l["ms_per_step"], "qps", l["value"], "e2e", l["e2e"]["value"])
r=l["roofline"]; print("frac", r["frac"], r["bound"], "tc ms/launch", r["per_launch"]["ms"], r["step"])
PY
cat gpurun_out/${tag}_launches.txt
