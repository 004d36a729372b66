#!/bin/bash
# GPU tool: per-kernel durations (ncu launch list of one bench step) for several library variants
# usage: tools/exp_variants.sh <kernel regex> <variant name>...   (variants built by tools/build_variant.sh)
rx=$1; shift
for v in "$@"; do
  AUNCEL_LIB=$PWD/_variants/libauncel_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      -k regex:"$rx" --csv --log-file gpurun_out/var_$v.csv python tools/exp_profile.py > gpurun_out/var_$v.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/var_$v.csv", errors="ignore")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; ci={k:i for i,k in enumerate(rows[h])}
t=[(r[ci["Kernel Name"]].split("(")[0][-28:], float(r[ci["Metric Value"]].replace(",",""))/1e6 if r[ci["Metric Unit"]] in ("ns","nsecond") else float(r[ci["Metric Value"]].replace(",",""))/1e3) for r in rows[h+1:] if len(r)>=len(rows[h])]
print("$v", "total %.3f ms" % sum(x[1] for x in t), [(n, round(x,3)) for n,x in t])
PY
done
