for d in 1 2; do
AUNCEL_TC_KERNEL=3 AUNCEL_TC_DRY=$d ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:tc_filter3 --csv --log-file gpurun_out/dry$d.csv python tools/exp_profile.py > gpurun_out/dry$d.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/dry$d.csv", errors="ignore")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; ci={k:i for i,k in enumerate(rows[h])}
t=[float(r[ci["Metric Value"]].replace(",",""))/1e6 if r[ci["Metric Unit"]] in ("ns","nsecond") else float(r[ci["Metric Value"]].replace(",",""))/1e3 for r in rows[h+1:] if len(r)>=len(rows[h])]
print("dry $d (dry, real per round):", [round(x,3) for x in t])
PY
done
