#!/bin/bash
# GPU tool: ncu --set full of the step's kernels that are neither scan nor filter (usage: tools/ncu_misc.sh <tag> <regex> <count>)
tag=${1:-misc}; rx=${2:-"rerank|merge_check|heap_order"}; cnt=${3:-8}
out=gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"$rx" -c $cnt -f \
    -o $out/${tag} python tools/exp_profile.py > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
python tools/ncu_digest.py $out/${tag}_raw.csv
