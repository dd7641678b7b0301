#!/bin/bash
# per-instruction (SASS) execution counts of one kernel launch: scripts/ncu_source.sh <tag> <target> <kernel_regex> <skip>
tag=$1; name=$2; k=$3; skip=${4:-2}
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${tag}_$name \
    python scripts/profile_targets.py $name > gpurun_out/${tag}_${name}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_$name.ncu-rep --page source --csv > gpurun_out/${tag}_${name}_source.csv 2>&1
rm -f gpurun_out/${tag}_$name.ncu-rep
wc -l gpurun_out/${tag}_${name}_source.csv
