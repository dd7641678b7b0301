#!/bin/bash
# ncu --set full captures of the non-headline kernels (run under gpurun; one GPU).  Usage: scripts/ncu_capture.sh <tag> target:kernel_regex:skip ...
tag=$1; shift
mkdir -p gpurun_out
for t in "$@"; do
  IFS=: read -r name k skip <<< "$t"
  ncu --set full --clock-control none --import-source on -k regex:$k -s ${skip:-2} -c 1 -f -o gpurun_out/${tag}_$name \
      python scripts/profile_targets.py $name > gpurun_out/${tag}_${name}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_$name.ncu-rep --page details > gpurun_out/${tag}_${name}_ncu_details.txt 2>&1
  ncu -i gpurun_out/${tag}_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_${name}_ncu_raw.csv 2>&1
  [ -n "$KEEP_REP" ] || rm -f gpurun_out/${tag}_$name.ncu-rep      # the reports (with sources) exceed gpurun's 64 MiB return limit
  grep -E "Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Executed Ipc Active" gpurun_out/${tag}_${name}_ncu_details.txt | head -8
done
