#!/bin/bash
# ncu --set full of the depth-0 / depth-1 launches of C5 (16.8 M triangles, 4K) at 4 spp; raw page exported on the box
T=${1:-r5f}
mkdir -p gpurun_out
python scripts/profile_workload.py C5 4 > gpurun_out/${T}_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 6 -f -o /tmp/${T}_c5 \
    python scripts/profile_workload.py C5 4 > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_c5.ncu-rep --page raw --csv > gpurun_out/${T}_c5_raw.csv 2>/dev/null
tail -2 gpurun_out/${T}_ncu_full.log; ls -la gpurun_out/${T}_c5_raw.csv
