#!/bin/bash
# ncu --set full of the shade launches (depth 0..2) of a bench workload: scripts/gpu_ncu_workload.sh TAG WORKLOAD SPP
tag=${1:-prof}; wl=${2:-CM}; spp=${3:-32}
mkdir -p gpurun_out
python scripts/profile_workload.py $wl $spp > gpurun_out/${tag}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 9 -f -o gpurun_out/${tag}_prof \
    python scripts/profile_workload.py $wl $spp > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log; ls -la gpurun_out/${tag}_prof.ncu-rep
