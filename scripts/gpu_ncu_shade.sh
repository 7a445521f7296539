#!/bin/bash
# ncu --set full of the shade / shadow / extend launches of depth 0 and 1 of the profile target (source import on)
tag=${1:-prof}
mkdir -p gpurun_out
python scripts/profile_target.py cbbunny_area_light_transforms 32 > gpurun_out/${tag}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 6 -f -o gpurun_out/${tag}_prof \
    python scripts/profile_target.py cbbunny_area_light_transforms 32 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out/${tag}_prof.ncu-rep
