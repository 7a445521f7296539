# A/B plan of the current session (sourced by gpu_session.sh)
cp ab/new.so $LIB
one precise C3 X=1
python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 26 -f -o /tmp/${TAG}_prof \
    python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_raw.csv
# source-level hot spots of the depth-1 shade launch (launch order per depth: extend, shade, shadow)
ncu --set full --clock-control none --import-source on -k regex:'^k_shade$' -s 10 -c 1 -f -o gpurun_out/${TAG}_shade \
    python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_ncu_shade.log 2>&1
ls -la gpurun_out/${TAG}_shade.ncu-rep
