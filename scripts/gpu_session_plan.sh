# A/B plan of the current session (sourced by gpu_session.sh)
one dense C3 X=1
one new C3 X=1
one dense C5s X=1
one new C5s X=1
one dense C4 X=1
one new C4 X=1
cp ab/new.so $LIB
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_(extend|shade|shadow)<' -s 26 -c 26 -f -o gpurun_out/${TAG}_prof \
    python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
