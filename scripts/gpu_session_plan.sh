# A/B plan of the current session (sourced by gpu_session.sh)
cp ab/new.so $LIB
python -m pytest tests/test_gpu_parity.py -q -k "watertight or multi_device or c3_primary" 2>&1 | tail -3
one new C3 X=1
one twopass C3 X=1
one twopass4 C3 X=1
one twopass C4 X=1
cp ab/new.so $LIB
python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 26 -f -o gpurun_out/${TAG}_prof \
    python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
