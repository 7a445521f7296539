#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (C3), ncu launch list of the bench command and one
# `--set full` capture of the three wavefront kernels on the profile target. Outputs under gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -n 3 gpurun_out/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${tag}_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_c3.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_c3.log
if [ "$2" != "noncu" ]; then
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
python scripts/profile_target.py > gpurun_out/${tag}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_shadow' -s 38 -c 3 -f -o gpurun_out/${tag}_prof \
    python scripts/profile_target.py > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_extend -c 72 --csv \
    --log-file gpurun_out/${tag}_extend_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_traffic_run.log 2>&1
fi
echo done
