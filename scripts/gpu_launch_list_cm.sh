mkdir -p gpurun_out
python bench.py --workload CM --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:'k_shade|k_extend|k_shadow' -s 60 -c 60 --csv --log-file gpurun_out/r5d_launches.csv python bench.py --workload CM --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5d_ncu.log 2>&1
tail -2 gpurun_out/r5d_ncu.log
