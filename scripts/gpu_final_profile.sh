#!/bin/bash
# Final measurement session of a round (1 GPU): bench lines of every workload, the reference arm, the ncu launch list of the bench
# command and one `--set full` capture of all hot launches of a C3 frame at 64 spp (raw page exported here: the report is ~110 MB).
T=${1:-r5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${T}_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; tail -c 300 gpurun_out/${T}_bench_c3.json
for wl in C2 C4 C3s C5 CM CD; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${wl}.json 2> gpurun_out/${T}_bench_${wl}.err
  python - $T $wl <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/%s_bench_%s.json" % (sys.argv[1], sys.argv[2])) if l.startswith("{")][-1])
print(sys.argv[2], "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "mrays", round(d["mrays_per_s"],1))
PY
done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 400 gpurun_out/${T}_bench_reference.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${T}_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_(extend|shade|shadow)$' -s 26 -c 26 -f -o /tmp/${T}_full \
    python scripts/profile_target.py cbbunny_area_light_transforms 64 > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
ls -la /tmp/${T}_full.ncu-rep gpurun_out/${T}_raw.csv
echo done
