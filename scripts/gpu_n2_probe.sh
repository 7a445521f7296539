#!/bin/bash
# 2-GPU check of the one-call multi-device path: the multi-device tests, wall-clock traces of rc.render(num_devices=2), the bench lines
mkdir -p gpurun_out
T=${1:-r4k}
timeout 600 python -m pytest tests -m gpu -q -k "multi_device or tile_partition or tile_size or sample_range" > gpurun_out/${T}_pytest_n2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_n2.log; tail -3 gpurun_out/${T}_pytest_n2.log
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C3 2 4 > gpurun_out/${T}_probe_c3.log 2> gpurun_out/${T}_probe_c3.trace; cat gpurun_out/${T}_probe_c3.log
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C5 2 3 > gpurun_out/${T}_probe_c5.log 2> gpurun_out/${T}_probe_c5.trace; cat gpurun_out/${T}_probe_c5.log
for wl in C3 C5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_n2.json 2> gpurun_out/${T}_bench_${wl}_n2.err
python - $T $wl <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/%s_bench_%s_n2.json" % (sys.argv[1], sys.argv[2])) if l.startswith("{")][-1])
print(sys.argv[2], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["breakdown"])
PY
done
