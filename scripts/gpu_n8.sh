#!/bin/bash
# 8-GPU session: multi-device tests on distinct GPUs, bench lines (driver-style launch) for C3 and C5, one-call traces
mkdir -p gpurun_out
T=${1:-r4m}
N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${T}_gpu.txt
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests -m gpu -q -k "multi_device" > gpurun_out/${T}_pytest_n${N}.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_n${N}.log; tail -3 gpurun_out/${T}_pytest_n${N}.log
for wl in C3 C5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_n${N}.json 2> gpurun_out/${T}_bench_${wl}_n${N}.err
python - $T $wl $N <<'PY'
import json, sys
try:
    d=json.loads([l for l in open("gpurun_out/%s_bench_%s_n%s.json" % tuple(sys.argv[1:4])) if l.startswith("{")][-1])
    print(sys.argv[2], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["breakdown"], d.get("per_rank_render_ms"))
except Exception as e:
    print("no line", e)
PY
tail -3 gpurun_out/${T}_bench_${wl}_n${N}.err
done
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C3 $N 4 > gpurun_out/${T}_probe_c3.log 2> gpurun_out/${T}_probe_c3.trace; cat gpurun_out/${T}_probe_c3.log
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C5 $N 3 > gpurun_out/${T}_probe_c5.log 2> gpurun_out/${T}_probe_c5.trace; cat gpurun_out/${T}_probe_c5.log
