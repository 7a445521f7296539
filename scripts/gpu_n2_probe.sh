#!/bin/bash
# 2-GPU diagnostics of the one-call multi-device path: the failing upload test alone and in sequence, then wall-clock traces
mkdir -p gpurun_out
T=${1:-r4i}
timeout 600 python -m pytest tests -m gpu -q -x -k "large_mesh" > gpurun_out/${T}_pytest_alone.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_alone.log; tail -4 gpurun_out/${T}_pytest_alone.log
timeout 600 python -m pytest tests -m gpu -q -k "multi_device" > gpurun_out/${T}_pytest_seq.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_seq.log; tail -4 gpurun_out/${T}_pytest_seq.log
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C3 2 4 > gpurun_out/${T}_probe_c3.log 2> gpurun_out/${T}_probe_c3.trace; cat gpurun_out/${T}_probe_c3.log
RTCUDA_TRACE=1 timeout 600 python scripts/e2e_probe_multi.py C5 2 3 > gpurun_out/${T}_probe_c5.log 2> gpurun_out/${T}_probe_c5.trace; cat gpurun_out/${T}_probe_c5.log
RTCUDA_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload C3 --no-cpu-baseline > gpurun_out/${T}_bench_c3_n2.json 2> gpurun_out/${T}_bench_c3_n2.trace
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/%s_bench_c3_n2.json" % "r4i") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["breakdown"])
PY
