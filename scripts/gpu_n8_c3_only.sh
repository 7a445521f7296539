#!/bin/bash
# driver-style 8-GPU line of the default workload (C3) only
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_c3_n8.json 2> gpurun_out/r8_bench_c3_n8.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r8_bench_c3_n8.json") if l.startswith("{")][-1])
print("N=8 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["e2e"]["breakdown"], d["render_ms_per_rank"])
PY
