#!/bin/bash
# multi-GPU bench lines of the small configs (C2, C4), driver-style launch: scripts/gpu_nx_small.sh TAG N
T=${1:-r5h}; N=${2:-8}
mkdir -p gpurun_out
for wl in C2 C4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/${T}_bench_${wl}_n${N}.json 2> gpurun_out/${T}_bench_${wl}_n${N}.err
python - $T $wl $N <<'PY'
import json, sys
try:
    d=json.loads([l for l in open("gpurun_out/%s_bench_%s_n%s.json" % tuple(sys.argv[1:4])) if l.startswith("{")][-1])
    print(sys.argv[2], "N", sys.argv[3], "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["e2e"]["breakdown"])
except Exception as e:
    print("no line", e)
PY
done
