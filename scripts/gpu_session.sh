#!/bin/bash
# One GPU-box session (tuning aid): the -m gpu suite, then A/B of the library variants under ab/ with environment toggles.
# usage: scripts/gpu_session.sh TAG   -> writes gpurun_out/TAG_*.log
TAG=${1:-sess}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
LIB=opencl-raytracing_b200/libraytracing_cuda.so
cp $LIB /tmp/_keep.so
one() {  # variant workload [ENV=...]
  local v=$1 wl=$2; shift 2
  cp ab/$v.so $LIB
  env "$@" python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); sh = d['roofline']['kernel_share_of_step']; ms = d['ms_per_step']
        print('$v', '$wl', '$*', {k: round(d[k], 1) for k in ('value', 'mrays_per_s', 'ms_per_step')}, {k: round(x * ms, 1) for k, x in sh.items()}, 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'])
    elif 'rror' in l or 'Traceback' in l: print(l.strip())
"
}
{
if [ -f scripts/gpu_session_plan.sh ]; then source scripts/gpu_session_plan.sh; fi
} > gpurun_out/${TAG}_ab.log 2>&1
cp /tmp/_keep.so $LIB
cat gpurun_out/${TAG}_ab.log
