#!/bin/bash
# A/B of prebuilt library variants under ab/*.so (tuning aid): each is copied over the product library and benched.
# usage: scripts/ab_bench.sh "C3 C5s" base v1 v2
WLS="$1"; shift
LIB=opencl-raytracing_b200/libraytracing_cuda.so
cp $LIB /tmp/_keep.so
for v in "$@"; do
  cp ab/$v.so $LIB
  for wl in $WLS; do
    python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); sh = d['roofline']['kernel_share_of_step']; ms = d['ms_per_step']
        print('$v', '$wl', {k: round(d[k], 1) for k in ('value', 'mrays_per_s', 'ms_per_step')}, {k: round(x * ms, 1) for k, x in sh.items()}, 'e2e', round(d['e2e']['value'], 1))
    elif 'rror' in l: print(l.strip())
"
  done
done
cp /tmp/_keep.so $LIB
