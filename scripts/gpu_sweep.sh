#!/bin/bash
# batch-size sweep of the C3 bench (tuning aid)
for p in 4194304 16777216 67108864; do
  echo "== RTCUDA_MAX_PATHS=$p"
  RTCUDA_MAX_PATHS=$p python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); sh = d['roofline']['kernel_share_of_step']; ms = d['ms_per_step']
        print({k: round(d[k], 1) for k in ('value', 'mrays_per_s', 'ms_per_step')}, {k: round(v * ms, 1) for k, v in sh.items()}, 'e2e', round(d['e2e']['value'], 1))
    elif 'rror' in l: print(l.strip())
"
done
