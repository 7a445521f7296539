#!/bin/bash
# C3 and C5s one-liners (tuning aid)
for wl in C3 C5s; do
python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); sh = d['roofline']['kernel_share_of_step']; ms = d['ms_per_step']
        print('$wl', {k: round(d[k], 1) for k in ('value', 'mrays_per_s', 'ms_per_step')}, {k: round(v * ms, 1) for k, v in sh.items()}, 'e2e', round(d['e2e']['value'], 1))
    elif 'rror' in l: print(l.strip())
"
done
