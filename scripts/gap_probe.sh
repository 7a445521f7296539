#!/bin/bash
# how much of a timed step is not inside a kernel span, with and without the clock sampler (diagnostic)
for i in 1 2 3; do for nc in "" 1; do
BENCH_NO_CLOCKS=$nc python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); sh = d['roofline']['kernel_share_of_step']; ms = d['ms_per_step']
        print('noclocks=$nc', round(ms, 1), 'kernels', round(sum(sh.values()) * ms, 1), 'gap', round((1 - sum(sh.values())) * ms, 1), d['clocks'], 'e2e', [round(b['device_render_ms'],1) for b in d['e2e']['breakdown']])
"
done; done
