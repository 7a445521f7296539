#!/bin/bash
# Final check of a round (1 GPU): the -m gpu suite with its parity figures printed, smoke, the default bench line and the reference arm
T=${1:-r6}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
grep "^\[parity\]" gpurun_out/${T}_pytest.log > gpurun_out/${T}_parity_on_b200.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
python - $T <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/%s_bench_default.json" % sys.argv[1]) if l.startswith("{")][-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], d["clocks"])
r=json.loads([l for l in open("gpurun_out/%s_bench_reference.json" % sys.argv[1]) if l.startswith("{")][-1])
print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
