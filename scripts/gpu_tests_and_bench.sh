mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r5e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5e_pytest.log; tail -3 gpurun_out/r5e_pytest.log
python __graft_entry__.py smoke > gpurun_out/r5e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r5e_smoke.log
for wl in CM CD C4; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5e_bench_${wl}.json 2> gpurun_out/r5e_bench_${wl}.err
  python - $wl <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/r5e_bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "mrays", round(d["mrays_per_s"],1))
PY
done
