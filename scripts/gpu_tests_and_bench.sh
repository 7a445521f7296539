mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4l_pytest.log; tail -3 gpurun_out/r4l_pytest.log
RTCUDA_TRACE=1 python scripts/e2e_probe_multi.py C3 1 4 2> gpurun_out/r4l_probe_c3.trace | tee gpurun_out/r4l_probe_c3.log
RTCUDA_TRACE=1 python scripts/e2e_probe_multi.py C5 1 3 2> gpurun_out/r4l_probe_c5.trace | tee gpurun_out/r4l_probe_c5.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r4l_bench_c3.json 2> gpurun_out/r4l_bench_c3.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r4l_bench_c3.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['e2e']['breakdown'])"
