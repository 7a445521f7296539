mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4h_pytest.log; tail -3 gpurun_out/r4h_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r4h_bench_c3.json 2> gpurun_out/r4h_bench_c3.err; tail -c 600 gpurun_out/r4h_bench_c3.json
