mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r4q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4q_pytest.log; tail -3 gpurun_out/r4q_pytest.log
bash scripts/gpu_ncu_shade.sh r4q
