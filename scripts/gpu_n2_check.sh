mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r4a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "multi_device or tile_partition or tile_size or sample_range" > gpurun_out/r4a_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4a_pytest_n2.log
tail -3 gpurun_out/r4a_pytest_n2.log
for wl in C3 C5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl > gpurun_out/r4a_bench_${wl}_n2.json 2> gpurun_out/r4a_bench_${wl}_n2.err
tail -c 1500 gpurun_out/r4a_bench_${wl}_n2.json
done
