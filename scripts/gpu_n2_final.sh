mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r7_bench_c3_n2.json 2> gpurun_out/r7_bench_c3_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r7_bench_c3_n2.json") if l.startswith("{")][-1])
print("N=2 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["e2e"]["breakdown"], d["clocks"])
PY
timeout 300 python -m pytest tests -m gpu -q -k "multi_device or page_locked" 2>&1 | tail -2
