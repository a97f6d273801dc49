# 4-GPU sanity after the prefill rewrite: TP=4 real-width parity + the driver's N=4 bench line
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "real_width and True-4" > gpurun_out/r2C_tp4_tests.log 2>&1; echo "tp4 tests rc=$?"; grep -E "passed|failed|skipped" gpurun_out/r2C_tp4_tests.log | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2C_bench_tp4.json 2> gpurun_out/r2C_bench_tp4.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2C_bench_tp4.json').read())
print(round(d['value'],1), d['decode_step'])"
