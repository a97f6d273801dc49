# 8-GPU sanity of the driver's N=8 line after the prefill rewrite
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2E_bench_tp8.json 2> gpurun_out/r2E_bench_tp8.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2E_bench_tp8.json').read())
print(round(d['value'],1), d['decode_step'])"
