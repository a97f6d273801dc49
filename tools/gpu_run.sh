timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" > gpurun_out/r2F_gemm.log 2>&1; rc=$?
echo "gemm tests rc=$rc"; tail -3 gpurun_out/r2F_gemm.log
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_realwidth.py -m gpu -q > gpurun_out/r2F_stack.log 2>&1; echo "stack rc=$?"; tail -3 gpurun_out/r2F_stack.log
  timeout 600 python bench.py --batch 32 --prompt-len 64 --new-tokens 64 --steps 3 --warmup 2 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r2F_bench.json 2> gpurun_out/r2F_bench.err; echo "bench rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2F_bench.json').read())
print(round(d['value'],1), d['decode_step'])"
fi
