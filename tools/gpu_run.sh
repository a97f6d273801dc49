# pair GEMM with eight epilogue warps: kernel tests, real-width stack, bench
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" > gpurun_out/r2B_gemm.log 2>&1; rc=$?
echo "gemm tests rc=$rc"; tail -3 gpurun_out/r2B_gemm.log
if [ $rc -eq 0 ]; then
  timeout 900 python -m pytest tests/test_gpu_realwidth.py tests/test_gpu_engine.py -m gpu -q > gpurun_out/r2B_stack.log 2>&1; echo "stack rc=$?"; tail -3 gpurun_out/r2B_stack.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2B_bench.json 2> gpurun_out/r2B_bench.err; echo "bench rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2B_bench.json').read())
print(round(d['value'],1), d['decode_step']); c=d['configs']['configs[2]']; print(round(c['value'],1), c['decode_step'])"
fi
