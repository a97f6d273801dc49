# RMSNorm handoff in the gemv chain: parity tests then bench with / without
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_realwidth.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r2D_stack.log 2>&1; echo "stack rc=$?"; grep -E "passed|failed|^E " gpurun_out/r2D_stack.log | head -12
for v in 0 1; do
  if [ $v -eq 1 ]; then export VDD_NO_NORM_HANDOFF=1; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2D_bench_$v.json 2> gpurun_out/r2D_bench_$v.err
  echo "no_handoff=$v rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r2D_bench_$v.json').read());print(round(d['value'],1), d['decode_step']['ms_per_token_step'], d['roofline']['frac'])")"
done
