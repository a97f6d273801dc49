export VDD_STEP=1
for t in test_gpu_engine test_gpu_fullsize test_gpu_realwidth; do
  timeout 600 python -m pytest tests/$t.py -m gpu -q -x > gpurun_out/r2h_$t.log 2>&1; echo "$t rc=$?"; tail -3 gpurun_out/r2h_$t.log | head -2
done
VDD_TRACE=1 timeout 300 python tools/trace_step.py gpurun_out/r2h_trace_step.json > gpurun_out/r2h_trace.log 2>&1; echo trace_rc=$?; tail -12 gpurun_out/r2h_trace.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "rc=$?"; grep headline gpurun_out/r2h_bench.err
VDD_STEP_DEPTH=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2h_bench_d1.json 2> gpurun_out/r2h_bench_d1.err; echo "d1 rc=$?"; grep headline gpurun_out/r2h_bench_d1.err
