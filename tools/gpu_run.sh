timeout 900 python -m pytest tests/test_gpu_realwidth.py -m gpu -q -k "fused_rope" > gpurun_out/r2x.log 2>&1; echo "rc=$?"; grep -E "^E|passed|failed" gpurun_out/r2x.log | head
