timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r2o_tests.log | tail -8
