# full attention capture: engine tests that touch attentions + driver shim
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_driver_shim.py -m gpu -q -k "attention or attentions or shim or calibrate or driver" > gpurun_out/r2t_attn.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2t_attn.log
