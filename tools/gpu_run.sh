# 8-GPU session: TP=8 parity at real width, 7B headline at TP=8 (per-layer launches vs whole-step kernel), BASELINE configs[4]
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "real_width and 8-" > gpurun_out/r2k_tp8_tests.log 2>&1; echo "tp8 tests rc=$?"; tail -3 gpurun_out/r2k_tp8_tests.log | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_tp8.json 2> gpurun_out/r2k_tp8.err; echo "tp8 rc=$?"; grep headline gpurun_out/r2k_tp8.err | head -1
VDD_STEP=1 timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2k_tp8_step.json 2> gpurun_out/r2k_tp8_step.err; echo "tp8 step rc=$?"; grep headline gpurun_out/r2k_tp8_step.err | head -1
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 2 --model llava-v1.5-13b --batch 64 --flags both --no-cpu-baseline > gpurun_out/r2k_cfg4_13b_b64_tp8.json 2> gpurun_out/r2k_cfg4_13b_b64_tp8.err; echo "cfg4 rc=$?"; grep headline gpurun_out/r2k_cfg4_13b_b64_tp8.err | head -1
tail -3 gpurun_out/r2k_cfg4_13b_b64_tp8.err
