# 4-GPU session: TP=4 parity at real width, 7B headline at TP=4, BASELINE configs[3] (13B, batch 8, use_dd_unk, TP=4)
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "real_width and 4-" > gpurun_out/r2j_tp4_tests.log 2>&1; echo "tp4 tests rc=$?"; tail -3 gpurun_out/r2j_tp4_tests.log | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_tp4.json 2> gpurun_out/r2j_tp4.err; echo "tp4 rc=$?"; grep headline gpurun_out/r2j_tp4.err | head -1
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 2 --model llava-v1.5-13b --batch 8 --flags unk --no-cpu-baseline > gpurun_out/r2j_cfg3_13b_b8_tp4.json 2> gpurun_out/r2j_cfg3_13b_b8_tp4.err; echo "cfg3 rc=$?"; grep headline gpurun_out/r2j_cfg3_13b_b8_tp4.err | head -1
tail -3 gpurun_out/r2j_cfg3_13b_b8_tp4.err
