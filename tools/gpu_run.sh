# 8-GPU session: TP=8 parity at real width (7B, 4 layers), BASELINE configs[4] with the NCCL all-reduce the rows x ranks rule selects
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "real_width and (True-8 or True-4)" > gpurun_out/r2p_tp8_tests.log 2>&1; echo "tp tests rc=$?"; grep -E "passed|failed" gpurun_out/r2p_tp8_tests.log | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 2 --model llava-v1.5-13b --batch 64 --flags both --no-cpu-baseline > gpurun_out/r2p_cfg4_13b_b64_tp8_nccl.json 2> gpurun_out/r2p_cfg4.err; echo "cfg4 rc=$?"; grep headline gpurun_out/r2p_cfg4.err | head -1
