# launch list of one configs[2]-shape prefill with the pair GEMM, then ncu --set full of one pair GEMM and one attention launch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2v_prefill_launches.csv \
  python bench.py --profile --batch 32 --prompt-len 64 --new-tokens 2 --steps 1 --warmup 0 --no-configs > gpurun_out/r2v.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel -s 120 -c 2 -o gpurun_out/r2v_pair_gemm -f \
  python bench.py --profile --batch 32 --prompt-len 64 --new-tokens 2 --steps 1 --warmup 0 --no-configs > gpurun_out/r2v2.log 2>&1; echo "ncu pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc_kernel -s 30 -c 1 -o gpurun_out/r2v_fa -f \
  python bench.py --profile --batch 32 --prompt-len 64 --new-tokens 2 --steps 1 --warmup 0 --no-configs > gpurun_out/r2v3.log 2>&1; echo "ncu fa rc=$?"
ls -la gpurun_out/*.ncu-rep
