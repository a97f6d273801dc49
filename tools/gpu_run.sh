# launch list of one configs[2]-shape prefill (+ ViT) under ncu: where do the 370 ms go
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2s_prefill_launches.csv \
  python bench.py --profile --batch 32 --prompt-len 64 --new-tokens 2 --steps 1 --warmup 0 --no-configs > gpurun_out/r2s.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2s_prefill_launches.csv
