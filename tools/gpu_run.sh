# L2 prefetch of the next chain's weights from the few-row attention kernel: sweep
for mb in 0 32 64 100 160; do
  VDD_ATTN_L2_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2q_l2_$mb.json 2> gpurun_out/r2q_l2_$mb.err
  echo "mb=$mb rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r2q_l2_$mb.json').read());print(round(d['value'],1), d['decode_step']['ms_per_token_step'], d['roofline']['frac'])")"
done
