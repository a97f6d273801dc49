for st in 0 3 4 5; do
  VDD_DEEP_STAGES=$st timeout 300 python bench.py --batch 32 --prompt-len 64 --new-tokens 64 --steps 3 --warmup 2 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r2y_$st.json 2> gpurun_out/r2y_$st.err
  echo "stages=$st rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r2y_$st.json').read());print(round(d['value'],1), d['decode_step']['ms_per_token_step'], d['roofline']['frac'])")"
done
