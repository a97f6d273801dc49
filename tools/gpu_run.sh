# full 1-GPU validation: all GPU tests, smoke, bench (both arms)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2G_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -E "passed|failed" gpurun_out/r2G_gpu_tests.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2G_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2G_smoke.log
timeout 900 python bench.py > gpurun_out/r2G_bench.json 2> gpurun_out/r2G_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2G_bench_ref.json 2> gpurun_out/r2G_bench_ref.err; echo "ref rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2G_bench.json').read())
print(round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['decode_step']); c=d['configs']['configs[2]']; print(round(c['value'],1), c['decode_step'], c['roofline']['frac'])
for x in d['logit_combine']: print(x['B'], round(x['us_per_launch'],1), round(x['frac_of_hbm_peak'],3))
print(d['clocks'])
r=json.loads(open('gpurun_out/r2G_bench_ref.json').read()); print(r['value'], r['ms_per_step'])"
