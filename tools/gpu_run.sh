timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log | head -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2l_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; grep -E "headline|configs" gpurun_out/r2l_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 1 --warmup 0 --new-tokens 6 --profile > gpurun_out/r2l_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 40 -c 2 -o gpurun_out/r2l_ncu_gemv python bench.py --steps 1 --warmup 0 --new-tokens 6 --profile > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu gemv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:combine_sample_rows_kernel -c 1 -o gpurun_out/r2l_ncu_k8rows python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2l_ncu3.log 2>&1; echo "ncu k8 rc=$?"
ls -la gpurun_out | head -30
