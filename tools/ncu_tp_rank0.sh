#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs bench.py under ncu (a few gemv_kernel launches of the decode steps, single-pass metrics so
# that no kernel is replayed -- a replayed tensor-parallel kernel would wait for peer arrivals that never come again), the other
# ranks run it plainly.  Usage: python -m torch.distributed.run --nproc-per-node N --no-python bash tools/ncu_tp_rank0.sh <bench args>
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics "$VDD_NCU_METRICS" --clock-control none -k regex:gemv_kernel -s 200 -c 8 --csv --log-file gpurun_out/r2H_tp_ncu.csv python bench.py "$@"
else
  exec python bench.py "$@"
fi
