#!/bin/bash
# final pass after the allocation-race fix: the whole single-GPU pass of gpu_r2_final1.sh, then the GPU tests with the
# int8 tests FIRST (the order that exposed the race), then a fresh launch list of the default bench
cd "$(dirname "$0")/.."
bash tools/gpu_r2_final1.sh
( time timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_multi.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_list_final.log 2>&1
echo "ncu list rc=$?"
