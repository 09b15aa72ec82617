#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_tests8.log 2>&1
tail -8 gpurun_out/r2_tests8.log
timeout 300 python tools/bench_configs.py c3 > gpurun_out/r2_c3_occ2.jsonl 2>&1; cut -c1-500 gpurun_out/r2_c3_occ2.jsonl | head -3
BTF_PG_OCC=3 timeout 300 python tools/bench_configs.py c3 > gpurun_out/r2_c3_occ3.jsonl 2>&1; cut -c1-500 gpurun_out/r2_c3_occ3.jsonl | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pg_draw_kernel" -s 3 -c 1 -o gpurun_out/r2_ncu_pg python tools/bench_configs.py c3 > gpurun_out/r2_ncu_pg.log 2>&1; echo "ncu pg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"band_lookahead_kernel" -s 2 -c 1 -o gpurun_out/r2_ncu_band32 python tools/bench_configs.py k32 > gpurun_out/r2_ncu_band32.log 2>&1; echo "ncu band32 rc=$?"
ls -la gpurun_out/*.ncu-rep
