#!/bin/bash
# final build: full ncu capture of the three kernel families that make up the C2 sweep (after a plain run of the same command)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain_final.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"i8gemm2_kernel|sf_kernel|band_lookahead_kernel" -s 5 -c 10 -o gpurun_out/r2_ncu_c2_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_final.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2_ncu_c2_final.ncu-rep
