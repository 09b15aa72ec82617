#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_tests5.log 2>&1
tail -8 gpurun_out/r2_tests5.log
BTF_B200_LIB=$PWD/functionalmf_b200/libbtf_b200_prof.so timeout 300 python tools/band_profile.py > gpurun_out/r2_band_profile.log 2>&1; echo "band profile rc=$?"; cat gpurun_out/r2_band_profile.log | cut -c1-400
