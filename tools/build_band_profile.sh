#!/bin/bash
# Instrumented build of the library (per-phase clock64 totals of the band kernels, -DBTF_BAND_PROFILE) next to the product
# library: functionalmf_b200/libbtf_b200_prof.so; use it with BTF_B200_LIB=<path> python tools/band_profile.py
set -e
HERE="$(cd "$(dirname "$0")/../functionalmf_b200/csrc" && pwd)"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
mkdir -p "$HERE/_obj" "$HERE/_obj_prof"
$NVCC $FLAGS -DBTF_BAND_PROFILE -c "$HERE/band_lookahead.cu" -o "$HERE/_obj_prof/band_lookahead_prof.o"
objs=$(ls "$HERE"/_obj/*.o | grep -v band_lookahead)
$NVCC -shared -o "$HERE/../libbtf_b200_prof.so" $objs "$HERE/_obj_prof/band_lookahead_prof.o" -ldl
echo built "$HERE/../libbtf_b200_prof.so"
