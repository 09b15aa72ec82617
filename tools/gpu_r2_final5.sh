#!/bin/bash
# column linear block with a split contraction on one GPU: parity at the full size, then the A/B timing
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/sweep_ab.py split: 
