#!/bin/bash
# 2-GPU pass with the final build: sharded == single GPU, engines on two devices of one process
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 ) 2>&1 | tail -20
