timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_posterior.py -q -x 2>&1 | tail -4
timeout 600 python tools/bench_configs.py c3 c4 2>&1 | grep -v "^$" | cut -c1-800
