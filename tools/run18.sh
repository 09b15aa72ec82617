timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py -q -x -k "negbin" 2>&1 | tail -3
BTF_NB_NO_HIST=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "negbin" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py c4 c1 2>&1 | grep -v "^$" | cut -c1-700
timeout 600 python tools/bench_configs.py k32 > gpurun_out/plain_k32.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 12 -c 2 -o gpurun_out/prof_stats_k32_r1 python tools/bench_configs.py k32 > gpurun_out/ncu_k32.log 2>&1
echo "ncu k32 rc=$?"
