nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python - > gpurun_out/peaks.txt 2>&1 <<'PY'
from functionalmf_b200.engine import fp64_peak, hbm_copy_gbs
print('dfma TF', fp64_peak(0, 0, 20000))
print('dmma TF', fp64_peak(0, 1, 20000))
print('hbm copy GB/s', hbm_copy_gbs(0, 1<<30, 5))
PY
cat gpurun_out/peaks.txt
timeout 900 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 1500 tools/gpu_isolated_tests.sh > /dev/null 2>&1; tail -1 gpurun_out/isolated_tests.log
timeout 600 python bench.py --workload small --steps 5 --warmup 3 --cpu-budget 5 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?"; tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
