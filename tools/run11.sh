nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 1500 python bench.py --workload c5 --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_c5_g1.json 2> gpurun_out/bench_c5_g1.err; echo "bench c5 rc=$?"
tail -c 2500 gpurun_out/bench_c5_g1.json; tail -8 gpurun_out/bench_c5_g1.err
