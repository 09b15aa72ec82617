set -x
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 4 -c 2 -o gpurun_out/prof_stats_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/
