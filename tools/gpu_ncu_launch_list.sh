timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 6 -c 2 -o gpurun_out/prof_stats_r1_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stats.log 2>&1
echo "ncu stats rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:band_lookahead -s 3 -c 1 -o gpurun_out/prof_bandblk_r1_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_band.log 2>&1
echo "ncu band rc=$?"
timeout 600 python tools/bench_configs.py c4 > gpurun_out/plain_c4.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_c4.csv python tools/bench_configs.py c4 > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 600 python tools/bench_configs.py c3 > gpurun_out/plain_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pg_draw -s 2 -c 1 -o gpurun_out/prof_pg_r1 python tools/bench_configs.py c3 > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
