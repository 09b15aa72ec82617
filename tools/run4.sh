timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:band_solve -s 3 -c 1 -o gpurun_out/prof_band_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_band.log 2>&1
echo "ncu band rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 6 -c 2 -o gpurun_out/prof_stats_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stats.log 2>&1
echo "ncu stats rc=$?"
