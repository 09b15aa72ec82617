timeout 1200 python -m pytest tests/test_gpu_posterior.py -q -x -s 2>&1 | tail -15
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:band_blocked -s 3 -c 1 -o gpurun_out/prof_bandblk_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_band.log 2>&1
echo "ncu band rc=$?"
