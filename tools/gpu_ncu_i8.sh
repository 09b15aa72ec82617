# launch list + one full ncu capture of the int8 GEMM (row and column launch) and of the linear-block kernel
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_i8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"i8gemm|sf_kernel" -s 8 -c 4 -o gpurun_out/prof_i8_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i8.log 2>&1
echo "ncu i8 rc=$?"
