#!/bin/bash
# round 2 (one GPU): C2 bench with the 2-CTA GEMM on/off, band-solve phase profile, ncu launch list + full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_$name.json'))
    print('  $name', round(d['value'],1), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'W', d['clocks'].get('power_w_max'))
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run gemm2 BTF_DUMMY=1
run gemm2_noovl BTF_NO_OVERLAP=1
run gemm1 BTF_I8_GEMM2=0
timeout 300 python tools/bench_configs.py k32 c3 > gpurun_out/r2_bench_configs2.jsonl 2> gpurun_out/r2_bench_configs2.err; echo "configs rc=$?"; cut -c1-600 gpurun_out/r2_bench_configs2.jsonl
# ncu: launch list of the bench command, then full captures of the top kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"i8gemm2_kernel|sf_kernel|band_lookahead_kernel" -c 12 -o gpurun_out/r2_ncu_c2_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2_ncu_c2_top.ncu-rep
