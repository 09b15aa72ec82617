#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_tests7.log 2>&1
tail -8 gpurun_out/r2_tests7.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_$name.json'))
    print('  $name', round(d['value'],1), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'W', d['clocks'].get('power_w_max'), d['state'])
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run band7 BTF_DUMMY=1
run band7_sfdeep BTF_SF_DEEP=1
timeout 300 python tools/bench_configs.py k32 c4 c1 > gpurun_out/r2_bench_configs4.jsonl 2> gpurun_out/r2_bench_configs4.err; echo "configs rc=$?"; cut -c1-700 gpurun_out/r2_bench_configs4.jsonl
