#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_$name.json'))
    print('  $name', round(d['value'],1), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['state']['nu2'])
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run sf1bar BTF_DUMMY=1
timeout 300 python tools/bench_configs.py k32 > gpurun_out/r2_k32_sf1bar.jsonl 2>&1; cut -c1-700 gpurun_out/r2_k32_sf1bar.jsonl | head -2
