#!/bin/bash
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
    print('  $name', round(d['value'],1), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'W', d['clocks'].get('power_w_max'), d['state']['nu2'])
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run base BTF_DUMMY=1
run st4 BTF_I8_G2_STAGES=4
run st4co BTF_I8_G2_STAGES=4 BTF_SF_AFTER_DIGITS=1
run st6co BTF_SF_AFTER_DIGITS=1
timeout 300 python tools/bench_configs.py c3 > gpurun_out/r2_c3_pf.jsonl 2>&1; cut -c1-400 gpurun_out/r2_c3_pf.jsonl | head -3
timeout 600 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_i8.py -x -q -m gpu 2>&1 | tail -3
BTF_I8_G2_STAGES=4 timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize_oracle.py -x -q -m gpu 2>&1 | tail -3
