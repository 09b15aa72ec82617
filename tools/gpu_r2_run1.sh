#!/bin/bash
# round 2, call 1 (one GPU): full GPU test suite, then the C2 bench with the stream-overlap variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2_tests1.log 2>&1
tail -8 gpurun_out/r2_tests1.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_$name.json'))
    print('  $name', round(d['value'],1), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run default BTF_DUMMY=1
run noovl BTF_NO_OVERLAP=1
run s3res BTF_I8_STAGES=3 BTF_SF_RESIDENT=1
run res BTF_SF_RESIDENT=1
run s3 BTF_I8_STAGES=3
run s2res BTF_I8_STAGES=2 BTF_SF_RESIDENT=1
