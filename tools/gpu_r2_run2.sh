#!/bin/bash
# round 2, call (one GPU): full GPU test suite, then the C2 bench with the chunk-pipeline variants, C3 (Binomial) bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2_tests2.log 2>&1
tail -15 gpurun_out/r2_tests2.log
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
    r=d['roofline']; print('   roofline', r['kernel'][:30], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), '| sweep hbm_frac', round(r['sweep']['hbm_frac'],3), '| i8 peak', [round(v['peak']) for k,v in list(r['other_kernels'].items())+[('x',r)] if 'i8gemm' in v['kernel']])
except Exception as e:
    print('  $name: no json', e)
PY
  tail -2 gpurun_out/r2_bench_$name.err
}
run chunks4 BTF_COL_CHUNKS=4
run chunks1 BTF_COL_CHUNKS=1
run chunks2 BTF_COL_CHUNKS=2
run chunks8 BTF_COL_CHUNKS=8
run chunks4noprio BTF_COL_CHUNKS=4 BTF_NO_PRIO=1
run noguard BTF_COL_CHUNKS=4 BTF_I8_NO_GUARD=1
timeout 600 python tools/bench_configs.py > gpurun_out/r2_bench_configs.jsonl 2> gpurun_out/r2_bench_configs.err; echo "configs rc=$?"; cat gpurun_out/r2_bench_configs.jsonl | cut -c1-400
