#!/bin/bash
# round 2 (NG GPUs): sharded == single-GPU check (both statistics paths), then C2 (strong) and C5 (weak) benches
cd "$(dirname "$0")/.."
NG=${NG:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ "${SKIP_CHECK:-0}" != "1" ]; then
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_multi_test_g$NG.log 2>&1
tail -12 gpurun_out/r2_multi_test_g$NG.log
fi
bench() {  # name, extra args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG "$@" > gpurun_out/r2_bench_${name}_g$NG.json 2> gpurun_out/r2_bench_${name}_g$NG.err
  echo "bench $name x$NG rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_${name}_g$NG.json') if l.startswith('{')][-1])
    print('  $name', round(d['value'],2), 'sweeps/s', round(d['ms_per_step'],3), 'ms; e2e', d['e2e'].get('value'), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
    print('  ', {k:round(v,3) for k,v in d['phases_ms'].items()})
except Exception as e:
    print('  $name: no json', e)
PY
  tail -3 gpurun_out/r2_bench_${name}_g$NG.err
}
ENVV="BTF_DUMMY=1" bench c2 --steps 20 --warmup 3
if [ "${SKIP_C5:-0}" != "1" ]; then
ENVV="BTF_DUMMY=1" bench c5 --workload c5 --steps 5 --warmup 3
fi
