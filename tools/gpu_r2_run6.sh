#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err ) 2>&1 | tail -4; echo "ref rc=$?"; cut -c1-1500 gpurun_out/r2_bench_reference.json; tail -3 gpurun_out/r2_bench_reference.err
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | tail -4; echo "ours rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_default.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
print('cpu_baseline', json.dumps(d.get('cpu_baseline'))[:900])
print('port', json.dumps(d.get('cpu_baseline_port'))[:500])
r=d['roofline']; print('roofline', r['kernel'][:40], r['achieved'], r['peak'], r['frac'], r['traffic']); print(json.dumps(r['sweep']))
PY
tail -3 gpurun_out/r2_bench_default.err
