#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_final_bench.json') if l.startswith('{')][-1])
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['kernel'][:30], d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['kind'], d['cpu_baseline']['value'])
print({k: round(v, 3) for k, v in d['phases_ms'].items()})
PY
