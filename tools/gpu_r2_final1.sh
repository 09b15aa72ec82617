#!/bin/bash
# final single-GPU pass: smoke, full GPU suite, default bench (with the CPU baselines), reference arm, the other configurations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r2_tests_final.log 2>&1
tail -8 gpurun_out/r2_tests_final.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench.json')); r=json.load(open('gpurun_out/r2_final_reference.json'))
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
print('ref', r['value'], r['ms_per_step'], 'ratio e2e', d['e2e']['value']/r['value'], 'ratio device', d['value']/r['value'])
print('roofline', d['roofline']['kernel'][:30], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['sweep']['hbm_frac'])
PY
timeout 600 python tools/bench_configs.py > gpurun_out/r2_final_configs.jsonl 2> gpurun_out/r2_final_configs.err; echo "configs rc=$?"; cut -c1-330 gpurun_out/r2_final_configs.jsonl
