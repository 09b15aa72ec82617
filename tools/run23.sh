timeout 900 python -m pytest tests/test_gpu_constrained.py tests/test_gpu_posterior.py -q -x 2>&1 | tail -4
timeout 600 python tools/bench_configs.py c1 2>&1 | grep -v "^$" | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['upload_seconds'])"
