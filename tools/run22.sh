timeout 600 python -m pytest tests/test_gpu_posterior.py -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['upload_seconds'], d['e2e'].get('cpu_affinity'), d['e2e']['host_buffers'])"
done
