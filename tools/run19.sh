timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4
timeout 600 python tools/bench_configs.py k32 2>&1 | grep -v "^$" | cut -c1-700
BTF_STATS_NO_HALVES=1 timeout 600 python tools/bench_configs.py k32 2>&1 | grep tflops
timeout 1500 python bench.py --workload c5 --gpus 1 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 slice value', d['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved']); print(d['phases_ms'])"
