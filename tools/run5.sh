nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi2.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_g2.json 2> gpurun_out/bench_c2_g2.err; echo "bench2 rc=$?"; tail -c 1200 gpurun_out/bench_c2_g2.json; tail -5 gpurun_out/bench_c2_g2.err
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_c2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
print(d['phases_ms'])
PY
