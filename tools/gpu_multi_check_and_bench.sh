NG=${NG:-2}
nvidia-smi -L | wc -l
# sharded == single GPU, on the FP64 statistics path and on the (forced) integer-tensor-core path
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi$NG.log 2>&1; echo "multi rc=$?"; grep -E "shape|MULTI" gpurun_out/multi$NG.log
BTF_STATS_FORCE_I8=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/multi_gpu_check.py > gpurun_out/multi${NG}_i8.log 2>&1; echo "multi i8 rc=$?"; grep -E "shape|MULTI" gpurun_out/multi${NG}_i8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_c2_g$NG.json 2> gpurun_out/bench_c2_g$NG.err; echo "bench c2 x$NG rc=$?"; tail -c 2600 gpurun_out/bench_c2_g$NG.json; tail -3 gpurun_out/bench_c2_g$NG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_g$NG.json 2> gpurun_out/bench_c5_g$NG.err; echo "bench c5 x$NG rc=$?"; tail -c 2600 gpurun_out/bench_c5_g$NG.json; tail -3 gpurun_out/bench_c5_g$NG.err
