#!/bin/bash
cd "$(dirname "$0")/.."
NG=${NG:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r2_multi_dbg_fp64.log 2>&1; echo "multi fp64 rc=$?"; grep -E "shape|MULTI|model-level|Error|error|Traceback|btf_b200" gpurun_out/r2_multi_dbg_fp64.log | head -30
BTF_STATS_FORCE_I8=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/multi_gpu_check.py > gpurun_out/r2_multi_dbg_i8.log 2>&1; echo "multi i8 rc=$?"; grep -E "shape|MULTI|model-level|Error|error|Traceback|btf_b200" gpurun_out/r2_multi_dbg_i8.log | head -30
BTF_STATS_FORCE_I8=1 BTF_I8_GEMM2=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29515 tools/multi_gpu_check.py > gpurun_out/r2_multi_dbg_i8old.log 2>&1; echo "multi i8 (old gemm) rc=$?"; grep -E "shape|MULTI|model-level|Error|error|Traceback|btf_b200" gpurun_out/r2_multi_dbg_i8old.log | head -30
# single-GPU i8 tests with the new GEMM forced and not
BTF_I8_GEMM2=1 timeout 600 python -m pytest tests/test_gpu_i8.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize_oracle.py -x -q -m gpu 2>&1 | tail -15
