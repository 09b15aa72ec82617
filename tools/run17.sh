timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.out 2>&1
echo "memcheck smoke rc=$?"; tail -5 gpurun_out/memcheck_smoke.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and (small or binom or negbin_all or k8)" > gpurun_out/parity_plain.log 2>&1 && \
timeout 2400 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_parity.log python -m pytest tests/test_gpu_parity.py -q -x -k "golden and (small or binom or negbin_all or k8)" > gpurun_out/memcheck_parity.out 2>&1
echo "memcheck parity rc=$?"; tail -5 gpurun_out/memcheck_parity.log; tail -3 gpurun_out/memcheck_parity.out
