#!/bin/bash
# Run every -m gpu test in its own process (a CUDA fault in one test must not poison the
# rest) and summarise into gpurun_out/isolated_tests.log.  Debug aid for gpurun calls.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/isolated_tests.log
: > $LOG
ids=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep '::')
pass=0; fail=0
for id in $ids; do
  out=$(timeout 300 python -m pytest "$id" -x -q --tb=short -p no:cacheprovider 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id" >> $LOG
  else fail=$((fail+1)); echo "FAIL($rc) $id" >> $LOG; echo "$out" | tail -40 >> $LOG; echo "-----" >> $LOG; fi
done
echo "isolated: $pass passed, $fail failed" | tee -a $LOG
