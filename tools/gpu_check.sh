#!/bin/bash
# tools/gpu_check.sh — run every GPU test in its own process (a CUDA fault in one test must not
# poison the others) with a per-test timeout; summary in gpurun_out/gpu_check.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/gpu_check.log
: > $LOG
nvidia-smi --query-gpu=name,memory.total --format=csv >> $LOG 2>&1
tests=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::")
pass=0; fail=0
for t in $tests; do
  out=$(timeout ${PER_TEST_TIMEOUT:-300} python -m pytest "$t" -x -q -p no:cacheprovider 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $t" >> $LOG
  else fail=$((fail+1)); echo "FAIL($rc) $t" >> $LOG; echo "$out" | tail -40 >> $LOG; fi
done
echo "passed=$pass failed=$fail" | tee -a $LOG
[ $fail -eq 0 ]
