#!/bin/bash
# tools/gpu_r2.sh — one GPU session of round 2: gpu tests (new file first), then the bench lines named in BENCH_LIST.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout ${TEST_TIMEOUT:-1500} python -m pytest ${TESTS:-tests} -m gpu -q -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
fi
i=0
IFS=';' read -ra LINES <<< "${BENCH_LIST:-}"
for args in "${LINES[@]}"; do
  i=$((i+1))
  echo "== bench $i: $args"
  MGX_BATCH_TRACE=${TRACE:-0} timeout ${BENCH_TIMEOUT:-900} python bench.py $args > gpurun_out/bench_$i.log 2> gpurun_out/bench_$i.err
  echo "rc=$?"; tail -c 2500 gpurun_out/bench_$i.log; tail -5 gpurun_out/bench_$i.err
done
