#!/bin/bash
# tools/gpu_round.sh — one GPU session: all gpu tests, the headline bench (plain), then the ncu launch list of
# the same bench command. Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
fi
MGX_BATCH_TRACE=1 MGX_BUILD_TRACE=1 timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.log; tail -12 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launch list rc=$?"
fi
