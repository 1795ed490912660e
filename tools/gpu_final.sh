#!/bin/bash
# tools/gpu_final.sh — the evidence run of a round: GPU tests, headline bench with the CPU baseline, the CPU
# reference arm, then (after those exited 0) the ncu launch list of the bench command and one --set full capture of
# the dominant kernel. Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_reference.log
export BENCH_NO_CLOCKS=1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
KERNEL=${KERNEL:-df_tile_kernel}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 2 -c 1 \
    -o gpurun_out/prof_${KERNEL}_10m -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
if [ -n "${KERNEL2:-}" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL2 -s 2 -c 1 \
      -o gpurun_out/prof_${KERNEL2}_10m -f $CMD > gpurun_out/ncu_full2.log 2>&1
  echo "second capture rc=$?"
fi
if [ "${EXPANDED:-0}" = "1" ]; then
  # fuzzy / synonym / MGIX measurement at 10M documents with the CPU oracle beside it
  timeout 400 python tools/bench_expanded.py --queries 100 --cpu-sample 16 --out gpurun_out/expanded_10m.json \
      > gpurun_out/expanded.log 2>&1
  echo "expanded paths rc=$?"
fi
if [ -n "${FUZZ_SECONDS:-}" ]; then
  timeout $((FUZZ_SECONDS + 60)) python tools/fuzz_gpu.py "$FUZZ_SECONDS" 7 > gpurun_out/fuzz.log 2>&1
  echo "fuzz rc=$?"; tail -1 gpurun_out/fuzz.log
fi
