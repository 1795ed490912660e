#!/bin/bash
# session 3, call a (1 GPU): GPU tests incl. the shared-compile test, a C5-shaped single-shard run (12.5M documents,
# 65536-query batches) and one full ncu capture of and_tile_kernel at that shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
CMD="python bench.py --config c5 --docs 12500000 --steps 2 --warmup 3 --no-cpu-baseline --parity off --min-seconds 0.5"
timeout 600 $CMD > gpurun_out/c5shape_1gpu.json 2> gpurun_out/c5shape_1gpu.err
echo "c5 shape rc=$?"; tail -c 1500 gpurun_out/c5shape_1gpu.json
export BENCH_NO_CLOCKS=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:and_tile -s 3 -c 1 \
    -o gpurun_out/prof_and_tile_c5shape -f $CMD > gpurun_out/ncu_c5.log 2>&1
echo "capture rc=$?"; tail -3 gpurun_out/ncu_c5.log
