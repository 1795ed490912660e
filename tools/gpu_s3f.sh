#!/bin/bash
# session 3, call f (1 GPU): the stream import test, the whole GPU suite, a second capture of and_tile at the C5 shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_expanded.py -m gpu -x -q -p no:cacheprovider -k "mgix" > gpurun_out/pytest_mgix.log 2>&1
echo "mgix tests rc=$?"; tail -15 gpurun_out/pytest_mgix.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
export BENCH_NO_CLOCKS=1
CMD="python bench.py --config c5 --docs 12500000 --steps 2 --warmup 3 --no-cpu-baseline --parity off --min-seconds 0.5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:and_tile -s 3 -c 1 \
    -o gpurun_out/prof_and_tile_c5shape_b -f $CMD > gpurun_out/ncu_c5_b.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_c5_b.log
