#!/bin/bash
# session 4, call f (1 GPU): where a one-sweep pass spends its time (clock64 per phase, MGX_OS_TIMING variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_ostiming.so
MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/c3_10m_ostiming.json 2> gpurun_out/c3_10m_ostiming.trace
echo "rc=$?"; tail -22 gpurun_out/c3_10m_ostiming.trace
