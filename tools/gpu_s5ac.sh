#!/bin/bash
# session 5, call ac (1 GPU): ncu --set full of df_units_kernel in the final binary (128-entry pieces)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:df_units_kernel -s 4 -c 1 \
    -o gpurun_out/prof_df_units_final -f python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --parity off \
    --min-seconds 0 > gpurun_out/ncu_final.log 2>&1
echo "ncu rc=$?"
