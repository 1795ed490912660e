#!/bin/bash
# session 5, call r (1 GPU): randomised GPU-vs-oracle shake-out of the final build (payload signatures in the df stage)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python tools/fuzz_gpu.py 200 77 > gpurun_out/fuzz_s5r.log 2>&1
echo "fuzz rc=$?"; tail -5 gpurun_out/fuzz_s5r.log
