#!/bin/bash
# session 4, call a (1 GPU): wide-key (n-gram sizes 4..10) parity tests, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide_ngrams.py -x -q -m gpu > gpurun_out/pytest_wide.log 2>&1
echo "wide rc=$?"; tail -25 gpurun_out/pytest_wide.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_wide_ngrams.py > gpurun_out/pytest_gpu_s4a.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/pytest_gpu_s4a.log
