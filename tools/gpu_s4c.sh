#!/bin/bash
# session 4, call c (1 GPU): scored boolean programs; whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scored_boolean or mixed_boolean" > gpurun_out/pytest_s4c_a.log 2>&1
echo "scored rc=$?"; tail -15 gpurun_out/pytest_s4c_a.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_s4c.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/pytest_gpu_s4c.log
