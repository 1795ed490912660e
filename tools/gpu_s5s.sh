#!/bin/bash
# session 5, call s (1 GPU): commits build the next generation beside the readers: mutation / filter / streamed tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py tests/test_cabi_host.py -x -q -m gpu \
    -k "commit or mutation or add_update or additive or filter or journal or concurrent or adapter or statistics or mgix or load" > gpurun_out/pytest_s5s.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/pytest_s5s.log
