#!/bin/bash
# session 5, call ad (1 GPU): overlapped commits (reads never commit, mgx_index_commit publishes): commit test, adapter
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_cabi_host.py -x -q -m gpu -k "commit or adapter or add_update" > gpurun_out/pytest_s5ad.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_s5ad.log
