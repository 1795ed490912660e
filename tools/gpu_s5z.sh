#!/bin/bash
# session 5, call z (1 GPU): whole GPU suite + smoke of the final binary (128-entry df pieces)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final_1gpu.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/pytest_gpu_final_1gpu.log
