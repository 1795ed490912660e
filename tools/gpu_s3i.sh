#!/bin/bash
# session 3, call i (1 GPU): randomised GPU-vs-oracle shake-out of the new paths, single-call classes (two passes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python tools/fuzz_gpu.py 150 11 > gpurun_out/fuzz_r02.log 2>&1
echo "fuzz rc=$?"; tail -4 gpurun_out/fuzz_r02.log
timeout 600 python tools/bench_expanded.py --queries 100 --gpu-only --out gpurun_out/expanded_10m_r02c.json > gpurun_out/expanded_r02c.log 2>&1
echo "expanded rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/expanded_10m_r02c.json'))
for k, v in d['classes'].items():
    print(k, round(v['gpu_ms_per_query'], 3), 'ms', round(v['result_docs_mean']))
PY
