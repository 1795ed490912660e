#!/bin/bash
# session 4, call b (1 GPU): optimistic program evaluation -- parity of the expanded paths, then the single-call classes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_expanded.py tests/test_gpu_wide_ngrams.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/pytest_s4b.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/pytest_s4b.log
timeout 900 python tools/bench_expanded.py --queries 100 --gpu-only --out gpurun_out/expanded_10m_s4b.json > gpurun_out/expanded_s4b.log 2>&1
echo "expanded rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/expanded_10m_s4b.json'))
for k, v in d['classes'].items():
    print(k, round(v['gpu_ms_per_query'], 3), 'ms', round(v['result_docs_mean']))
PY
