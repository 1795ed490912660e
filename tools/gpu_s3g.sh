#!/bin/bash
# session 3, call g (1 GPU): whole GPU suite (list-driven SearchOr / threshold, C++ adapter assertions, stream import),
# C5-shaped run after the stride fix, single-call classes at 10M documents with the CPU oracle beside them.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c5 --docs 12500000 --steps 2 --warmup 3 --no-cpu-baseline --parity off --min-seconds 0.5 \
  > gpurun_out/c5shape_1gpu_c.json 2> gpurun_out/c5shape_1gpu_c.err
echo "c5 shape rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5shape_1gpu_c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], {k.split(' ')[0]: round(v['ms'],2) for k,v in d['kernels'].items()})
PY
timeout 700 python tools/bench_expanded.py --queries 100 --cpu-sample 16 --out gpurun_out/expanded_10m_r02.json \
  > gpurun_out/expanded_r02.log 2>&1
echo "expanded rc=$?"; tail -c 3000 gpurun_out/expanded_r02.log
