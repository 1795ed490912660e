#!/bin/bash
# session 3, call b (1 GPU): GPU tests, the C5-shaped single-shard run after the pruning change, the C2 headline run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c5 --docs 12500000 --steps 2 --warmup 3 --no-cpu-baseline --parity off --min-seconds 0.5 \
  > gpurun_out/c5shape_1gpu_b.json 2> gpurun_out/c5shape_1gpu_b.err
echo "c5 shape rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5shape_1gpu_b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['kernels'], d['config']['streamed_batches'])
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_1gpu_b.json 2> gpurun_out/c2_1gpu_b.err
echo "c2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c2_1gpu_b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['kernels'], d['parity'])
PY
