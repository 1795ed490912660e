#!/bin/bash
# session 3, call e (2 GPUs): the 2-rank NCCL tests, then C2 at N=2 with the shared compile ring and without it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_streamed.py -m gpu -x -q -p no:cacheprovider -k "nccl or shared" > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"
  python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['per_step_ms'], d['parity'])
PY
}
run c2_n2_share "--parity gpu"
run c2_n2_noshare "--parity gpu --no-share"
