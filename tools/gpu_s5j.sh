#!/bin/bash
# session 5, call j (2 GPUs): the sharded pipeline over NCCL after the payload-signature work: 2-GPU tests, C2 at N=2
# and N=1 on the same box, C5 shape (25M documents, 64k-query batches) at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_streamed.py -x -q -m gpu > gpurun_out/pytest_2gpu_s5j.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_2gpu_s5j.log
export C2_NS="2 1" RUN_C3=0 C5_N=2 C5_ARGS="--docs 25000000" RUN_TIMEOUT=900
bash tools/gpu_r2_scale.sh
for f in scale_c2_n2 scale_c2_n1 scale_c5_n8; do
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e'].get('per_step_ms'), d.get('parity',{}).get('ok'))
    print('  ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
except Exception as e: print('$f', e)
P
done
