#!/bin/bash
# session 5, call n (8 GPUs): C2 strong scaling at N = 8 and 4 after the payload-signature work (share ring, NCCL lanes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export C2_NS="8 4" RUN_C3=0 RUN_C5=0 RUN_TIMEOUT=600
bash tools/gpu_r2_scale.sh
for f in scale_c2_n8 scale_c2_n4; do
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e'].get('per_step_ms'), d['e2e'].get('host_compile'), d.get('parity',{}).get('ok'))
    print('  ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
except Exception as e: print('$f', e)
P
done
