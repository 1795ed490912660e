#!/bin/bash
# session 5, call p (2 GPUs): batches in flight 3 / 6 at N=2 (4 communicator lanes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 3 6; do
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
   bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --in-flight $fl > gpurun_out/c2_s5p_n2_$fl.json 2> gpurun_out/c2_s5p_n2_$fl.err
echo "rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5p_n2_$fl.json').read().strip().splitlines()[-1])
print('N=2 inflight $fl:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e']['per_step_ms'], d.get('parity',{}).get('ok'))
P
done
