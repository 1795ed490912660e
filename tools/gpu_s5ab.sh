#!/bin/bash
# session 5, call ab (2 GPUs): default bench of the final binary at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 \
   bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_final3_n2.json 2> gpurun_out/bench_final3_n2.err; echo "n2 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/bench_final3_n2.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('parity',{}).get('ok'), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],3))
P
