#!/bin/bash
# session 5, call w (2 GPUs): whole GPU suite of the final build (two-rank NCCL tests included), default bench at N=1 and N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
   bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_final2_n2.json 2> gpurun_out/bench_final2_n2.err; echo "n2 rc=$?"
for f in bench_final2_n1 bench_final2_n2; do
python - <<P
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('parity',{}).get('ok'), d['clocks'])
P
done
