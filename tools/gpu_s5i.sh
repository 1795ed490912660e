#!/bin/bash
# session 5, call i (1 GPU): whole GPU suite after the payload-signature work; default bench line (c2) with the CPU arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_s5i.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/pytest_gpu_s5i.log
timeout 900 python bench.py > gpurun_out/bench_c2_s5i.json 2> gpurun_out/bench_c2_s5i.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_c2_s5i.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_c2_s5i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('parity'))
print(d['roofline'])
print(d['cpu_baseline'])
print(d['clocks'])
P
