#!/bin/bash
# session 5, call y (1 GPU): 128-entry pieces in the df collecting loop (4 payload words per lane): df / streamed parity
# subset, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_ or large_batch or query_batch or kat or streamed or full_size" > gpurun_out/pytest_s5y.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/pytest_s5y.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_s5y.json 2> gpurun_out/bench_s5y.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/bench_s5y.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('parity',{}).get('ok'), {k.split(' ')[0]:round(v['ms'],3) for k,v in d['kernels'].items()})
print(d['roofline']['frac'], d['roofline']['dram_frac'], d['roofline']['avg_launch_ms'])
P
