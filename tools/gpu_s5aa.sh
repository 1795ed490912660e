#!/bin/bash
# session 5, call aa (1 GPU): batches in flight 6 / 8 / 12 on the final binary
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
for fl in 6 8 12; do
  timeout 600 python bench.py --no-cpu-baseline --parity off --in-flight $fl > gpurun_out/c2_s5aa_$fl.json 2> gpurun_out/c2_s5aa_$fl.err
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5aa_$fl.json').read().strip().splitlines()[-1])
print('in flight $fl:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e']['per_step_ms'])
P
done
