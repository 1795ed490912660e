#!/bin/bash
# session 5, call o (1 GPU): share of the device a persistent kernel takes, on the full shard and on an N=8-sized shard
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
for docs in 1250000 10000000; do
for pct in 100 50 34; do
  for fl in 3 6; do
  export MGX_PERSIST_PCT=$pct
  timeout 900 python bench.py --config c2 --docs $docs --steps 10 --warmup 3 --no-cpu-baseline --parity off --in-flight $fl \
      > gpurun_out/c2_s5o_${docs}_${pct}_$fl.json 2> gpurun_out/c2_s5o_${docs}_${pct}_$fl.err
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5o_${docs}_${pct}_$fl.json').read().strip().splitlines()[-1])
print('docs $docs pct $pct inflight $fl:', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), {k.split(' ')[0]:round(v['ms'],3) for k,v in d['kernels'].items()})
P
  done
done
done
