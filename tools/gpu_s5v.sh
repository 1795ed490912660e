#!/bin/bash
# session 5, call v (1 GPU): C5 shape with the document lengths pinned in L2 for the intersect launch (A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
for v in off on; do
  unset MGX_L2_DOCLEN
  if [ $v = on ]; then export MGX_L2_DOCLEN=1; fi
  timeout 900 python bench.py --config c5 --docs 12500000 --steps 3 --warmup 3 --no-cpu-baseline --parity off \
     > gpurun_out/c5shape_s5v_$v.json 2> gpurun_out/c5shape_s5v_$v.err
  echo "== $v rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/c5shape_s5v_$v.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), {k.split(' ')[0]:round(v['ms'],3) for k,v in d['kernels'].items()})
P
done
