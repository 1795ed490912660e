#!/bin/bash
# session 5, call b (1 GPU): df kernel restructured (collect across pieces, membership per 128 candidates):
# parity subset, then c2 for the unit-size / occupancy variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_stream or large_batch or query_batch or kat or streamed" > gpurun_out/pytest_s5b.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/pytest_s5b.log
export BENCH_NO_CLOCKS=1
for v in default u1k u1ko5 u1ko4 u512o5 u2ko5; do
  unset MGX_LIB_PATH
  if [ $v != default ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5b_$v.json 2> gpurun_out/c2_s5b_$v.err
  echo "== $v rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5b_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
done
