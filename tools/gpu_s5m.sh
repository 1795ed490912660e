#!/bin/bash
# session 5, call m (1 GPU): occupancy / unit variants of the shipped df kernel; whole GPU suite of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
for v in default o2 u4k u1k; do
  unset MGX_LIB_PATH
  if [ $v != default ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5m_$v.json 2> gpurun_out/c2_s5m_$v.err
  echo "== $v rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5m_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
done
unset MGX_LIB_PATH
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_s5m.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/pytest_gpu_s5m.log
