#!/bin/bash
# session 5, call d (1 GPU): df collect phase strided over 256-entry pieces, candidates kept as indices, four comparisons in flight per lane
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_stream or large_batch or query_batch or kat or streamed" > gpurun_out/pytest_s5d.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/pytest_s5d.log
export BENCH_NO_CLOCKS=1
for v in default u1ko6 u1ko5 u1ko4 u1ko3 u2ko4; do
  unset MGX_LIB_PATH
  if [ $v != default ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5d_$v.json 2> gpurun_out/c2_s5d_$v.err
  echo "== $v rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5d_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
print({k:d['batch_stats_per_step'][k] for k in ('df_candidates','df_scanned_docs')})
P
done
export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_u1ko4.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:df_units_kernel -s 4 -c 1 \
    -o gpurun_out/prof_df_units_s5d -f python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --parity off \
    --min-seconds 0 > gpurun_out/ncu_s5d.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_s5d.log
