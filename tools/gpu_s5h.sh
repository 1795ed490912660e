#!/bin/bash
# session 5, call h (1 GPU): filter kernel with prefetch + block-reserved 16-byte records; occupancy variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_stream or large_batch or query_batch or kat or streamed or split" > gpurun_out/pytest_s5h.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/pytest_s5h.log
export BENCH_NO_CLOCKS=1
for v in split ff4 ff6 fused; do
  unset MGX_DF_FUSED MGX_LIB_PATH
  if [ $v = fused ]; then export MGX_DF_FUSED=1; fi
  if [ $v = ff4 ] || [ $v = ff6 ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5h_$v.json 2> gpurun_out/c2_s5h_$v.err
  echo "== $v rc=$?"; tail -3 gpurun_out/c2_s5h_$v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5h_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
print({k:d['batch_stats_per_step'][k] for k in ('df_candidates','df_scanned_docs','launches')}, d['run']['streamed_batches'][-60:])
P
done
unset MGX_DF_FUSED MGX_LIB_PATH
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:df_ -s 40 -c 12 --csv \
    --log-file gpurun_out/ncu_s5h_df.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --parity off \
    --min-seconds 0 > gpurun_out/ncu_s5h.log 2>&1
echo "ncu rc=$?"; tail -14 gpurun_out/ncu_s5h_df.csv | cut -c1-260
