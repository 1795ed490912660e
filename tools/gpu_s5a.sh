#!/bin/bash
# session 5, call a (1 GPU): neighbour signatures in the posting payload -- payload test, df / build parity tests,
# c2 with the pre-filter on and off, build phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_stream or build or tokenizer or mutation or add_update or large_batch or query_batch or kat or streamed" \
    > gpurun_out/pytest_s5a.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/pytest_s5a.log
export BENCH_NO_CLOCKS=1
for v in sig nosig; do
  unset MGX_DF_NO_SIG
  if [ $v = nosig ]; then export MGX_DF_NO_SIG=1; fi
  MGX_BUILD_TRACE=1 timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5a_$v.json 2> gpurun_out/c2_s5a_$v.err
  echo "== $v rc=$?"; grep "mgx build" gpurun_out/c2_s5a_$v.err | tail -14
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5a_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print(json.dumps(d['kernels']))
print({k:d['batch_stats_per_step'][k] for k in ('df_candidates','df_scanned_docs','n_df_tiles','algo_bytes_df')})
P
done
