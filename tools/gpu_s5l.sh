#!/bin/bash
# session 5, call l (1 GPU): without the L2 prefetch of call k;: verification arena (0xFF after every document) + positions of the recorded occurrences in
# it: candidates compared with one read of the text. Parity subset, C2 with and without it, build trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
true # tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q -m gpu \
    -k "payload or df_ or build or large_batch or query_batch or kat or streamed or mutation or add_update" > gpurun_out/pytest_s5l.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/pytest_s5l.log
export BENCH_NO_CLOCKS=1
for v in vtext novtext; do
  unset MGX_DF_NO_VTEXT
  if [ $v = novtext ]; then export MGX_DF_NO_VTEXT=1; fi
  MGX_BUILD_TRACE=1 timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5l_$v.json 2> gpurun_out/c2_s5l_$v.err
  echo "== $v rc=$?"; grep "mgx build" gpurun_out/c2_s5l_$v.err | tail -12
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5l_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'), d['run']['index_resident_gb'])
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
print({k:d['batch_stats_per_step'][k] for k in ('df_candidates','df_scanned_docs','algo_bytes_df')})
P
done
