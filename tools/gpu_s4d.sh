#!/bin/bash
# session 4, call d (1 GPU): build-path A/B at 10M documents -- warp-striped CSR kernels, ballot vs match.any ranking in
# the one-sweep passes; build parity tests first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide_ngrams.py -x -q -m gpu -k "build or mutation or add_update or tokenizer" > gpurun_out/pytest_s4d.log 2>&1
echo "build tests rc=$?"; tail -4 gpurun_out/pytest_s4d.log
export BENCH_NO_CLOCKS=1
for v in default matchany; do
  if [ $v = matchany ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_matchany.so; fi
  MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 --no-cpu-baseline \
      > gpurun_out/c3_10m_$v.json 2> gpurun_out/c3_10m_$v.trace
  echo "== $v rc=$?"; tail -16 gpurun_out/c3_10m_$v.trace
  python -c "
import json,sys
d=json.loads(open('gpurun_out/c3_10m_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], d['ms_per_step'])"
done
