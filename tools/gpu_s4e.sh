#!/bin/bash
# session 4, call e (1 GPU): build path after the tokenizer / histogram / look-back changes: parity tests, phase traces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide_ngrams.py -x -q -m gpu -k "build or mutation or add_update or tokenizer" > gpurun_out/pytest_s4e.log 2>&1
echo "build tests rc=$?"; tail -4 gpurun_out/pytest_s4e.log
export BENCH_NO_CLOCKS=1
for v in default ballot match; do
  unset MGX_OS_RANK
  if [ $v != default ]; then export MGX_OS_RANK=$v; fi
  MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 --no-cpu-baseline \
      > gpurun_out/c3_10m_e_$v.json 2> gpurun_out/c3_10m_e_$v.trace
  echo "== $v rc=$?"; tail -12 gpurun_out/c3_10m_e_$v.trace
  python -c "
import json,sys
d=json.loads(open('gpurun_out/c3_10m_e_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], d['ms_per_step'])"
done
