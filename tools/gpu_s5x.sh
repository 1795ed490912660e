#!/bin/bash
# session 5, call x (1 GPU): payload words per lane and piece (4 / 8 / 16) in the df collecting loop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
for v in default a4 a6; do
  unset MGX_LIB_PATH
  if [ $v != default ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --parity gpu \
      > gpurun_out/c2_s5x_$v.json 2> gpurun_out/c2_s5x_$v.err
  echo "== $v rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/c2_s5x_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('ok'))
print({k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
done
# tokenizer variant: neighbour signatures computed in the window loop (libmgx_tok.so)
export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_tok.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "payload or build or tokenizer or add_update" > gpurun_out/pytest_s5x_tok.log 2>&1
echo "tok tests rc=$?"; tail -3 gpurun_out/pytest_s5x_tok.log
for v in default tok; do
  unset MGX_LIB_PATH
  if [ $v != default ]; then export MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx_$v.so; fi
  MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 --no-cpu-baseline \
      > gpurun_out/c3_s5x_$v.json 2> gpurun_out/c3_s5x_$v.err
  echo "== c3 $v rc=$?"; grep "tokenize: fused\|csr  " gpurun_out/c3_s5x_$v.err | tail -2
  python -c "
import json
d=json.loads(open('gpurun_out/c3_s5x_$v.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])"
done
