#!/bin/bash
# session 5, call u (1 GPU): adapter assertions (batch forms, overlapped commits); and_tiles at the C5 shape under ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cabi_host.py -x -q -m gpu > gpurun_out/pytest_s5u.log 2>&1
echo "adapter tests rc=$?"; tail -4 gpurun_out/pytest_s5u.log
export BENCH_NO_CLOCKS=1
timeout 900 python bench.py --config c5 --docs 12500000 --steps 2 --warmup 3 --no-cpu-baseline --parity off \
   > gpurun_out/c5shape_s5u.json 2> gpurun_out/c5shape_s5u.err
echo "c5 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/c5shape_s5u.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), {k.split(' ')[0]:round(v['ms'],3) for k,v in d['kernels'].items()})
print({k:v for k,v in d['batch_stats_per_step'].items() if k in ('n_and_tiles','driver_entries','result_docs','launches')})
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:and_tile -s 3 -c 1 \
    -o gpurun_out/prof_and_tile_c5shape_s5u -f python bench.py --config c5 --docs 12500000 --steps 1 --warmup 1 \
    --no-cpu-baseline --parity off --min-seconds 0 > gpurun_out/ncu_s5u.log 2>&1
echo "ncu rc=$?"
