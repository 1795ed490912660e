#!/bin/bash
# session 3, call h (1 GPU): GPU suite, A/B of the df unit size / bound-search hints on C2, single-call classes again
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for v in "" _nohint _u512 _u1024; do
  MGX_LIB_PATH=$PWD/mygram-db_b200/libmgx$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --parity off \
    > gpurun_out/c2_ab$v.json 2> gpurun_out/c2_ab$v.err
  echo "variant '$v' rc=$?"
  python - "$v" <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/c2_ab{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1] or 'default', round(d['value']), round(d['e2e']['value']), {k.split(' ')[0]: round(v['ms'], 3) for k, v in d['kernels'].items()})
PY
done
timeout 600 python tools/bench_expanded.py --queries 100 --gpu-only --out gpurun_out/expanded_10m_r02b.json > gpurun_out/expanded_r02b.log 2>&1
echo "expanded rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/expanded_10m_r02b.json'))
for k, v in d['classes'].items():
    print(k, round(v['gpu_ms_per_query'], 3), 'ms', round(v['result_docs_mean']))
PY
