#!/bin/bash
# session 5, call t (1 GPU): batch forms of the fuzzy / synonym calls against the single calls and the oracle; timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
true

timeout 900 python tools/bench_expanded.py --gpu-only --queries 200 --out gpurun_out/expanded_s5t.json > gpurun_out/expanded_s5t.log 2>&1
echo "bench_expanded rc=$?"; python - <<P
import json
d=json.load(open('gpurun_out/expanded_s5t.json'))
for k,v in d['classes'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
P
