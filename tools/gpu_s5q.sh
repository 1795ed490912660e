#!/bin/bash
# session 5, call q (1 GPU): evidence run of the final build: default bench line, ncu launch list of the same command,
# C3 (10M-document build on one GPU, with the CPU arm) and C4 lines, smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s5q.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_s5q.log
timeout 900 python bench.py > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_c2_reference.json 2> gpurun_out/bench_final_c2_reference.err; echo "ref rc=$?"
tail -c 700 gpurun_out/bench_final_c2_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity off --min-seconds 0 > gpurun_out/ncu_launches_final.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 > gpurun_out/bench_final_c3_10m.json 2> gpurun_out/bench_final_c3_10m.err; echo "c3 rc=$?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err; echo "c4 rc=$?"
for f in bench_final_c2 bench_final_c3_10m bench_final_c4; do
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), d['unit'], round(d['ms_per_step'],3), d['e2e'], d.get('parity',{}) and d['parity'].get('ok'))
    print('   roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','traffic','dram_frac','avg_launch_ms') if k in d['roofline']})
    print('   cpu', d.get('cpu_baseline'))
except Exception as e: print('$f', e)
P
done
