#!/bin/bash
# tools/profile.sh — ncu evidence for the bench command (run under gpurun, 1 GPU).
#   1. plain run of the exact command (must exit 0)
#   2. launch list: every kernel with its device time (cold-cache, serialised: compare SHARES)
#   3. one --set full capture of the top kernel (default: and_tile_kernel)
# Outputs go to gpurun_out/; copy the summaries worth judging into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --docs ${DOCS:-2000000} --steps 1 --warmup 1 --no-cpu-baseline"
KERNEL=${KERNEL:-and_tile_kernel}
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profile_plain.log; exit 1; }
tail -c 400 gpurun_out/profile_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-1200} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-2} -c 2 \
    -o gpurun_out/prof_$KERNEL -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
