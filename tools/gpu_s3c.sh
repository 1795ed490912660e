#!/bin/bash
# session 3, call c (1 GPU): one-sweep sort -- parity first (under a short timeout: a look-back bug would spin), then
# the 10M-document build with both sorts (phase trace), then the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "build" > gpurun_out/pytest_build.log 2>&1
rc=$?; echo "build tests rc=$rc"; tail -5 gpurun_out/pytest_build.log
if [ $rc -ne 0 ]; then exit 1; fi
for mode in classic onesweep; do
  MGX_SORT=$mode MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 \
    > gpurun_out/c3_10m_$mode.json 2> gpurun_out/c3_10m_$mode.err
  echo "c3 $mode rc=$?"; tail -c 700 gpurun_out/c3_10m_$mode.json; grep -E "sort|tokenize|csr|total" gpurun_out/c3_10m_$mode.err | tail -14
done
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
