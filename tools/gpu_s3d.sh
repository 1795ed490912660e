#!/bin/bash
# session 3, call d (1 GPU): one-sweep sort after moving the look-back behind the ranking phase
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "build" > gpurun_out/pytest_build.log 2>&1
rc=$?; echo "build tests rc=$rc"; tail -3 gpurun_out/pytest_build.log
if [ $rc -ne 0 ]; then exit 1; fi
MGX_BUILD_TRACE=1 timeout 600 python bench.py --config c3 --docs 10000000 --steps 3 --warmup 1 --no-cpu-baseline \
    > gpurun_out/c3_10m_onesweep2.json 2> gpurun_out/c3_10m_onesweep2.err
echo "c3 rc=$?"; grep -E "sort|tokenize|csr" gpurun_out/c3_10m_onesweep2.err | tail -11
