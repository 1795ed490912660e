#!/bin/bash
# session 3, call j (1 GPU): full ncu captures of the build kernels (tokenizer, one-sweep pass) at 10M documents
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export BENCH_NO_CLOCKS=1
CMD="python bench.py --config c3 --docs 10000000 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_flat -s 1 -c 1 \
    -o gpurun_out/prof_tokenize_flat_10m -f $CMD > gpurun_out/ncu_tok.log 2>&1
echo "tokenizer capture rc=$?"; tail -2 gpurun_out/ncu_tok.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 4 -c 2 \
    -o gpurun_out/prof_onesweep_10m -f $CMD > gpurun_out/ncu_os.log 2>&1
echo "one-sweep capture rc=$?"; tail -2 gpurun_out/ncu_os.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_write -s 1 -c 1 \
    -o gpurun_out/prof_csr_write_10m -f $CMD > gpurun_out/ncu_csr.log 2>&1
echo "csr capture rc=$?"; tail -2 gpurun_out/ncu_csr.log
