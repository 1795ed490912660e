#!/bin/bash
# tools/gpu_r2_scale.sh — the multi-GPU session of round 2 on ONE 8-GPU box: C2 at N = 8, 4, 2, 1 (strong scaling, parity in
# every line), C5 (100M documents, 64k-query batches) at N = 8 and C3 (50M-document build) at N = 8. Logs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/scale_gpus.txt 2>&1
nproc >> gpurun_out/scale_gpus.txt; free -g >> gpurun_out/scale_gpus.txt
run() {  # name, n, args...
  local name=$1 n=$2; shift 2
  echo "== $name (N=$n): $*"
  if [ "$n" = "1" ]; then
    timeout ${RUN_TIMEOUT:-600} python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout ${RUN_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
      --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  echo "rc=$? $(tail -c 600 gpurun_out/$name.json | head -c 600)"; tail -3 gpurun_out/$name.err | cut -c1-300
}
for n in ${C2_NS:-8 4 2 1}; do
  run scale_c2_n$n $n --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${C2_ARGS:-}
done
if [ "${RUN_C5:-1}" = "1" ]; then
  run scale_c5_n8 ${C5_N:-8} --config c5 --steps 3 --warmup 2 --parity ${C5_PARITY:-gpu} --no-cpu-baseline ${C5_ARGS:-}
fi
if [ "${RUN_C3:-1}" = "1" ]; then
  run scale_c3_n8 ${C3_N:-8} --config c3 --steps 3 --warmup 1 ${C3_ARGS:-}
fi
