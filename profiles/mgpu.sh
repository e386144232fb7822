#!/bin/bash
# multi-GPU checks on the GPU box: profiles/mgpu.sh N tag [modes...]; then the headline bench on N GPUs
N=$1; tag=$2; shift 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29511
for mode in "$@"; do
  port=$((port+1))
  timeout 240 $TR --master-port $port tests/mgpu_check.py 40 $mode 2> gpurun_out/mg_${tag}_${mode}.err | grep '^{' | tail -1 | tee gpurun_out/mg_${tag}_${mode}.json
  tail -3 gpurun_out/mg_${tag}_${mode}.err | cut -c1-400
done
