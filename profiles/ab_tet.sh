#!/bin/bash
# tetramer force kernel: tile size / threads sweep on one GPU (side 100 = 4M atoms): variant:homes:smemKB
for cfg in "$@"; do
  IFS=: read v homes kb <<< "$cfg"
  if [ "$v" = default ]; then unset MRMD_B200_LIB_VARIANT; else export MRMD_B200_LIB_VARIANT=$v; fi
  MRMD_B200_MOL_HOMES=$homes MRMD_B200_MOL_SMEM_KB=$kb python bench.py --workload tetramer --side 100 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/abt.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', 'ms/step %.4f' % d['ms_per_step'], 'force %.4f' % d['roofline']['kernel_ms_per_launch'])" || tail -3 gpurun_out/abt.err
done
