#!/bin/bash
# A/B measurement of library variants on the GPU box (profiles/variants.py builds them):
#   profiles/ab.sh tag variant1 variant2 ...   ("default" = the product library)
# per variant: the headline bench (200 steps, no profiler) and an ncu launch list of a 40-step timed region
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset MRMD_B200_LIB_VARIANT; else export MRMD_B200_LIB_VARIANT=$v; fi
  python bench.py --only-headline --no-e2e --no-cpu-baseline --steps 200 --warmup 5 ${AB_ARGS} > gpurun_out/ab_${tag}_$v.json 2> gpurun_out/ab_${tag}_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${tag}_$v.json"))
    print("$v", "ms/step %.4f" % d["ms_per_step"], "force %.4f" % d["roofline"]["kernel_ms_per_launch"], "value %.3e" % d["value"])
except Exception as e:
    print("$v", "FAILED", e)
PY
  MRMD_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/ll_${tag}_$v.csv python bench.py --only-headline --no-e2e --no-cpu-baseline --steps 40 --warmup 5 ${AB_ARGS} > /dev/null 2> gpurun_out/ll_${tag}_$v.err
  python profiles/launch_summary.py gpurun_out/ll_${tag}_$v.csv 40 | head -6 || true
done
