#!/bin/bash
# round-2 evidence run on the GPU box: full -m gpu suite, ncu --set full captures of the dominant kernels, launch lists of
# the three workloads, then the driver's default bench invocation (profiles/r02_final.sh)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_final_tests.log
bash profiles/ncu_capture.sh 2>&1 | tail -6
ll() {  # tag bench-args...
  tag=$1; shift
  MRMD_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/r02_launches_$tag.csv python bench.py --only-headline --no-e2e --no-cpu-baseline --steps 40 --warmup 5 "$@" > /dev/null 2> gpurun_out/r02_launches_$tag.err
  python profiles/launch_summary.py gpurun_out/r02_launches_$tag.csv 40 > gpurun_out/r02_launches_$tag.txt
  head -5 gpurun_out/r02_launches_$tag.txt
}
ll lj
ll adress --workload adress --side 200
ll tetramer --workload tetramer --side 160
python bench.py > gpurun_out/r02_n1_driver.json 2> gpurun_out/r02_n1_driver.err
tail -c 600 gpurun_out/r02_n1_driver.json
