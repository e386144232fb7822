#!/bin/bash
# one `ncu --set full` capture per dominant kernel, first launch inside the bench's timed region (GPU box):
#   profiles/ncu_capture.sh  ->  gpurun_out/r02_<kernel>.ncu-rep
cap() {  # name kernel-regex bench-args...
  name=$1; k=$2; shift 2
  MRMD_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f \
      -o gpurun_out/r02_$name python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 "$@" > /dev/null 2> gpurun_out/r02_$name.err
  ls -la gpurun_out/r02_$name.ncu-rep
}
cap lj_force ljForceTiledKernel --only-headline
cap build verletBuildTiledKernel --only-headline
cap adress_force adressForceTiledKernel --workload adress --side 200
cap molecule_force moleculeForceTiledKernel --workload tetramer --side 160
cap integrate integratePreKernel --only-headline
