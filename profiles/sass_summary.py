#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels in mrmd_b200/libmrmd_b200.so (cuobjdump -sass, sm_100a):
python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mrmd_b200", "libmrmd_b200.so")
HOT = ["ljForceTiledKernelILb1ELb0ELb0E", "verletBuildTiledKernelILb0E", "adressForceTiledKernelILb1ELb0ELb0E",
       "moleculeForceTiledKernelILi4ELb1ELb0ELb0E", "integratePreKernelILb1ELb1E", "haloPushCountedKernel",
       "maxDisplacementGatherKernel", "rsScatterKernel", "permuteAtomsKernel", "shakeFusedKernelILi4E", "shakeAllPairsKernelILi4E", "rattleAllPairsKernelILi4ELb1E"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        kernels[name][m.group(1)] += 1
print(f"{os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, arch sm_100a\n")
for key in HOT:
    for name, ops in kernels.items():
        if key in name:
            total = sum(ops.values())
            fam = collections.Counter()
            for op, c in ops.items():
                fam[op.split(".")[0]] += c
            wide = sum(c for op, c in ops.items() if ".256" in op or ".ENL2" in op)
            print(f"{name}\n  {total} instructions; FP64 {fam['DFMA'] + fam['DADD'] + fam['DMUL'] + fam['DSETP']} "
                  f"(DFMA {fam['DFMA']}, DADD {fam['DADD']}, DMUL {fam['DMUL']}, DSETP {fam['DSETP']}), MUFU {fam['MUFU']}, "
                  f"LDS {fam['LDS']}, STS {fam['STS']}, LDG {fam['LDG']} ({wide} 256-bit), STG {fam['STG']}, SHFL {fam['SHFL']}, "
                  f"VOTE {fam['VOTE']}, BAR {fam['BAR']}, ATOM/RED {fam['ATOM'] + fam['ATOMG'] + fam['RED'] + fam['ATOMS']}, BRA {fam['BRA']}")
            print("  top: " + ", ".join(f"{op} {c}" for op, c in ops.most_common(12)) + "\n")
