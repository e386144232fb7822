#!/usr/bin/env python
"""Times ljForceTiledKernel back to back on one melted 1M-atom state (CUDA events through torch on the launching stream):
separates the kernel's own duration from what the step loop adds around it (profiles/r02_build_experiments.md)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmd_b200 import api  # noqa: E402
from mrmd_b200.workloads import lattice_system  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pos, vel, box = lattice_system(side, seed=1)
n = len(pos)
sub = api.Subdomain([0, 0, 0], box, 2.6)
atoms = api.Atoms.from_arrays(pos, vel, mass=1.0)
md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=5, cellSort=True, fullList=2)
md.run(300)
st = md.run(100, timeForceKernel=True)
print("in the step loop: ms/step of the force kernel", st["forceKernelMs"] / 100, "rebuilds", st["rebuilds"])
# the same state through the operators: sort, tiled list, then the force kernel ten times in a row
n = atoms.numLocalAtoms
atoms.numGhostAtoms = 0
atoms.permute(api.LinkedCellList(0, n, [2.6, 2.6, 0.65], sub.minCorner, sub.maxCorner))
vl = api.FullVerletList()
vl.build_periodic(atoms, sub, 2.6, 1.0, 60)
lj = api.LennardJones(2.5, 1.0, 1.0, 0.7)
atoms.setForce(0.0)
lj.apply(atoms, vl)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for k in range(10):
    lj.apply(atoms, vl)  # apply reads the energy back: one synchronisation per call, <ENERGY> variant
    ev[k + 1].record()
torch.cuda.synchronize()
print("back to back (apply = <ENERGY> variant + read-back): ms per call", [round(ev[k].elapsed_time(ev[k + 1]), 4) for k in range(10)])
