#!/usr/bin/env python
"""The LJ and AdResS step loops on a 10^3 lattice for compute-sanitizer --tool initcheck (slow: a few steps only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmd_b200 import api  # noqa: E402
from mrmd_b200.workloads import lattice_system  # noqa: E402

pos, vel, box = lattice_system(10)
sub = api.Subdomain([0, 0, 0], box, 2.6)
md = api.MolecularDynamics(api.Atoms.from_arrays(pos, vel, mass=1.0), sub, langevin=True, cellSort=True, fullList=2)
print("lj", md.run(6)["pairInteractions"])
w = api.Slab(0.5 * box, 0.2 * box[0], 0.1 * box[0], 1)
md = api.MolecularDynamics(api.Atoms.from_arrays(pos, vel, mass=1.0), sub, langevin=True, cellSort=True, fullList=2, adress=True,
                           weight=w)
print("adress", md.run(6)["pairInteractions"])
