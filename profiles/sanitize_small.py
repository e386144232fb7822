#!/usr/bin/env python
"""Small invocations of every tiled kernel for compute-sanitizer (racecheck / memcheck): cell sort, tiled list build,
LJ force, AdResS force (slab region), molecule sort + list + force (spherical region), one short md run of each.
usage: compute-sanitizer --tool racecheck python profiles/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmd_b200 import api  # noqa: E402
from mrmd_b200.workloads import lattice_system, tetramer_system  # noqa: E402

pos, vel, box = lattice_system(12)
sub = api.Subdomain([0, 0, 0], box, 2.6)
md = api.MolecularDynamics(api.Atoms.from_arrays(pos, vel, mass=1.0), sub, langevin=True, cellSort=True, fullList=2)
print("lj", md.run(12)["pairInteractions"])
w = api.Slab(0.5 * box, 0.2 * box[0], 0.1 * box[0], 1)
md = api.MolecularDynamics(api.Atoms.from_arrays(pos, vel, mass=1.0), sub, langevin=True, cellSort=True, fullList=2, adress=True,
                           weight=w, thermo=dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=2,
                                                 updateInterval=5, sigma=2.0, range=2.0))
print("adress", md.run(12)["pairInteractions"])
pos, vel, box = tetramer_system(8)
sub = api.Subdomain([0, 0, 0], box, 2.6)
w = api.Spherical(0.5 * box, 3.0, 2.0, 2)
md = api.MolecularDynamics(api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25), sub, langevin=True, cellSort=True,
                           fullList=2, adress=True, weight=w, maxNeighbors=40, atomsPerMolecule=4, numConstraintIterations=3,
                           bondLength=1.0)
print("tetramer", md.run(10)["pairInteractions"])
