"""Wall-clock per phase of the tetramer AdResS step (configs[3]) assembled from the operator API with a device
synchronisation after every call: python profiles/tetramer_phases.py [molecules per edge] [steps].
Shows host-side costs (allocations, read-backs) that a kernel launch list does not."""
import collections
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mrmd_b200 import api  # noqa: E402
from mrmd_b200.workloads import tetramer_system  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pos, vel, box = tetramer_system(side)
n, nm = len(pos), len(pos) // 4
cutoff, skin, dt = 2.6, 0.1, 0.002
sub = api.Subdomain([0, 0, 0], box, cutoff)
w = api.Spherical(0.5 * box, 60.0 / 317.48 * box[0], 30.0 / 317.48 * box[0], 2)
atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25)
mols = api.Molecules(nm)
mols.resize(nm)
mols.set("atomsOffset", np.arange(nm) * 4)
mols.set("numAtoms", np.full(nm, 4))
mols.numLocalMolecules = nm
ghost, vl = api.MultiResGhostLayer(), api.HalfVerletList()
lj = api.LJ_IdealGas(0.7, 2.5, 1.0, 1.0, True)
lj.setAtomsPerMolecule(4)
mc = api.MoleculeConstraints(4, 3)
mc.setConstraints([(i, j, 1.0) for i in range(4) for j in range(i + 1, 4)])
integ = api.VelocityVerletLangevinThermostat(20.0, 1.5, 1234)
acc = collections.defaultdict(float)
cnt = collections.Counter()


def timed(name, fn, *a, **k):
    t0 = time.perf_counter()
    r = fn(*a, **k)
    api.sync()
    acc[name] += time.perf_counter() - t0
    cnt[name] += 1
    return r


max_disp = np.finfo(np.float64).max
for step in range(steps):
    if step == steps // 4:  # drop the first quarter (allocations, first rebuild)
        acc.clear()
        cnt.clear()
        t_start, s_start = time.perf_counter(), step
    timed("shake", mc.enforcePositionalConstraints, mols, atoms, dt)
    max_disp += timed("pre", integ.preForceIntegrate, atoms, dt)
    if max_disp >= skin * 0.5:
        max_disp = 0.0
        timed("rebuild:update_molecules", api.UpdateMolecules.update, mols, atoms, w)
        timed("rebuild:exchange", ghost.exchangeRealAtoms, mols, atoms, sub)
        timed("rebuild:create_ghosts", ghost.createGhostAtoms, mols, atoms, sub)
        timed("rebuild:update_molecules", api.UpdateMolecules.update, mols, atoms, w)
        timed("rebuild:verlet_build", vl.build, mols, 0, mols.numLocalMolecules, cutoff, 1.0, list(sub.minGhostCorner),
              list(sub.maxGhostCorner), 40)
    else:
        timed("ghost_update", ghost.updateGhostAtoms, atoms, sub)
        timed("update_molecules", api.UpdateMolecules.update, mols, atoms, w)
    timed("zero_force", lambda: (atoms.setForce(0.0), mols.setForce(0.0)))
    timed("lj_idealgas", lj.run, mols, vl, atoms, fetch=False)
    timed("contribute", api.ContributeMoleculeForceToAtoms.update, mols, atoms)
    timed("ghost_fold", ghost.contributeBackGhostToReal, atoms)
    timed("post", integ.postForceIntegrate, atoms, dt)
    timed("rattle", mc.enforceVelocityConstraints, mols, atoms, dt)
k = steps - s_start
total = time.perf_counter() - t_start
print(json.dumps({"atoms": n, "steps": k, "ms_per_step": 1e3 * total / k,
                  "phases_ms_per_step": {p: round(1e3 * t / k, 4) for p, t in sorted(acc.items(), key=lambda x: -x[1])},
                  "calls": dict(cnt)}))
