"""Times SPC::applyForces (mrmd_b200_spc_apply_forces) on a water box: python profiles/spc_timing.py [sites] [reps]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mrmd_b200 import api  # noqa: E402
from oracle.md_loop import spc_water_box  # noqa: E402  (start configuration only)

sites = int(sys.argv[1]) if len(sys.argv) > 1 else 40
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pos, vel, mass, q, rm, typ, box = spc_water_box(sites)
n, nm = len(pos), len(pos) // 3
atoms = api.Atoms.from_arrays(pos, vel, mass=mass, type=typ, relativeMass=rm)
atoms.set("charge", q)
mols = api.Molecules(nm)
mols.resize(nm)
mols.set("atomsOffset", np.arange(nm) * 3)
mols.set("numAtoms", np.full(nm, 3))
mols.numLocalMolecules = nm
api.UpdateMolecules.update(mols, atoms, api.Slab(0.5 * box, 1e3, 1.0, 7))
vl = api.HalfVerletList()
vl.build(mols, 0, nm, 1.3, 1.0, np.full(3, -0.2), box + 0.2, 230)
pairs = vl.info()["totalPairs"]
out = {}
for kind in (0, 1):
    spc = api.SPC(kind)
    L = api.L()
    for _ in range(3):
        L.mrmd_b200_spc_apply_forces(spc.h, mols.h, vl.h, atoms.h, None, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        L.mrmd_b200_spc_apply_forces(spc.h, mols.h, vl.h, atoms.h, None, None, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["dsf" if kind else "plain"] = {"ms": ms, "molecule_pairs_per_s": pairs / ms * 1e3, "site_pairs_per_s": 10 * pairs / ms * 1e3}
print(json.dumps({"molecules": nm, "molecule_pairs": pairs, **out}))
