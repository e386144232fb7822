import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from mrmd_b200 import api
from mrmd_b200.workloads import lattice_system
pos, vel, box = lattice_system(14, seed=4)
n = len(pos)
sub = api.Subdomain([0, 0, 0], box, 2.6)
extra = dict(adress=True, weight=api.Slab(0.5 * box, 0.25 * box[0], 0.12 * box[0], 1),
             thermo=dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=4, updateInterval=12, sigma=2.0, range=2.0))
if len(sys.argv) > 1 and sys.argv[1] == "nothermo":
    extra.pop("thermo")
def run(queued, chunks):
    if queued: os.environ.pop("MRMD_B200_NO_QUEUED_STEPS", None)
    else: os.environ["MRMD_B200_NO_QUEUED_STEPS"] = "1"
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0)
    md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=5, cellSort=True, fullList=2, **extra)
    if 'every' in sys.argv: md.setEnergyEveryStep(True)
    out = []
    for k in chunks:
        st = md.run(k)
        ids = atoms.get("id")[:n]
        o = np.argsort(ids)
        out.append((atoms.getPos()[:n][o], atoms.getForce()[:n][o], st["rebuilds"]))
    return out
chunks = [1, 17, 30] if 'long' in sys.argv else [1] + [3] * 12
a, b = run(True, chunks), run(False, chunks)
for i, ((pa, fa, ra), (pb, fb, rb)) in enumerate(zip(a, b)):
    print(i, "steps", sum(chunks[:i + 1]), "rebuilds", ra, rb, "dpos", np.abs(pa - pb).max(), "dforce", np.abs(fa - fb).max())
