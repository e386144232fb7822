"""Run under torchrun with >= 2 ranks (one per GPU): the x-slab NVE run must reproduce the single-GPU periodic
run of the same global system (positions matched atom by atom, pair counts equal, energy within 1e-9)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mrmd_b200 import api, slabs

    api.L().mrmd_b200_set_device(local)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    adress = len(sys.argv) > 2 and sys.argv[2].startswith("adress")
    balanced = len(sys.argv) > 2 and sys.argv[2] == "adress-cuts"
    # global system: jittered sc lattice, x-elongated box (world * 12 x 10 x 10 sites)
    nx, ny = 12 * world, 10
    rng = np.random.default_rng(42)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(ny), indexing="ij"), axis=-1).reshape(-1, 3)
    pos = (g + 0.5) * 1.25 + (rng.random(g.shape) - 0.5) * 0.5
    vel = (rng.random(g.shape) - 0.5) * 1.5
    vel -= vel.mean(axis=0)
    gmin, gmax = np.zeros(3), np.array([nx, ny, ny]) * 1.25
    phys = dict(dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cappingDistance=0.7, maxNeighbors=60)

    extra = {}
    if adress:
        # AT region across the rank boundary in the middle of the box, thermodynamic force updated during the run
        extra = dict(adress=True, weight=api.Slab(gmax / 2, 0.2 * gmax[0], 0.1 * gmax[0], 2), doShift=True,
                     thermo=dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=2, updateInterval=10,
                                 sigma=2.0, range=2.0))
    cuts = None
    if balanced:
        # cost-balanced slab widths: narrow over the AT + HY region, wide over the coarse-grained rest
        cuts = slabs.balanced_cuts(gmin, gmax, world, 0.3 * gmax[0], 0.7 * gmax[0], 2.0, quantum=1.25, min_width=5.2)
        if world == 2:  # the balanced cut of a symmetric region is the middle: take an uneven one instead
            cuts = np.array([0.0, round(0.4 * nx) * 1.25, gmax[0]])
        assert not np.allclose(np.diff(cuts), np.diff(cuts)[0])
    mine = slabs.select_slab(pos, gmin, gmax, rank, world, cuts)
    atoms = api.Atoms.from_arrays(pos[mine], vel[mine], mass=1.0, relativeMass=1.0)
    uid = slabs.broadcast_unique_id(rank)
    md = slabs.SlabMolecularDynamics(atoms, gmin, gmax, rank, world, uid, langevin=False, cuts=cuts, **phys, **extra)
    st = md.run(steps)
    n = st["numLocal"]
    my_pos, my_vel = atoms.getPos()[:n], atoms.getVel()[:n]

    # gather every rank's atoms on rank 0
    counts = [None] * world
    dist.all_gather_object(counts, int(n))
    gathered = [None] * world
    dist.all_gather_object(gathered, (my_pos, my_vel, st["pairInteractions"], st["rebuilds"]))
    ok = True
    msg = {}
    if rank == 0:
        all_pos = np.concatenate([x[0] for x in gathered])
        all_vel = np.concatenate([x[1] for x in gathered])
        pairs = gathered[0][2]  # already the sum over the ranks
        assert all(x[2] == pairs for x in gathered)
        # single-GPU periodic reference run of the same global system
        sub = api.Subdomain(gmin, gmax, phys["rc"] + phys["skin"])
        ref_atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0)
        ref = api.MolecularDynamics(ref_atoms, sub, langevin=False, cellSort=True, fullList=2, **phys, **extra)
        rst = ref.run(steps)
        rp, rv = ref_atoms.getPos()[:len(pos)], ref_atoms.getVel()[:len(pos)]
        from scipy.spatial import cKDTree

        box = gmax - gmin
        wrapped = np.mod(all_pos - gmin, box)
        tree = cKDTree(np.mod(rp - gmin, box), boxsize=box)
        dist_nn, idx = tree.query(wrapped)
        msg = {"atoms": int(len(all_pos)), "expected": int(len(pos)), "max_pos_err": float(dist_nn.max()),
               "bijective": bool(len(np.unique(idx)) == len(pos)),
               "max_vel_err": float(np.abs(all_vel - rv[idx]).max()), "pairs": int(pairs),
               "pairs_ref": int(rst["pairInteractions"]), "energy": st["energy"], "energy_ref": rst["energy"],
               "rebuilds": [x[3] for x in gathered], "rebuilds_ref": rst["rebuilds"], "per_rank": counts}
        ok = (msg["atoms"] == msg["expected"] and msg["bijective"] and msg["max_pos_err"] < 1e-8 and
              msg["max_vel_err"] < 1e-7 and msg["pairs"] == msg["pairs_ref"] and
              abs(msg["energy"] - msg["energy_ref"]) <= 1e-9 * abs(msg["energy_ref"]) and
              all(r == msg["rebuilds_ref"] for r in msg["rebuilds"]))
        print(json.dumps({"ok": ok, **msg}), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    md.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
