"""x-slab decomposition over the GPUs of one node: host-side geometry and the Python face of
mrmd_b200_slab_* (include/mrmd_b200.h).  One process per GPU; torch.distributed is only used to hand the NCCL
unique id of rank 0 to the other ranks (any backend: nccl on the GPU box, gloo in the CPU tests)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def balanced_cuts(global_min, global_max, nranks, active_lo, active_hi, force_cost_ratio, quantum=None,
                  min_width=0.0):
    """Slab boundaries cuts[0..nranks] with equal estimated cost per slab (narrow slabs over the
    AT / HY region [active_lo, active_hi), wide ones over the coarse-grained rest).  Cost per unit length is 1 outside
    the active region and 1 + force_cost_ratio inside (force kernel time / everything else, per atom).  Boundaries
    are rounded to multiples of `quantum` (a lattice plane spacing) and kept at least min_width apart."""
    gmin, gmax = float(global_min[0]), float(global_max[0])
    xs = np.array([gmin, min(max(active_lo, gmin), gmax), min(max(active_hi, gmin), gmax), gmax])
    dens = np.array([1.0, 1.0 + force_cost_ratio, 1.0])
    cum = np.concatenate([[0.0], np.cumsum(dens * np.diff(xs))])
    targets = cum[-1] * np.arange(1, nranks) / nranks
    inner = np.interp(targets, cum, xs)
    if quantum:
        inner = gmin + np.round((inner - gmin) / quantum) * quantum
    cuts = np.concatenate([[gmin], inner, [gmax]])
    for r in range(1, nranks):  # enforce the minimum width from the left, then from the right
        cuts[r] = max(cuts[r], cuts[r - 1] + min_width)
    for r in range(nranks - 1, 0, -1):
        cuts[r] = min(cuts[r], cuts[r + 1] - min_width)
    assert np.all(np.diff(cuts) >= min_width - 1e-12), "box too short for that many slabs"
    return cuts


def slab_bounds(global_min, global_max, rank, nranks, cuts=None):
    """[xlo, xhi) of rank `rank`: the caller's cuts, or equal-width slabs where the last one ends exactly at the
    global maximum (same arithmetic as mrmd_b200_slab_create)."""
    if cuts is not None:
        return float(cuts[rank]), float(cuts[rank + 1])
    gmin, gmax = float(global_min[0]), float(global_max[0])
    width = (gmax - gmin) / nranks
    lo = gmin + rank * width
    hi = gmax if rank == nranks - 1 else gmin + (rank + 1) * width
    return lo, hi


def neighbours(rank, nranks):
    """(left, right) ranks of the periodic chain"""
    return (rank - 1 + nranks) % nranks, (rank + 1) % nranks


def boundary_shifts(global_min, global_max, rank, nranks):
    """x shifts applied when an atom is sent (to the left, to the right): the periodic wrap is done by the two
    end ranks"""
    lx = float(global_max[0]) - float(global_min[0])
    return (lx if rank == 0 else 0.0), (-lx if rank == nranks - 1 else 0.0)


def owner_of(x, global_min, global_max, nranks, cuts=None):
    """rank that owns coordinate(s) x (x already wrapped into the global box)"""
    owners = np.zeros(np.shape(x), dtype=np.int64)
    for r in range(nranks):
        lo, hi = slab_bounds(global_min, global_max, r, nranks, cuts)
        owners[(np.asarray(x) >= lo) & (np.asarray(x) < hi)] = r
    return owners


def select_slab(pos, global_min, global_max, rank, nranks, cuts=None):
    """indices of the atoms of a global configuration that belong to `rank`"""
    lo, hi = slab_bounds(global_min, global_max, rank, nranks, cuts)
    return np.nonzero((pos[:, 0] >= lo) & (pos[:, 0] < hi))[0]


def bind_host_to_gpu(device_index):
    """Pins this process to the CPU cores of the NUMA node the GPU hangs on, so that the pinned host buffers of the
    host-buffer path (mrmd_b200_slab_run_host) are allocated node-local (first touch) and the copy streams are fed from
    that node: with one process per GPU and no binding, every rank's buffers end up on the node the launcher started on and
    half of the GPUs copy across the socket link.  Returns the previous affinity (restore with os.sched_setaffinity), or
    None when the topology cannot be read (no sysfs entry, node -1) -- then nothing is changed."""
    import os

    try:
        import torch

        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        before = os.sched_getaffinity(0)
        cpus &= before  # stay inside the cpuset the container grants
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return before
    except Exception:
        return None


def broadcast_unique_id(rank):
    """128-byte NCCL unique id of rank 0 on every rank (torch.distributed must be initialised)"""
    import torch
    import torch.distributed as dist

    buf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        check(_lib.load().mrmd_b200_nccl_unique_id(buf.ctypes.data))
    t = torch.from_numpy(buf)
    backend = dist.get_backend()
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return t.cpu().numpy().copy()


class SlabMolecularDynamics:
    """The LJ step loop of MolecularDynamics on this rank's x-slab (mrmd_b200_slab_*)."""

    def __init__(self, atoms, global_min, global_max, rank, nranks, unique_id, dt=0.002, rc=2.5, skin=0.1, sigma=1.0,
                 epsilon=1.0, cappingDistance=0.7, maxNeighbors=60, langevin=False, zeta=20.0, temperature=1.5,
                 seed=1234, adress=False, weight=None, doShift=True, thermo=None, cuts=None, atomsPerMolecule=1,
                 numConstraintIterations=0, bondLength=1.0):
        cfg = _lib.MdConfig()
        cfg.atomsPerMolecule, cfg.numConstraintIterations, cfg.bondLength = atomsPerMolecule, numConstraintIterations, bondLength
        cfg.dt, cfg.rc, cfg.skin, cfg.sigma, cfg.epsilon, cfg.cappingDistance = dt, rc, skin, sigma, epsilon, cappingDistance
        cfg.maxNeighbors, cfg.integrator, cfg.cellSort, cfg.fullList = maxNeighbors, int(langevin), 1, 2
        cfg.zeta, cfg.temperature, cfg.seed = zeta, temperature, seed
        cfg.adress, cfg.doShift = int(adress), int(doShift)
        if weight is not None:
            C.memmove(C.byref(cfg.weight), C.byref(weight), C.sizeof(_lib.Weight))
        if thermo is not None:
            cfg.useThermoForce = 1
            cfg.thermoTargetDensity, cfg.thermoBinWidth, cfg.thermoModulation = thermo["targetDensity"], thermo["binWidth"], thermo["modulation"]
            cfg.thermoSampleInterval, cfg.thermoUpdateInterval = thermo["sampleInterval"], thermo["updateInterval"]
            cfg.thermoSmoothingSigma, cfg.thermoSmoothingIntensity = thermo["sigma"], thermo["range"]
        self.cfg, self.atoms = cfg, atoms
        gmin = np.ascontiguousarray(global_min, dtype=np.float64)
        gmax = np.ascontiguousarray(global_max, dtype=np.float64)
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        self.h = C.c_void_p()
        cut_arr = None if cuts is None else np.ascontiguousarray(cuts, dtype=np.float64)
        assert cut_arr is None or cut_arr.size == nranks + 1
        check(_lib.load().mrmd_b200_slab_create_cuts(C.byref(self.h), C.byref(cfg), gmin.ctypes.data, gmax.ctypes.data,
                                                    None if cut_arr is None else cut_arr.ctypes.data, rank, nranks,
                                                    uid.ctypes.data, atoms.h, None))

    def run(self, nsteps, timeForceKernel=False, stream=None):
        st = _lib.MdStats()
        check(_lib.load().mrmd_b200_slab_run(self.h, nsteps, int(timeForceKernel), C.byref(st),
                                            C.c_void_p(stream) if stream else None))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def run_host(self, nsteps, posHostPtr, velHostPtr, scalarsHostPtr=None, stream=None):
        st = _lib.MdStats()
        check(_lib.load().mrmd_b200_slab_run_host(self.h, nsteps, posHostPtr, velHostPtr, scalarsHostPtr, C.byref(st),
                                                 C.c_void_p(stream) if stream else None))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def setEnergyEveryStep(self, enabled):
        check(_lib.load().mrmd_b200_slab_set_energy_every_step(self.h, int(enabled)))

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            _lib.load().mrmd_b200_slab_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parity_check(rank, world, steps=40, mode="lj", langevin=True, stream=None, sites_x_per_rank=12, sites_yz=10):
    """x-slab run over `world` ranks against the single-GPU periodic run of the same global system (computed on rank 0's
    GPU): atoms matched by their global ids, positions / velocities compared atom by atom, pair counts and rebuild
    counts equal, energy within 1e-9.  torch.distributed (NCCL) must be initialised; collective.  Langevin noise is
    keyed by the global atom id (Philox counter), so the thermostatted trajectories are comparable too.
    mode: "lj", "adress" (slab region across a rank boundary, thermodynamic force updated during the run),
    "adress-cuts" (uneven slab widths), "tetramer" (4-atom molecules with SHAKE / RATTLE, spherical region).
    Returns the result dict on rank 0 ({"ok": ...}), {"ok": <broadcast flag>} elsewhere."""
    import torch
    import torch.distributed as dist

    from . import api

    tetramer = mode == "tetramer"
    adress = mode.startswith("adress") or tetramer
    nx, ny = sites_x_per_rank * world, sites_yz
    rng = np.random.default_rng(42)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(ny), indexing="ij"), axis=-1).reshape(-1, 3)
    apm = 4 if tetramer else 1
    if tetramer:
        spacing = 1.98425
        sites = (g + 0.5) * spacing + (rng.random(g.shape) - 0.5) * 0.3
        tet = np.array([(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)], dtype=np.float64) / (2.0 * np.sqrt(2.0))
        pos = (sites[:, None, :] + tet[None, :, :]).reshape(-1, 3)
        vmol = (rng.random(g.shape) - 0.5) * 1.5
        vmol -= vmol.mean(axis=0)
        vel = np.repeat(vmol, 4, axis=0)
        owner_pos = sites
    else:
        spacing = 1.25
        pos = (g + 0.5) * spacing + (rng.random(g.shape) - 0.5) * 0.5
        vel = (rng.random(g.shape) - 0.5) * 1.5
        vel -= vel.mean(axis=0)
        owner_pos = pos
    gmin, gmax = np.zeros(3), np.array([nx, ny, ny]) * spacing
    phys = dict(dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cappingDistance=0.7,
                maxNeighbors=40 if tetramer else 60, langevin=langevin, zeta=20.0, temperature=1.5, seed=4321)
    extra = {}
    if tetramer:
        extra = dict(adress=True, weight=api.Spherical(gmax / 2, 0.2 * gmax[1], 0.12 * gmax[1], 2), doShift=True,
                     atomsPerMolecule=4, numConstraintIterations=3, bondLength=1.0)
    elif adress:
        extra = dict(adress=True, weight=api.Slab(gmax / 2, 0.2 * gmax[0], 0.1 * gmax[0], 2), doShift=True,
                     thermo=dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=2, updateInterval=10,
                                 sigma=2.0, range=2.0))
    cuts = None
    if mode == "adress-cuts":
        cuts = balanced_cuts(gmin, gmax, world, 0.3 * gmax[0], 0.7 * gmax[0], 2.0, quantum=spacing, min_width=5.2)
        if world == 2:  # the balanced cut of a symmetric region is the middle: take an uneven one instead
            cuts = np.array([0.0, round(0.4 * nx) * spacing, gmax[0]])
    mols = select_slab(owner_pos, gmin, gmax, rank, world, cuts)
    mine = (mols[:, None] * apm + np.arange(apm)[None, :]).reshape(-1)
    atoms = api.Atoms.from_arrays(pos[mine], vel[mine], mass=1.0, relativeMass=1.0 / apm, ids=mine)
    uid = broadcast_unique_id(rank)
    md = SlabMolecularDynamics(atoms, gmin, gmax, rank, world, uid, cuts=cuts, **phys, **extra)
    st = md.run(steps, stream=stream)
    n = st["numLocal"]
    my = (atoms.get("id")[:n], atoms.getPos()[:n], atoms.getVel()[:n], st["pairInteractions"], st["rebuilds"], st["energy"])
    gathered = [None] * world
    dist.all_gather_object(gathered, my)
    md.close()
    ok, msg = True, {}
    if rank == 0:
        ids = np.concatenate([x[0] for x in gathered])
        all_pos = np.concatenate([x[1] for x in gathered])
        all_vel = np.concatenate([x[2] for x in gathered])
        pairs = gathered[0][3]  # already the sum over the ranks
        sub = api.Subdomain(gmin, gmax, phys["rc"] + phys["skin"])
        ref_atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0 / apm)
        ref = api.MolecularDynamics(ref_atoms, sub, cellSort=not tetramer, fullList=0 if tetramer else 2, **phys, **extra)
        rst = ref.run(steps, stream=stream)
        rid = ref_atoms.get("id")[:len(pos)]
        rp, rv = ref_atoms.getPos()[:len(pos)], ref_atoms.getVel()[:len(pos)]
        bijective = bool(len(ids) == len(pos) and np.array_equal(np.sort(ids), np.arange(len(pos))))
        msg = {"mode": mode, "ranks": world, "steps": steps, "langevin": bool(langevin), "atoms": int(len(ids)),
               "expected": int(len(pos)), "bijective": bijective}
        if bijective:
            a, b = np.argsort(ids), np.argsort(rid)
            box = gmax - gmin
            d = all_pos[a] - rp[b]
            d -= box * np.round(d / box)  # the two runs may hold different periodic images of an atom
            msg.update({"max_pos_err": float(np.abs(d).max()), "max_vel_err": float(np.abs(all_vel[a] - rv[b]).max())})
        msg.update({"pairs": int(pairs), "pairs_ref": int(rst["pairInteractions"]), "energy": st["energy"],
                    "energy_ref": rst["energy"], "rebuilds": [int(x[4]) for x in gathered],
                    "rebuilds_ref": int(rst["rebuilds"]), "per_rank": [int(len(x[0])) for x in gathered]})
        ok = bool(bijective and msg["max_pos_err"] < 1e-9 and msg["max_vel_err"] < 1e-7 and
                  msg["pairs"] == msg["pairs_ref"] and
                  abs(msg["energy"] - msg["energy_ref"]) <= 1e-9 * abs(msg["energy_ref"]) and
                  all(x[3] == pairs for x in gathered) and all(r == msg["rebuilds_ref"] for r in msg["rebuilds"]))
        msg = {"ok": ok, **msg}
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    return msg if rank == 0 else {"ok": bool(int(flag) == 1)}
