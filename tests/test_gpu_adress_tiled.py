"""Tiled AdResS fast path (mrmd_b200_adress_run_periodic: UpdateMolecules + LJ_IdealGas::run +
ContributeMoleculeForceToAtoms for one-atom molecules on the periodic tiled list) against the oracle's assembled
reference step over ghost molecules, and the tiled AdResS step loop against the generic one."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def system(n_side, seed, spacing=1.25, jitter=0.5):
    rng = np.random.default_rng(seed)
    sites = (np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), axis=-1).reshape(-1, 3) + 0.5) * spacing
    pos = sites + (rng.random(sites.shape) - 0.5) * jitter
    vel = (rng.random(sites.shape) - 0.5) * 1.5
    return pos, vel - vel.mean(axis=0), n_side * spacing


def weights(api, oracle, kind, box):
    center = [box / 2] * 3
    if kind == "slab":
        return api.Slab(center, 4.0, 3.0, 2), oracle.make_weight(oracle.WEIGHT_SLAB, center, 4.0, 3.0, 2)
    return api.Spherical(center, 3.0, 3.0, 2), oracle.make_weight(oracle.WEIGHT_SPHERICAL, center, 3.0, 3.0, 2)


@pytest.mark.parametrize("weight_kind,num_types", [("slab", 2), ("slab", 1), ("spherical", 2)])
def test_adress_run_periodic_vs_oracle(api, oracle, weight_kind, num_types):
    pos, _, box = system(14, 33)
    N = len(pos)
    rc, skin = 2.5, 0.1
    cutoff = rc + skin
    types = (np.arange(N) % num_types).astype(np.int64)
    nt = num_types
    capv, rcv = np.full(nt * nt, 0.7), np.full(nt * nt, rc)
    sig = np.array([1.0, 0.9, 0.9, 0.8])[:nt * nt] if nt == 2 else np.array([1.0])
    eps = np.array([1.0, 1.1, 1.1, 1.2])[:nt * nt] if nt == 2 else np.array([1.0])
    w, ow = weights(api, oracle, weight_kind, box)

    sub = api.Subdomain([0, 0, 0], [box] * 3, cutoff)
    atoms = api.Atoms.from_arrays(pos, None, mass=1.0, type=types, relativeMass=1.0)
    atoms.permute(api.LinkedCellList(0, N, [cutoff] * 3, sub.minCorner, sub.maxCorner))
    spos, stype = atoms.getPos()[:N], atoms.getType()[:N]
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, cutoff, 1.0, 60)
    lj = api.LJ_IdealGas(capv, rcv, sig, eps, True, numTypes=nt)

    # oracle: the reference's step over ghost molecules, atoms in the sorted order
    L = oracle.lib()
    osub = oracle.subdomain([0, 0, 0], [box] * 3, cutoff)
    oa = np.zeros(5 * N, dtype=oracle.ATOM)
    oa["pos"][:N], oa["mass"][:N], oa["type"][:N], oa["relMass"][:N] = spos, 1.0, stype, 1.0
    om = np.zeros(5 * N, dtype=oracle.MOLECULE)
    om["atomsOffset"][:N], om["numAtoms"][:N] = np.arange(N), 1
    L.or_update_molecules(om.ctypes.data, N, oa.ctypes.data, C.byref(ow))
    corr = np.zeros(len(oa), dtype=np.int64)
    out = np.zeros(2, dtype=np.int64)
    assert L.or_mr_ghost_create_xyz(om.ctypes.data, N, len(om), oa.ctypes.data, N, len(oa), C.byref(osub),
                                    corr.ctypes.data, out.ctypes.data) == 0
    mg, ag = int(out[0]), int(out[1])
    oc, on = oracle.verlet_build(om, 13, N + mg, 0, N, cutoff, 1.0, np.array(osub.minGhostCorner),
                                 np.array(osub.maxGhostCorner), half=True, width=40)
    assert vl.info()["totalPairs"] == 2 * int(oc[:N].sum())
    arrs = [np.ascontiguousarray(x) for x in (capv, rcv, sig, eps)]
    oh = L.or_adress_create(*[x.ctypes.data for x in arrs], nt, 1)

    for run in range(3):
        L.or_update_molecules(om.ctypes.data, N + mg, oa.ctypes.data, C.byref(ow))
        oa["force"] = 0.0
        om["force"] = 0.0
        nact = C.c_int64()
        oe = L.or_adress_run(oh, om.ctypes.data, N, oc.ctypes.data, on.ctypes.data, on.shape[1], oa.ctypes.data,
                             C.byref(nact))
        L.or_contribute_molecule_force(om.ctypes.data, N + mg, oa.ctypes.data)
        L.or_ghost_fold_force(oa.ctypes.data, N, ag, corr.ctypes.data)

        atoms.setForce(0.0)
        e = lj.run_periodic(atoms, vl, w)
        assert lj.lastNumPairs == nact.value and nact.value > 0
        assert abs(e - oe) <= 1e-11 * abs(oe)
        f = atoms.getForce()[:N]
        assert np.abs(f - oa["force"][:N]).max() <= FORCE_RTOL * np.abs(oa["force"][:N]).max()
        mean = lj.getMeanCompensationEnergy()
        omean = np.ctypeslib.as_array(oh.contents.meanCompensationEnergy, shape=(200, nt))
        assert np.count_nonzero(omean) > 0
        assert np.abs(mean - omean).max() <= 1e-11 * np.abs(omean).max()
    # the force is accumulated like the reference's
    lj.run_periodic(atoms, vl, w)
    assert np.abs(atoms.getForce()[:N] - 2 * oa["force"][:N]).max() <= 4 * FORCE_RTOL * np.abs(oa["force"][:N]).max()
    L.or_adress_destroy(oh)


def test_adress_run_periodic_rejects_region_at_boundary(api):
    pos, _, box = system(12, 3)
    cutoff = 2.6
    sub = api.Subdomain([0, 0, 0], [box] * 3, cutoff)
    atoms = api.Atoms.from_arrays(pos, None, relativeMass=1.0)
    atoms.permute(api.LinkedCellList(0, len(pos), [cutoff] * 3, sub.minCorner, sub.maxCorner))
    vl = api.FullVerletList()
    vl.build_periodic(atoms, sub, cutoff, 1.0, 60)
    lj = api.LJ_IdealGas(0.7, 2.5, 1.0, 1.0, True)
    w = api.Slab([box / 2] * 3, box - 4.0, 1.0, 1)  # the hybrid region overlaps the ghost layer
    with pytest.raises(RuntimeError, match="periodic boundary"):
        lj.run_periodic(atoms, vl, w)


@pytest.mark.parametrize("mode", [0, 2])
def test_adress_md_vs_oracle_loop(api, oracle, mode):
    """the AdResS step loop (Langevin, thermodynamic force sampled / updated / applied) against the oracle's loop of
    the same operators (oracle/md_loop.py:OracleAdressMD), 30 steps"""
    from oracle.md_loop import OracleAdressMD

    pos, vel, box = system(14, 5)
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], [box] * 3, 2.6)
    w, ow = weights(api, oracle, "slab", box)
    th = dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=2, updateInterval=10, sigma=2.0, range=2.0)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0)
    md = api.MolecularDynamics(atoms, sub, langevin=True, zeta=20.0, temperature=1.5, seed=99, cellSort=True,
                               fullList=mode, adress=True, weight=w, thermo=th)
    omd = OracleAdressMD(pos, vel, [box] * 3, ow, langevin=True, zeta=20.0, temperature=1.5, seed=99, thermo=th)
    st, res = md.run(30), omd.run(30)
    assert st["rebuilds"] == res["rebuilds"] and st["rebuilds"] >= 2
    assert st["pairInteractions"] == res["pairInteractions"]
    assert abs(st["energy"] - res["energy"]) <= 1e-8 * abs(res["energy"])
    assert np.abs(atoms.getPos()[:n] - omd.atoms["pos"][:n]).max() < 1e-9
    assert np.abs(atoms.getVel()[:n] - omd.atoms["vel"][:n]).max() < 1e-8


@pytest.mark.parametrize("thermo", [False, True])
def test_adress_md_tiled_vs_generic(api, thermo):
    """40 NVE AdResS steps: tiled (fullList=2) against the generic operator sequence (half molecule list over ghost
    molecules); same rebuild count and pair interactions, trajectories within roundoff amplification"""
    pos, vel, box = system(14, 11)
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], [box] * 3, 2.6)
    w = api.Slab([box / 2] * 3, 4.0, 3.0, 2)
    th = dict(targetDensity=0.512, binWidth=0.5, modulation=2.0, sampleInterval=2, updateInterval=10, sigma=2.0,
              range=2.0) if thermo else None
    out = {}
    for mode in (0, 2):
        atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0)
        md = api.MolecularDynamics(atoms, sub, langevin=False, cellSort=True, fullList=mode, adress=True, weight=w,
                                   thermo=th)
        st = md.run(40)
        out[mode] = (st, atoms.getPos()[:n], atoms.getVel()[:n])
    s0, p0, v0 = out[0]
    s2, p2, v2 = out[2]
    assert s0["rebuilds"] == s2["rebuilds"] and s0["rebuilds"] >= 2
    assert s0["pairInteractions"] == s2["pairInteractions"] and s2["pairInteractions"] > 0
    assert abs(s0["energy"] - s2["energy"]) <= 1e-9 * abs(s0["energy"])
    assert np.abs(p0 - p2).max() < 1e-9 and np.abs(v0 - v2).max() < 1e-8
